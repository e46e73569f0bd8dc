import numpy as _np
import torch as _torch

try:
    import cupy  # the stand-in next to this package (or a real CuPy)
except ImportError:  # pragma: no cover
    cupy = None

available = _torch.cuda.is_available()


class Device(object):
    def __init__(self, device_id=-1):
        self.id = device_id

    def use(self):
        if self.id >= 0:
            _torch.cuda.set_device(self.id)

    def __enter__(self):
        if self.id >= 0:
            self._prev = _torch.cuda.current_device()
            _torch.cuda.set_device(self.id)
        return self

    def __exit__(self, *exc):
        if self.id >= 0:
            _torch.cuda.set_device(self._prev)
        return False


def _raw(x):
    return getattr(x, "data", x) if not isinstance(x, (_np.ndarray, _torch.Tensor)) and not _is_dev(x) else x


def _is_dev(x):
    return cupy is not None and isinstance(x, cupy.ndarray)


def get_device_from_id(device_id):
    return Device(-1 if device_id is None else int(device_id))


def get_device_from_array(*arrays):
    for a in arrays:
        a = _raw(a)
        if _is_dev(a):
            return Device(a.device.index)
        if isinstance(a, _torch.Tensor) and a.is_cuda:
            return Device(a.device.index)
    return Device(-1)


get_device = get_device_from_array


def get_array_module(*arrays):
    for a in arrays:
        if _is_dev(_raw(a)):
            return cupy
    return _np


def to_gpu(array, device=None):
    if _is_dev(array):
        return array
    with Device(-1 if device is None else int(device)):
        return cupy.asarray(array)


def to_cpu(array):
    if _is_dev(array):
        return array.get()
    if isinstance(array, _torch.Tensor):
        return array.detach().cpu().numpy()
    return _np.asarray(array)
