"""Minimal `cupy` stand-in backed by torch CUDA tensors.

Only what the reference's callers touch (tests/test_proposal_layer.py:36-47,
tests/test_region_proposal_network.py:29-36, forward.py:97-99): array creation, astype, asnumpy.
It exists so those files import and run unchanged on a box without CuPy; it is NOT a compute
library -- every computation happens in libfrcnn_b200.so.
"""
import numpy as _np
import torch as _torch

float32, float64, int32, int64 = _np.float32, _np.float64, _np.int32, _np.int64
_TORCH_OF = {_np.dtype("float32"): _torch.float32, _np.dtype("float64"): _torch.float64,
             _np.dtype("int32"): _torch.int32, _np.dtype("int64"): _torch.int64, _np.dtype("uint8"): _torch.uint8}
_NP_OF = {v: k for k, v in _TORCH_OF.items()}


class ndarray(object):
    """Device array: a thin view over a torch CUDA tensor (`.tensor`)."""

    def __init__(self, tensor):
        self.tensor = tensor

    shape = property(lambda s: tuple(s.tensor.shape))
    ndim = property(lambda s: s.tensor.dim())
    size = property(lambda s: s.tensor.numel())
    dtype = property(lambda s: _NP_OF[s.tensor.dtype])
    device = property(lambda s: s.tensor.device)

    def astype(self, dtype, copy=True):
        return ndarray(self.tensor.to(_TORCH_OF[_np.dtype(dtype)]))

    def get(self):
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, idx):
        return ndarray(self.tensor[idx])

    def reshape(self, *shape):
        return ndarray(self.tensor.reshape(*shape))

    def __repr__(self):
        return "cupy_compat.ndarray(shape=%s, dtype=%s)" % (self.shape, self.dtype)


def _dev(device=None):
    return _torch.device("cuda", _torch.cuda.current_device() if device is None else int(device))


def asarray(a, dtype=None):
    if isinstance(a, ndarray):
        return a if dtype is None else a.astype(dtype)
    if isinstance(a, _torch.Tensor):
        t = a.cuda()
    else:
        t = _torch.from_numpy(_np.ascontiguousarray(a)).to(_dev())
    return ndarray(t if dtype is None else t.to(_TORCH_OF[_np.dtype(dtype)]))


array = asarray


def asnumpy(a):
    return a.get() if isinstance(a, ndarray) else _np.asarray(a)


def zeros(shape, dtype=float32):
    return ndarray(_torch.zeros(shape, dtype=_TORCH_OF[_np.dtype(dtype)], device=_dev()))


def ones(shape, dtype=float32):
    return ndarray(_torch.ones(shape, dtype=_TORCH_OF[_np.dtype(dtype)], device=_dev()))


empty = zeros


class _Random(object):
    @staticmethod
    def rand(*shape):
        return ndarray(_torch.rand(*shape, dtype=_torch.float64, device=_dev()))

    @staticmethod
    def randn(*shape):
        return ndarray(_torch.randn(*shape, dtype=_torch.float64, device=_dev()))


random = _Random()
