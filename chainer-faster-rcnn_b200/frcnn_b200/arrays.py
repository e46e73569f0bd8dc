"""Array-family plumbing for the drop-in classes: accept what the reference's callers pass
(chainer.Variable wrapping numpy / cupy arrays, bare numpy, torch tensors), hand the kernels a
CUDA float32 torch tensor, and give results back in the caller's array family."""
import numpy as np
import torch

NUMPY, DEVICE, TORCH = "numpy", "cupy", "torch"


def raw(x):
    """Strip a Variable-like wrapper."""
    if isinstance(x, (np.ndarray, torch.Tensor)) or hasattr(x, "tensor"):
        return x
    return x.data if hasattr(x, "data") and not isinstance(x, (list, tuple)) else x


def family(x):
    x = raw(x)
    if isinstance(x, torch.Tensor):
        return TORCH
    if hasattr(x, "tensor"):          # compat cupy.ndarray
        return DEVICE
    return NUMPY


def to_device(x, dtype=torch.float32, device=None):
    x = raw(x)
    if hasattr(x, "tensor"):
        t = x.tensor
    elif isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=dev, dtype=dtype).contiguous()


def to_host_ints(x):
    x = raw(x)
    if hasattr(x, "tensor"):
        x = x.tensor
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x).reshape(-1)


def from_device(t, fam):
    if fam == TORCH:
        return t
    if fam == DEVICE:
        import cupy
        return cupy.ndarray(t)
    return t.detach().cpu().numpy()


def dtype_kind(x):
    x = raw(x)
    dt = x.dtype
    if isinstance(dt, torch.dtype):
        return "f" if dt.is_floating_point else "i"
    return np.dtype(dt).kind
