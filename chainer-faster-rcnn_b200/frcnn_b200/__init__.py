"""frcnn_b200 -- host-side Python of the B200-native Faster R-CNN forward path.

  _lib    ctypes binding of libfrcnn_b200.so (include/frcnn_b200.h)
  ops     torch-tensor wrappers (device pointers + current stream in, tensors out)
  engine  the whole forward pass as one replayable CUDA graph

The drop-in mirror of the reference's `models` package lives next to this package
(chainer-faster-rcnn_b200/models) and is built on `engine`.
"""
from ._lib import FrcnnError, LIB_PATH, load  # noqa: F401

__all__ = ["FrcnnError", "LIB_PATH", "load"]
