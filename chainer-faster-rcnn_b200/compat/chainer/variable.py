import numpy as _np


class Variable(object):
    """Array holder with the attributes the reference's callers and type checks read:
    .data .shape .dtype .ndim .size .volatile, to_gpu()/to_cpu()."""

    def __init__(self, data=None, volatile=False, name=None):
        self.data = data
        self.volatile = volatile
        self.name = name

    shape = property(lambda s: tuple(s.data.shape))
    ndim = property(lambda s: len(s.data.shape))
    size = property(lambda s: int(_np.prod(s.data.shape)))

    @property
    def dtype(self):
        return _np.dtype(self.data.dtype) if not hasattr(self.data.dtype, "is_floating_point") else \
            _np.dtype(str(self.data.dtype).replace("torch.", ""))

    def to_gpu(self, device=None):
        from . import cuda
        self.data = cuda.to_gpu(self.data, device)

    def to_cpu(self):
        from . import cuda
        self.data = cuda.to_cpu(self.data)

    def __len__(self):
        return self.data.shape[0]

    def __repr__(self):
        return "variable(shape=%s)" % (self.shape,)
