"""Minimal `chainer` stand-in so the reference's CALLERS (forward.py, tests/) import unchanged on a
box where Chainer v1 cannot be installed (no network; README.md:13 asks for "1.22.0+").

Scope: Variable, chainer.cuda.{available, cupy, to_gpu, to_cpu, get_array_module,
get_device_from_array, get_device_from_id}, chainer.serializers.load_npz/save_npz,
chainer.testing.{parameterize, product}, chainer.set_debug (SURVEY.md 8b census).  The model code
itself (models/) does not use Chainer functions/links at all: its math is libfrcnn_b200.so.
Installed on sys.path only when a real `chainer` is not importable (frcnn_b200.dropin.install()).
"""
import numpy as _np

from . import cuda, serializers, testing  # noqa: F401
from .variable import Variable  # noqa: F401

__version__ = "1.22.0-frcnn_b200-compat"
_debug = False


def set_debug(flag):
    global _debug
    _debug = bool(flag)


def is_debug():
    return _debug


class _Unavailable(object):
    def __init__(self, what):
        self._what = what

    def __getattr__(self, name):
        raise NotImplementedError("chainer.%s.%s: training-side Chainer API is outside the forward path "
                                  "(SURVEY.md 8f 'next' rows)" % (self._what, name))


computational_graph = _Unavailable("computational_graph")
optimizers = _Unavailable("optimizers")
