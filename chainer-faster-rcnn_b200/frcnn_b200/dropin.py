"""Make the reference's import paths resolve to this build.

    from frcnn_b200 import dropin; dropin.install()

puts `chainer-faster-rcnn_b200/` first on sys.path (so `import models.faster_rcnn`,
`models.cpu_nms`, ... resolve to the B200 mirror instead of the reference's Chainer/Cython package)
and, only if a real Chainer is not importable, adds the minimal `chainer` / `cupy` stand-ins
(compat/).  After that the reference's forward.py and tests/ import unchanged.
"""
import importlib
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_COMPAT = os.path.join(_PKG_ROOT, "compat")


def install(force_compat=False):
    if _PKG_ROOT not in sys.path:
        sys.path.insert(0, _PKG_ROOT)
    need = force_compat
    if not need:
        try:
            importlib.import_module("chainer")
        except ImportError:
            need = True
    if need and _COMPAT not in sys.path:
        sys.path.insert(1, _COMPAT)
        for name in ("chainer", "cupy"):
            sys.modules.pop(name, None)
        importlib.import_module("cupy")
        importlib.import_module("chainer")
    return need
