"""load_npz / save_npz for the reference's checkpoint format (forward.py:29): a NumPy .npz keyed by
link path, e.g. 'trunk/conv1_1/W', 'RPN/rpn_cls_score/b', 'fc6/W' (SURVEY.md 5)."""
import numpy as _np


def load_npz(filename, obj, strict=True):
    with _np.load(filename) as f:
        params = dict(obj.namedparams())
        for path, param in params.items():
            key = path.lstrip("/")
            if key not in f.files:
                if strict:
                    raise KeyError("%s: no array named %r" % (filename, key))
                continue
            arr = f[key]
            if param.data is not None and tuple(param.data.shape) != tuple(arr.shape):
                raise ValueError("%s: shape %s != %s" % (key, arr.shape, param.data.shape))
            param.data = _np.ascontiguousarray(arr, dtype=_np.float32)
    if hasattr(obj, "_params_changed"):
        obj._params_changed()


def save_npz(filename, obj, compression=True):
    arrays = {path.lstrip("/"): _np.asarray(p.data) for path, p in obj.namedparams()}
    (_np.savez_compressed if compression else _np.savez)(filename, **arrays)
