"""Tiny parameter containers with the attribute / path layout of the reference's Chainer links
(`model.trunk.conv1_1.W.data`, namedparams() -> '/trunk/conv1_1/W'), so checkpoints in the
reference's .npz format (forward.py:29) load by name.  No autograd, no compute: the arrays are
host float32 masters that get repacked once into the kernels' bf16 layouts."""
import numpy as np


class Param(object):
    """One named array.  ASSIGNING `param.data = array` (what chainer.serializers.load_npz and most callers do) marks the
    owning link as changed, so that packed device copies are rebuilt; mutating the array IN PLACE
    (`param.data[...] = v`) cannot be seen from here -- call `link._params_changed()` afterwards."""

    def __init__(self, data, owner=None):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_owner", owner)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        object.__setattr__(self, "_data", value)
        if self._owner is not None:
            self._owner._bump()

    shape = property(lambda s: s._data.shape)


class Link(object):
    def __init__(self):
        object.__setattr__(self, "_param_names", [])
        object.__setattr__(self, "_child_names", [])
        object.__setattr__(self, "_version", 0)
        object.__setattr__(self, "_device_id", -1)

    def add_param(self, name, array):
        self._param_names.append(name)
        object.__setattr__(self, name, Param(np.ascontiguousarray(array, dtype=np.float32), owner=self))

    def add_link(self, name, link):
        self._child_names.append(name)
        object.__setattr__(self, name, link)

    def namedparams(self, prefix=""):
        for n in self._param_names:
            yield prefix + "/" + n, getattr(self, n)
        for c in self._child_names:
            for item in getattr(self, c).namedparams(prefix + "/" + c):
                yield item

    def params(self):
        return [p for _, p in self.namedparams()]

    def param_dict(self, prefix=""):
        return {k.lstrip("/"): p.data for k, p in self.namedparams(prefix)}

    def _bump(self):
        object.__setattr__(self, "_version", self._version + 1)

    def _params_changed(self):
        """Declare that parameter arrays of this link or of any link below it were modified in place."""
        self._bump()
        for c in self._child_names:
            getattr(self, c)._params_changed()

    def version_key(self):
        """Changes whenever this link OR ANY DESCENDANT changed (a bump on a sub-link -- load_npz(path, model.trunk),
        rpn._params_changed(), param.data = ... -- is therefore seen by every cache held further up)."""
        return (self._version,) + tuple(getattr(self, c).version_key() for c in self._child_names)

    # -- device placement: the math always runs on the GPU; these only record the caller's intent so
    #    `xp` and the returned array family behave like the reference's links.
    def to_gpu(self, device=None):
        object.__setattr__(self, "_device_id", 0 if device is None else int(device))
        for c in self._child_names:
            getattr(self, c).to_gpu(device)
        return self

    def to_cpu(self):
        object.__setattr__(self, "_device_id", -1)
        for c in self._child_names:
            getattr(self, c).to_cpu()
        return self

    @property
    def xp(self):
        if self._device_id >= 0:
            import cupy
            return cupy
        return np


def conv_link(cin, cout, k, std, rng=np.random):
    l = Link()
    l.add_param("W", rng.normal(0.0, std, size=(cout, cin, k, k)))
    l.add_param("b", np.zeros(cout))
    return l


def linear_link(cin, cout, std, rng=np.random):
    l = Link()
    l.add_param("W", rng.normal(0.0, std, size=(cout, cin)))
    l.add_param("b", np.zeros(cout))
    return l
