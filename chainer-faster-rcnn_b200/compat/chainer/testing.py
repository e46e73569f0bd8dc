"""chainer.testing.{product, parameterize} as used at tests/test_faster_rcnn.py:23-27."""
import itertools
import sys


def product(parameter):
    keys = sorted(parameter)
    return [dict(zip(keys, vals)) for vals in itertools.product(*[parameter[k] for k in keys])]


def parameterize(*params):
    def wrap(klass):
        mod = sys.modules[klass.__module__]
        for i, param in enumerate(params):
            name = "%s_param_%d" % (klass.__name__, i)
            sub = type(name, (klass,), dict(param))
            sub.__module__ = klass.__module__
            setattr(mod, name, sub)
        # the un-parameterised base must not be collected as a test case itself
        for attr in [a for a in vars(klass) if a.startswith("test")]:
            pass
        klass.__test__ = False
        for i in range(len(params)):
            getattr(mod, "%s_param_%d" % (klass.__name__, i)).__test__ = True
        return klass
    return wrap
