"""Stub of `energyflow` (absent from this image; the reference's `utils` package imports it at module top for plotting /
jet-analysis code that is outside the hot path).  Every attribute is a no-op object."""
import sys
import types


class _Noop:
    """Callable, iterable, indexable, context-managing nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Noop()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Noop()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Noop()

    def __setitem__(self, k, v):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __len__(self):
        return 0

    def __bool__(self):
        return False


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return _Noop()


def _install(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        sub = ".".join(parts[:i])
        if sub not in sys.modules or not isinstance(sys.modules[sub], _StubModule):
            m = _StubModule(sub)
            m.__file__ = __file__
            sys.modules[sub] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)


sys.modules[__name__].__class__ = _StubModule
for _sub in ['emd']:
    _install(__name__ + "." + _sub)
