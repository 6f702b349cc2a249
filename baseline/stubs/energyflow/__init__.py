"""Stub of `energyflow` (absent from this image; the reference's `utils` package imports it at module top for plotting /
jet-analysis code that is outside the hot path).  Every attribute is a no-op object."""
import sys
import types


_PAIR_RETURNING = ("subplots", "get_legend_handles_labels")   # call sites unpack two values


class _Noop:
    """Callable, iterable, indexable, context-managing nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Noop()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name in _PAIR_RETURNING:
            return lambda *a, **k: (_Noop(), _Noop())
        return _Noop()

    def __iter__(self):
        return iter([_Noop() for _ in range(16)])   # `for ax, x in zip(axs, data)` must bind its loop variables

    def __getitem__(self, k):
        return _Noop()

    def __setitem__(self, k, v):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __bool__(self):
        return False

    def __len__(self):
        return 0

    def __float__(self):
        return 0.0

    def __int__(self):
        return 0

    def __index__(self):
        return 0

    def __array__(self, dtype=None, copy=None):
        import numpy
        return numpy.zeros(1, dtype=dtype or float)


def _arith(self, *a, **k):
    return _Noop()


for _ARITH in ("add", "radd", "sub", "rsub", "mul", "rmul", "truediv", "rtruediv", "pow", "rpow", "neg", "pos", "abs", "matmul", "rmatmul",
               "floordiv", "rfloordiv", "mod", "rmod"):
    setattr(_Noop, f"__{_ARITH}__", _arith)
for _CMP in ("lt", "le", "gt", "ge"):
    setattr(_Noop, f"__{_CMP}__", lambda self, other: False)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        if name in _PAIR_RETURNING:
            return lambda *a, **k: (_Noop(), _Noop())
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
