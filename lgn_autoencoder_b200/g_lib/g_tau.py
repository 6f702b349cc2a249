"""GTau: multiplicity (number of channels) of every SL(2,C) irrep (k, n) in a representation.
Mirrors the public behaviour of the reference's lgn/g_lib/g_tau.py:6-170."""
from __future__ import annotations


class GTau:
    def __init__(self, tau):
        if isinstance(tau, GTau):
            tau = tau._tau
        elif hasattr(tau, "tau") and not isinstance(tau, dict):
            tau = tau.tau._tau
        if not isinstance(tau, dict):
            raise ValueError(f"GTau expects a dict {{(k, n): channels}}, got {type(tau)}")
        clean = {}
        for key, val in tau.items():
            if not (isinstance(key, tuple) and len(key) == 2 and all(isinstance(x, int) for x in key)):
                raise ValueError(f"keys of a GTau must be tuples (k, n) of ints, got {key!r}")
            if int(val) != val:
                raise ValueError(f"multiplicities must be integers, got {val!r}")
            if val:
                clean[key] = int(val)
        self._tau = clean

    @property
    def maxdim(self):
        return max(max(k) for k in self._tau) + 1

    def keys(self):
        return self._tau.keys()

    def values(self):
        return self._tau.values()

    def items(self):
        return self._tau.items()

    def __iter__(self):
        yield from self._tau.items()

    def __getitem__(self, key):
        return self._tau[key]

    def __setitem__(self, key, val):
        self._tau[key] = val

    def __contains__(self, key):
        return key in self._tau

    def __len__(self):
        return len(self._tau)

    def __eq__(self, other):
        other = other._tau if isinstance(other, GTau) else dict(other)
        return self._tau == {k: v for k, v in other.items() if v}

    def __hash__(self):
        return hash(tuple(sorted(self._tau.items())))

    def get(self, key, default=0):
        return self._tau.get(key, default)

    @staticmethod
    def cat(tau_list):
        """Channel-wise concatenation: multiplicities add."""
        out = {}
        for tau in tau_list:
            for key, val in GTau(tau).items():
                out[key] = out.get(key, 0) + val
        return GTau(out)

    def __and__(self, other):
        return GTau.cat([self, other])

    def __rand__(self, other):
        return GTau.cat([other, self])

    def __add__(self, other):
        return GTau.cat([self, other])

    def __radd__(self, other):
        return self if other == 0 else GTau.cat([other, self])

    def __str__(self):
        return str(self._tau)

    __repr__ = __str__

    @staticmethod
    def from_rep(rep):
        if rep is None:
            return GTau({})
        if isinstance(rep, GTau):
            return rep
        if hasattr(rep, "tau"):
            return GTau(rep.tau)
        return GTau({key: part.shape[-2] for key, part in rep.items()})

    @property
    def tau(self):
        return self

    @property
    def channels(self):
        vals = set(self._tau.values())
        return vals.pop() if len(vals) == 1 else None

    def copy(self):
        return GTau(dict(self._tau))
