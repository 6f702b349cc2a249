"""Planar-complex irrep containers: dicts {(k, n): tensor} whose leading axis is (re, im).

  GVec     part shape (2, *batch, channels, (k+1)(n+1))     reference: lgn/g_lib/g_vec.py:11-70
  GScalar  part shape (2, *batch, channels)                 reference: lgn/g_lib/g_scalar.py:7-44
  GWeight  part shape (2, channels_out, channels_in)        reference: lgn/g_lib/g_weight.py:10-47

Iterating a container yields (key, part) pairs, as in the reference (g_tensor.py:191-196); the rotation
helpers rebuild containers with ``rep.__class__({...})`` (g_lib/rotations.py:36-51), which is supported."""
from __future__ import annotations

import torch

from .g_tau import GTau


def irrep_dim(key):
    return (key[0] + 1) * (key[1] + 1)


class GTensor:
    zdim = 0
    bdim = None
    cdim = None
    rdim = None

    def __init__(self, data, ignore_check=False):
        if isinstance(data, GTensor):
            data = data._data
        if not isinstance(data, dict):
            raise ValueError("data must be a dictionary {(k, n): tensor}")
        data = {k: v for k, v in data.items() if isinstance(k, tuple) and torch.is_tensor(v) and v.numel() > 0}
        if not ignore_check:
            self.check_data(data)
        self._data = data

    # ---- to be specialised --------------------------------------------------------------------------
    def check_data(self, data):
        for key, val in data.items():
            if val.shape[self.zdim] != 2:
                raise ValueError(f"complex axis of part {key} must have size 2, got shape {tuple(val.shape)}")

    # ---- dict behaviour ---------------------------------------------------------------------------------
    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def pop(self, key):
        return self._data.pop(key)

    def __iter__(self):
        yield from self._data.items()

    def __len__(self):
        return len(self._data)

    def __contains__(self, key):
        return key in self._data

    def __getitem__(self, key):
        if type(key) is not tuple:
            raise ValueError(f"Keys of G tensors must be tuples of ints! {key}")
        return self._data[key]

    def __setitem__(self, key, val):
        self._data[key] = val

    def __eq__(self, other):
        if set(self.keys()) != set(other.keys()):
            return False
        return all(bool((self[k] == other[k]).all()) for k in self.keys())

    __hash__ = None

    def __str__(self):
        return str(dict(self._data))

    __repr__ = __str__

    # ---- geometry -------------------------------------------------------------------------------------------
    @property
    def data(self):
        return self._data

    @property
    def maxdim(self):
        return max(max(k) for k in self._data) + 1

    @property
    def tau(self):
        return GTau({key: part.shape[self.cdim] for key, part in self._data.items()})

    @property
    def channels(self):
        return self.tau.channels

    @property
    def shapes(self):
        return {key: part.shape for key, part in self._data.items()}

    @property
    def bshape(self):
        shapes = {tuple(part.shape[1:self.cdim]) for part in self._data.values()}
        if len(shapes) != 1:
            raise ValueError(f"parts have different batch shapes: {shapes}")
        return shapes.pop()

    @property
    def device(self):
        return next(iter(self._data.values())).device

    @property
    def dtype(self):
        return next(iter(self._data.values())).dtype

    def truncate(self, maxdim):
        return type(self)({k: v for k, v in self._data.items() if max(k) < maxdim})

    # ---- tensor-like helpers ----------------------------------------------------------------------------------
    def _map(self, fn):
        return type(self)({k: fn(v) for k, v in self._data.items()}, ignore_check=True)

    def to(self, *args, **kwargs):
        self._data = {k: v.to(*args, **kwargs) for k, v in self._data.items()}
        return self

    def cpu(self):
        return self.to("cpu")

    def cuda(self, **kwargs):
        self._data = {k: v.cuda(**kwargs) for k, v in self._data.items()}
        return self

    def double(self):
        return self.to(torch.float64)

    def float(self):
        return self.to(torch.float32)

    def clone(self):
        return self._map(torch.clone)

    def detach(self):
        return self._map(torch.Tensor.detach)

    def requires_grad_(self, requires_grad=True):
        for v in self._data.values():
            v.requires_grad_(requires_grad)
        return self

    @property
    def grad(self):
        return type(self)({k: v.grad for k, v in self._data.items() if v.grad is not None}, ignore_check=True)

    def abs(self):
        return self._map(torch.abs)

    def max(self):
        return {k: v.max() for k, v in self._data.items()}

    def min(self):
        return {k: v.min() for k, v in self._data.items()}

    def squeeze(self, dim):
        return self._map(lambda t: t.squeeze(dim))

    def unsqueeze(self, dim):
        return self._map(lambda t: t.unsqueeze(dim))

    @staticmethod
    def allclose(rep1, rep2, **kwargs):
        return set(rep1.keys()) == set(rep2.keys()) and all(torch.allclose(rep1[k], rep2[k], **kwargs) for k in rep1.keys())

    # ---- arithmetic (see g_torch) ----------------------------------------------------------------------------
    def __add__(self, other):
        from . import g_torch
        return g_torch.add(self, other)

    __radd__ = __add__

    def __sub__(self, other):
        from . import g_torch
        return g_torch.sub(self, other)

    def __mul__(self, other):
        from . import g_torch
        return g_torch.mul(self, other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        from . import g_torch
        return g_torch.div(self, other)

    add, sub, mul, div, complex_mul = __add__, __sub__, __mul__, __truediv__, __mul__

    def __and__(self, other):
        from . import g_torch
        return g_torch.cat([self, other])

    def __rand__(self, other):
        from . import g_torch
        return g_torch.cat([other, self])

    @classmethod
    def _shape(cls, batch, key, channels):
        raise NotImplementedError

    @classmethod
    def _make(cls, fn, tau, batch, device=None, dtype=None, requires_grad=False):
        batch = (batch,) if isinstance(batch, int) else tuple(batch)
        return cls({key: fn(cls._shape(batch, key, ch), device=device, dtype=dtype, requires_grad=requires_grad)
                    for key, ch in GTau(tau).items()})

    @classmethod
    def rand(cls, tau, batch, device=None, dtype=None, requires_grad=False):
        return cls._make(torch.rand, tau, batch, device, dtype, requires_grad)

    @classmethod
    def randn(cls, tau, batch, device=None, dtype=None, requires_grad=False):
        return cls._make(torch.randn, tau, batch, device, dtype, requires_grad)

    @classmethod
    def zeros(cls, tau, batch, device=None, dtype=None, requires_grad=False):
        return cls._make(torch.zeros, tau, batch, device, dtype, requires_grad)

    @classmethod
    def ones(cls, tau, batch, device=None, dtype=None, requires_grad=False):
        return cls._make(torch.ones, tau, batch, device, dtype, requires_grad)


class GVec(GTensor):
    cdim = -2
    rdim = -1
    bdim = slice(1, -2)

    def check_data(self, data):
        super().check_data(data)
        for key, val in data.items():
            if val.dim() < 3 or val.shape[-1] != irrep_dim(key):
                raise ValueError(f"part {key} must have last dimension {irrep_dim(key)}, got shape {tuple(val.shape)}")
        shapes = {tuple(v.shape[1:-2]) for v in data.values()}
        if len(shapes) > 1:
            raise ValueError(f"all parts of a GVec must share the batch shape, got {shapes}")

    @classmethod
    def _shape(cls, batch, key, channels):
        return (2,) + batch + (channels, irrep_dim(key))


class GScalar(GTensor):
    cdim = -1
    rdim = None
    bdim = slice(1, -1)

    @classmethod
    def _shape(cls, batch, key, channels):
        return (2,) + batch + (channels,)


class GWeight(GTensor):
    cdim = 2
    rdim = None
    bdim = None

    def check_data(self, data):
        super().check_data(data)
        for key, val in data.items():
            if val.dim() != 3:
                raise ValueError(f"weight {key} must have shape (2, C_out, C_in), got {tuple(val.shape)}")

    @property
    def tau_in(self):
        return GTau({k: v.shape[2] for k, v in self._data.items()})

    @property
    def tau_out(self):
        return GTau({k: v.shape[1] for k, v in self._data.items()})

    tau = tau_in

    @classmethod
    def _wmake(cls, fn, tau_in, tau_out, device=None, dtype=None, requires_grad=False):
        tau_in, tau_out = GTau(tau_in), GTau(tau_out)
        return cls({k: fn((2, tau_out[k], tau_in[k]), device=device, dtype=dtype, requires_grad=requires_grad) for k in tau_in.keys()})

    @classmethod
    def rand(cls, tau_in, tau_out, device=None, dtype=None, requires_grad=False):
        return cls._wmake(torch.rand, tau_in, tau_out, device, dtype, requires_grad)

    @classmethod
    def randn(cls, tau_in, tau_out, device=None, dtype=None, requires_grad=False):
        return cls._wmake(torch.randn, tau_in, tau_out, device, dtype, requires_grad)

    @classmethod
    def zeros(cls, tau_in, tau_out, device=None, dtype=None, requires_grad=False):
        return cls._wmake(torch.zeros, tau_in, tau_out, device, dtype, requires_grad)

    @classmethod
    def ones(cls, tau_in, tau_out, device=None, dtype=None, requires_grad=False):
        return cls._wmake(torch.ones, tau_in, tau_out, device, dtype, requires_grad)
