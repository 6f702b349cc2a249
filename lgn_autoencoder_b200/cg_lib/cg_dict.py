"""Clebsch-Gordan coefficients of SL(2,C) irreps (k, n), dimension (k+1)(n+1).

Public behaviour of the reference's CGDict (lgn/cg_lib/cg_dict.py:11-225): a mapping
``((k1,n1),(k2,n2)) -> {(k,n): matrix}`` for all k, n < maxdim, each matrix stored transposed and flattened to
``(dim(k,n), d1*d2)`` real fp64, shared between encoder, decoder and the Lorentz-D matrices of the
equivariance test (which index ``cg[((k,0),(0,n))][(k,n)]``).

The numbers are generated here from first principles (not taken from the reference): SU(2) coefficients by
Racah's closed form in exact integer arithmetic on doubled spins, and the SL(2,C) block by recoupling
(k/2 (x) n/2 -> l) on the two inputs and on the output.  Components of (k,n) are ordered
by l = |k-n|/2 .. (k+n)/2, m = -l..l ascending."""
from __future__ import annotations

import itertools
from fractions import Fraction
from functools import lru_cache
from math import factorial, sqrt

import numpy as np
import torch


@lru_cache(maxsize=None)
def su2_cg(J1: int, M1: int, J2: int, M2: int, J3: int, M3: int) -> float:
    """<j1 m1 j2 m2 | j3 m3> with all arguments DOUBLED (J = 2j, M = 2m), Condon-Shortley phases."""
    if M1 + M2 != M3 or not (abs(J1 - J2) <= J3 <= J1 + J2) or (J1 + J2 + J3) % 2:
        return 0.0
    if abs(M1) > J1 or abs(M2) > J2 or abs(M3) > J3 or (J1 + M1) % 2 or (J2 + M2) % 2 or (J3 + M3) % 2:
        return 0.0
    f = lambda x2: factorial(x2 // 2)   # argument is an even doubled integer
    pref = Fraction((J3 + 1) * f(J3 + J1 - J2) * f(J3 - J1 + J2) * f(J1 + J2 - J3), f(J1 + J2 + J3 + 2))
    pref *= f(J3 + M3) * f(J3 - M3) * f(J1 - M1) * f(J1 + M1) * f(J2 - M2) * f(J2 + M2)
    total = Fraction(0)
    for v in range(0, (J1 + J2 - J3) // 2 + 1):
        args = (J1 + J2 - J3 - 2 * v, J1 - M1 - 2 * v, J2 + M2 - 2 * v, J3 - J2 + M1 + 2 * v, J3 - J1 - M2 + 2 * v)
        if any(a < 0 for a in args):
            continue
        den = factorial(v)
        for a in args:
            den *= f(a)
        total += Fraction((-1) ** v, den)
    return float(total) * sqrt(pref.numerator / pref.denominator) if pref.denominator else 0.0


def _su2_block(J1: int, J2: int, J3: int) -> np.ndarray:
    """[j1+m1, j2+m2, j3+m3] array of SU(2) coefficients (doubled spins)."""
    out = np.zeros((J1 + 1, J2 + 1, J3 + 1))
    for a in range(J1 + 1):
        for b in range(J2 + 1):
            M1, M2 = 2 * a - J1, 2 * b - J2
            M3 = M1 + M2
            if abs(M3) <= J3 and (J3 + M3) % 2 == 0:
                out[a, b, (J3 + M3) // 2] = su2_cg(J1, M1, J2, M2, J3, M3)
    return out


def _recouple(k: int, n: int) -> np.ndarray:
    """(k/2) (x) (n/2) -> (+)_l l with the l multiplets stacked in ascending l: shape (k+1, n+1, (k+1)(n+1))."""
    return np.concatenate([_su2_block(k, n, L) for L in range(abs(k - n), k + n + 1, 2)], axis=-1)


def sl2c_cg(rep1, rep2, rep) -> np.ndarray:
    """CG block (d1, d2, d_out) coupling irreps rep1 (x) rep2 -> rep."""
    (k1, n1), (k2, n2), (k, n) = rep1, rep2, rep
    left = _su2_block(k1, k2, k)       # [a1, a2, a]
    right = _su2_block(n1, n2, n)      # [b1, b2, b]
    in1, in2, out = _recouple(k1, n1), _recouple(k2, n2), _recouple(k, n)   # [a, b, x]
    return np.einsum("pqa,rsb,prx,qsy,abz->xyz", left, right, in1, in2, out)


class CGDict:
    """Dictionary of CG matrices for all irreps with k, n < maxdim."""

    def __init__(self, maxdim=None, transpose=True, dtype=torch.float64, device=None):
        self.dtype = dtype
        self.device = device if device is not None else torch.device("cpu")
        self._transpose = transpose
        self._maxdim = None
        self._cg_dict = {}
        if maxdim is not None:
            self.update_maxdim(maxdim)

    @property
    def transpose(self):
        return self._transpose

    @property
    def maxdim(self):
        return self._maxdim

    def update_maxdim(self, new_maxdim):
        if self._maxdim is not None and new_maxdim <= self._maxdim:
            return self
        for k1, n1, k2, n2 in itertools.product(range(new_maxdim), repeat=4):
            key = ((k1, n1), (k2, n2))
            if key in self._cg_dict:
                continue
            entry = {}
            for k in range(abs(k1 - k2), k1 + k2 + 1, 2):
                for n in range(abs(n1 - n2), n1 + n2 + 1, 2):
                    block = sl2c_cg((k1, n1), (k2, n2), (k, n))
                    mat = block.reshape(-1, block.shape[-1])
                    if self._transpose:
                        mat = mat.T
                    entry[(k, n)] = torch.from_numpy(np.ascontiguousarray(mat)).to(device=self.device, dtype=self.dtype)
            self._cg_dict[key] = entry
        self._maxdim = new_maxdim
        return self

    def to(self, dtype=None, device=None):
        dtype = self.dtype if dtype is None else dtype
        device = self.device if device is None else device
        if dtype != self.dtype or device != self.device:
            self._cg_dict = {key: {k: v.to(device=device, dtype=dtype) for k, v in entry.items()} for key, entry in self._cg_dict.items()}
            self.dtype, self.device = dtype, device
        return self

    def keys(self):
        return self._cg_dict.keys()

    def values(self):
        return self._cg_dict.values()

    def items(self):
        return self._cg_dict.items()

    def __getitem__(self, idx):
        if not self:
            raise ValueError("CGDict has not been initialized")
        return self._cg_dict[idx]

    def __contains__(self, idx):
        return idx in self._cg_dict

    def __bool__(self):
        return self._maxdim is not None and self._maxdim >= 0
