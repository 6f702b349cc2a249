"""Basis changes, Minkowski norms and zonal functions (reference: lgn/cg_lib/zonal_functions.py:10-446).
Layer-level composites on device tensors; the fused kernels compute the same quantities in registers."""
import functools
import math

import torch

from ..g_lib import GVec
from .cg_module import CGModule
from .cg_ops import cg_product

R2 = 1.0 / math.sqrt(2.0)


def eps(data):
    """1e-16 for fp64 (the only supported dtype; reference zonal_functions.py:441-446)."""
    return 1e-16 if data.dtype == torch.float64 else None


@functools.lru_cache(maxsize=None)
def _cartesian4(dtype, device):
    """Constant basis-change matrix, built once per (dtype, device): no host-to-device copy per call (CUDA-graph capturable)."""
    re = [[1, 0, 0, 0], [0, R2, 0, 0], [0, 0, 0, 1], [0, -R2, 0, 0]]
    im = [[0, 0, 0, 0], [0, 0, -R2, 0], [0, 0, 0, 0], [0, 0, -R2, 0]]
    return torch.complex(torch.tensor(re, dtype=dtype, device=device), torch.tensor(im, dtype=dtype, device=device))


def normsq4(p):
    """Minkowski square of real Cartesian 4-vectors (reference zonal_functions.py:201-218)."""
    # Near-massless particles make this a cancellation (p^2 ~ 1e-10 E^2), so the reference's CPU rounding sequence is kept
    # explicitly: squares, then the four terms added left to right (no device-side reduction, no FMA contraction).
    psq = p * p
    s = ((psq[..., 0] + psq[..., 1]) + psq[..., 2]) + psq[..., 3]
    return 2 * psq[..., 0] - s


def p_to_rep(p):
    """Real Cartesian (...,4) -> GVec {(1,1): (2,...,1,4)} in the canonical basis."""
    m = _cartesian4(p.dtype, p.device)
    z = torch.matmul(p.to(m.dtype), m.T)
    return GVec({(1, 1): torch.stack((z.real, z.imag), 0).unsqueeze(-2)}, ignore_check=True)


def p_cplx_to_rep(p):
    """Complex Cartesian (2,...,4) -> GVec {(1,1): (2,...,4)} (no channel axis added, as in the reference)."""
    m = _cartesian4(p.dtype, p.device)
    z = torch.matmul(torch.complex(p[0], p[1]), m.T)
    return GVec({(1, 1): torch.stack((z.real, z.imag), 0)}, ignore_check=True)


def rep_to_p(rep):
    """Canonical (1,1) part -> complex Cartesian (2,...,4): the Hermitian transpose of p_to_rep."""
    part = rep[(1, 1)] if not torch.is_tensor(rep) else rep
    m = _cartesian4(part.dtype, part.device)
    z = torch.matmul(torch.complex(part[0], part[1]), m.conj())
    return torch.stack((z.real, z.imag), 0)


@functools.lru_cache(maxsize=None)
def _metric_cached(dtype, device):
    return torch.tensor([[1.0, 0, 0, 0], [0, 0, 0, 1.0], [0, 0, -1.0, 0], [0, 1.0, 0, 0]], dtype=dtype, device=device)


def metric(dtype=torch.float64, device=None):
    """Canonical-basis metric (built once per dtype and device; callers get a copy they may modify)."""
    return _metric_cached(dtype, device).clone()


def repdot(rep1, rep2):
    """Invariant complex bilinear form on canonical 4-vectors -> (2, ..., 1)."""
    a, b = (r[(1, 1)] if not torch.is_tensor(r) else r for r in (rep1, rep2))
    g = metric(a.dtype, a.device).to(torch.complex128 if a.dtype == torch.float64 else torch.complex64)
    z = torch.einsum("...a,ab,...b->...", torch.complex(a[0], a[1]), g, torch.complex(b[0], b[1]))
    return torch.stack((z.real, z.imag), 0).unsqueeze(-1)


def _higher(zf, p11, max_zf, cg_dict):
    for l in range(2, max_zf + 1):
        new = cg_product(cg_dict, GVec({(l - 1, l - 1): zf[(l - 1, l - 1)]}, ignore_check=True), GVec({(1, 1): p11}, ignore_check=True),
                         maxdim=l + 1, ignore_check=True)[(l, l)]
        zf[(l, l)] = new * math.sqrt(2 * l / (l + 1))


def zonal_functions4(cg_dict, p, maxdim, normalize=False):
    """Real Cartesian momenta -> (zonal GVec, norm, norm_sq)."""
    e = eps(p)
    norm_sq = normsq4(p).unsqueeze(-1) + e
    norm = torch.where(norm_sq != 0, norm_sq / norm_sq.abs().sqrt(), norm_sq)
    p11 = p_to_rep(p)[(1, 1)]
    zf = {(0, 0): torch.ones(p11.shape[:-1] + (1,), dtype=p.dtype, device=p.device), (1, 1): p11}
    _higher(zf, p11, maxdim, cg_dict)
    return GVec(zf, ignore_check=True), norm.squeeze(-1), norm_sq.squeeze(-1)


def zonal_functions_canonical(cg_dict, p, maxdim, normalize=False):
    """Complex canonical momenta (2,...,4)."""
    e = eps(p)
    norm_sq = repdot(p, p) + e
    norm = norm_sq / norm_sq.abs().sqrt()
    zf = {(0, 0): torch.ones(p.shape[:-1] + (1,), dtype=p.dtype, device=p.device), (1, 1): p}
    _higher(zf, p, maxdim, cg_dict)
    zf = {k: v.unsqueeze(-2) for k, v in zf.items()}
    return GVec(zf, ignore_check=True), norm.squeeze(-1), norm_sq.squeeze(-1)


def zonal_functions(cg_dict, p, maxdim, normalize=False, basis="cartesian"):
    if basis == "cartesian":
        return zonal_functions4(cg_dict, p, maxdim, normalize)
    if basis == "canonical":
        return zonal_functions_canonical(cg_dict, p, maxdim, normalize)
    raise ValueError(f"basis must be 'cartesian' or 'canonical', got {basis}")


def zonal_functions_rel(cg_dict, p1, p2, maxdim, normalize=False, basis="cartesian"):
    rel = p1.unsqueeze(-2) - p2.unsqueeze(-3)
    return zonal_functions(cg_dict, rel, maxdim, normalize, basis)


class ZonalFunctions(CGModule):
    def __init__(self, maxdim, normalize=False, basis="cartesian", cg_dict=None, dtype=None, device=None):
        self.normalize, self.basis, self._zf_maxdim = normalize, basis, maxdim
        super().__init__(cg_dict=cg_dict, maxdim=maxdim + 1, device=device, dtype=dtype)

    def forward(self, p):
        return zonal_functions(self.cg_dict, p, self._zf_maxdim, self.normalize, self.basis)


class ZonalFunctionsRel(CGModule):
    def __init__(self, maxdim, normalize=False, basis="cartesian", cg_dict=None, dtype=None, device=None):
        self.normalize, self.basis, self._zf_maxdim = normalize, basis, maxdim
        super().__init__(cg_dict=cg_dict, maxdim=maxdim + 1, device=device, dtype=dtype)

    def forward(self, p1, p2):
        return zonal_functions_rel(self.cg_dict, p1, p2, self._zf_maxdim, self.normalize, self.basis)
