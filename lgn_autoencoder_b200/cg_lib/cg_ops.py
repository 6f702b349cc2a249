"""Clebsch-Gordan products of GVecs (reference: lgn/cg_lib/cg_ops.py:10-298, cg_ops_tau.py:6-44).

Layer-level API for stand-alone calls and for configurations outside the fused path (maxdim 3, 'mix' pooling):
the product of every (irrep, irrep) pair is one launch of the generic term-list kernels in csrc/lgae_cg.cu.  The
training hot path of the LGAE at maxdim 2 never comes here: there the whole level is one kernel (csrc/lgae_level.cu)."""
import torch

from .. import layer_ops
from ..g_lib import GTau, GVec
from .cg_module import CGModule


def cg_product_tau(tau1, tau2, maxdim=float("inf")):
    """Multiplicities of a CG product when every channel of rep1 is paired with every channel of rep2
    (reference cg_ops_tau.py:6-44).  CGProduct.tau_out uses it with unit multiplicities because the product
    implemented by cg_product is channel-wise."""
    tau1, tau2 = GTau(tau1), GTau(tau2)
    tau = {}
    for (k1, n1), c1 in tau1.items():
        for (k2, n2), c2 in tau2.items():
            if max(k1, n1, k2, n2) >= maxdim:
                continue
            kmax = min(k1 + k2, maxdim - 1)
            nmax = min(n1 + n2, maxdim - 1)
            for k in range(abs(k1 - k2), int(kmax) + 1, 2):
                for n in range(abs(n1 - n2), int(nmax) + 1, 2):
                    tau[(k, n)] = tau.get((k, n), 0) + c1 * c2
    return GTau(tau)


def cg_product(cg_dict, rep1, rep2, maxdim=float("inf"), aggregate=False, ignore_check=False):
    """out[(k,n)][c] = H_(k,n) . vec(z1[c] (x) z2[c]); results for the same output irrep are concatenated on the
    channel axis in loop order (rep1 outer, rep2 inner).  With aggregate=True one operand carries an extra
    neighbour axis (2,B,N,N,C,d) and the product is summed over it: out_i = sum_j H (node_j (x) edge_ij).

    Runs on the term-list kernels of csrc/lgae_cg.cu (forward and adjoint); the Kronecker product the reference
    materialises (cg_ops.py:221-298) never exists."""
    if not ignore_check and cg_dict.maxdim is not None and maxdim < float("inf") and cg_dict.maxdim < maxdim:
        raise ValueError(f"CG dictionary maxdim ({cg_dict.maxdim}) is smaller than the requested maxdim ({maxdim})")
    keys1, keys2 = list(rep1.keys()), list(rep2.keys())
    top = max(max(k for k, _ in keys1) + max(k for k, _ in keys2), max(n for _, n in keys1) + max(n for _, n in keys2)) + 1
    max_dim = int(min(top, maxdim))
    out_keys, outs = layer_ops.cg_pairs(cg_dict, keys1, [rep1[k] for k in keys1], keys2, [rep2[k] for k in keys2], max_dim, aggregate)
    return GVec(dict(zip(out_keys, outs)), ignore_check=True)


class CGProduct(CGModule):
    """Module wrapper with tau bookkeeping (reference cg_ops.py:10-132)."""

    def __init__(self, tau1=None, tau2=None, aggregate=False, maxdim=float("inf"), cg_dict=None, dtype=None, device=None):
        self.aggregate = aggregate
        if maxdim == float("inf") and cg_dict:
            maxdim = cg_dict.maxdim
        elif maxdim == float("inf") and tau1 and tau2:
            maxdim = max(len(tau1), len(tau2))
        elif maxdim == float("inf"):
            raise ValueError("maxdim is not defined, and was unable to retrieve get maxdim from cg_dict or tau1 and tau2")
        super().__init__(cg_dict=cg_dict, maxdim=int(maxdim), device=device, dtype=dtype)
        self.set_taus(tau1, tau2)

    def set_taus(self, tau1=None, tau2=None):
        self._tau1 = GTau(tau1) if tau1 else None
        self._tau2 = GTau(tau2) if tau2 else None
        if self._tau1 and self._tau2:
            if not self._tau1.channels or self._tau1.channels != self._tau2.channels:
                raise ValueError(f"The number of fragments must be same for each part! {self._tau1} {self._tau2}")

    @property
    def tau1(self):
        return self._tau1

    @property
    def tau2(self):
        return self._tau2

    @property
    def tau_out(self):
        if not self._tau1 or not self._tau2:
            raise ValueError("Module not intialized with input type!")
        ones1 = {k: 1 for k, v in self._tau1.items() if v > 0}
        ones2 = {k: 1 for k, v in self._tau2.items() if v > 0}
        nchan = self._tau1.channels
        return GTau({k: nchan * t for k, t in cg_product_tau(ones1, ones2, maxdim=self.maxdim).items()})

    tau = tau_out

    def forward(self, rep1, rep2):
        if self._tau1 and GTau.from_rep(rep1) != self._tau1:
            raise ValueError("Input rep1 does not match predefined tau!")
        if self._tau2 and GTau.from_rep(rep2) != self._tau2:
            raise ValueError("Input rep2 does not match predefined tau!")
        return cg_product(self.cg_dict, rep1, rep2, maxdim=self.maxdim, aggregate=self.aggregate)
