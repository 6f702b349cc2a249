"""Reference-faithful algorithmic FLOP counts of the LGAE hot path: the numerator of every
roofline figure in bench.py.  Real flops only: complex mul = 6, complex add = 2, complex x real-coefficient MAC = 4,
reciprocal / sqrt / div = 1, only non-zero CG coefficients.  The counts describe the REFERENCE's algorithm
(O(N^2) decoder aggregation, full radial functions in every level); algebraic savings made by the kernels do not
shrink them."""
from __future__ import annotations

from typing import Dict, List

# maxdim-2 CG bookkeeping for a level with node irreps {(0,0),(1,1)} x edge irreps {(0,0),(1,1)}:
#   products: S.e0 (1 comp), V.e0 (4), S.e1 (4), V.e1 -> (0,0) (1 comp from 4 products)
_D_EDGE = 5            # sum_l d_l of the zonal functions (1 + 4)
_AGG_PER_EDGE_CH = 150  # 6*#products + 4*nnz + 2*sum(out comps), SURVEY 8(d)
_POW_PER_NODE_CH = 6 * 10 + 4 * 14   # self product: same structure without the neighbour sum


def level_flops(n: int, c_in: int, c_out: int, k_basis: int, encoder: bool, mlp_width: int, mlp_hidden: int) -> Dict[str, float]:
    """Forward flops of one LGN level for one jet, split by stage."""
    out = {}
    out["radial"] = n * n * (6 * k_basis + 2 * k_basis * 2 * c_in * 2) if encoder else 0.0
    out["edge"] = n * n * c_in * 6 * _D_EDGE
    out["aggregate"] = n * n * c_in * _AGG_PER_EDGE_CH
    out["power"] = n * c_in * _POW_PER_NODE_CH
    out["mix"] = n * 8 * (1 + 4) * c_out * 5 * c_in
    w = mlp_width
    out["mlp"] = n * 2 * (2 * c_out * w + (mlp_hidden - 1) * w * w + w * 2 * c_out) if mlp_hidden else 0.0
    return out


def model_flops(n: int, channels: List[int], k_basis: int, encoder: bool, mlp_width_mul: int, mlp_hidden: int) -> Dict[str, float]:
    tot: Dict[str, float] = {}
    for l in range(len(channels) - 1):
        lf = level_flops(n, channels[l], channels[l + 1], k_basis, encoder, mlp_width_mul * 2 * channels[l + 1], mlp_hidden)
        for k, v in lf.items():
            tot[k] = tot.get(k, 0.0) + v
    return tot


def step_flops_per_jet(n: int, enc_channels: List[int], dec_channels: List[int], num_basis_fn: int = 10, mlp_width_mul: int = 6,
                       mlp_hidden: int = 6, backward: bool = True) -> float:
    """Encoder + decoder forward (x3 with the backward pass)."""
    e = model_flops(n, enc_channels, 2 * num_basis_fn, True, mlp_width_mul, mlp_hidden)
    d = model_flops(n, dec_channels, 2 * num_basis_fn, False, mlp_width_mul, mlp_hidden)
    fwd = sum(e.values()) + sum(d.values())
    return fwd * (3.0 if backward else 1.0)
