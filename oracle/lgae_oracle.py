"""CPU oracle for the LGAE hot path -- TEST INFRASTRUCTURE ONLY.

A functional, CPU/float64 (torch) restatement of the algorithm the reference implements for
``LGNEncoder -> LGNDecoder -> chamfer loss``.  It is the checker the
CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
(``lgn_autoencoder_b200``) never does and fails loudly without its CUDA library.

Parity pinning: the reference holds no golden vectors, so this oracle is
pinned against outputs of the reference itself, imported in the build container from
``/root/reference`` by ``tests/golden/make_golden.py``; the vectors are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.

Every function cites the reference file:line whose arithmetic it follows.  The restatement
keeps the reference's *order of operations* where rounding matters (mass scalar, pair norms,
arg-min/arg-max selection) and its container iteration orders (channel order inside
concatenations), because both are observable in the outputs.

Reps are plain dicts ``{(k, n): tensor}``; a "vec" part has shape ``(2, *batch, C, (k+1)(n+1))``
and a "scalar" part ``(2, *batch, C)`` with index 0/1 of the leading axis = real/imaginary.
"""
from __future__ import annotations

import itertools
import math
from fractions import Fraction
from typing import Dict, List, Tuple

import numpy as np
import torch

Key = Tuple[int, int]
Rep = Dict[Key, torch.Tensor]

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------------------------
# Clebsch-Gordan coefficients  (lgn/cg_lib/cg_dict.py:190-281, 370-436)
# ----------------------------------------------------------------------------------------------
def _fact(x: Fraction) -> int:
    assert x.denominator == 1 and x >= 0, x
    return math.factorial(int(x))


def su2_cg(j1, m1, j2, m2, j3, m3) -> float:
    """Condon-Shortley <j1 m1 j2 m2 | j3 m3> by Racah's formula (cg_dict.py:370-436)."""
    j1, m1, j2, m2, j3, m3 = (Fraction(x).limit_denominator(2) for x in (j1, m1, j2, m2, j3, m3))
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    pref = Fraction(
        int(2 * j3 + 1) * _fact(j3 + j1 - j2) * _fact(j3 - j1 + j2) * _fact(j1 + j2 - j3)
        * _fact(j3 + m3) * _fact(j3 - m3),
        _fact(j1 + j2 + j3 + 1) * _fact(j1 - m1) * _fact(j1 + m1) * _fact(j2 - m2) * _fact(j2 + m2),
    )
    s = Fraction(0)
    for v in range(vmin, vmax + 1):
        sign = -1 if int(v + j2 + m2) % 2 else 1
        s += Fraction(
            sign * _fact(j2 + j3 + m1 - v) * _fact(j1 - m1 + v),
            _fact(Fraction(v)) * _fact(j3 - j1 + j2 - v) * _fact(j3 + m3 - v) * _fact(v + j1 - j2 - m3),
        )
    return math.sqrt(float(pref)) * float(s)


def _su2_mat(j1, j2, j3) -> np.ndarray:
    """Array [j1+m1, j2+m2, j3+m1+m2] of SU(2) CG coefficients (cg_dict.py:228-238)."""
    j1, j2, j3 = Fraction(j1).limit_denominator(2), Fraction(j2).limit_denominator(2), Fraction(j3).limit_denominator(2)
    mat = np.zeros((int(2 * j1 + 1), int(2 * j2 + 1), int(2 * j3 + 1)))
    if int(2 * j3) in range(int(2 * abs(j1 - j2)), int(2 * (j1 + j2)) + 1, 2):
        for a in range(int(2 * j1 + 1)):
            for b in range(int(2 * j2 + 1)):
                m1, m2 = a - j1, b - j2
                if abs(m1 + m2) <= j3:
                    mat[a, b, int(j3 + m1 + m2)] = su2_cg(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _recoupling(k: int, n: int) -> np.ndarray:
    """(k/2) x (n/2) -> sum_l l, multiplets stacked with l ascending (cg_dict.py:244-250)."""
    return np.concatenate(
        [_su2_mat(Fraction(k, 2), Fraction(n, 2), Fraction(i, 2)) for i in range(abs(k - n), k + n + 1, 2)], axis=-1
    )


def sl2c_cg(rep1: Key, rep2: Key, rep: Key) -> np.ndarray:
    """SL(2,C) CG block of shape (d1, d2, d_out) (cg_dict.py:241-281)."""
    (k1, n1), (k2, n2), (k, n) = rep1, rep2, rep
    out_rc = _recoupling(k, n)                                   # [a, b, c]
    left = _su2_mat(Fraction(k1, 2), Fraction(k2, 2), Fraction(k, 2))   # [d, e, a]
    right = _su2_mat(Fraction(n1, 2), Fraction(n2, 2), Fraction(n, 2))  # [g, h, b]
    in1 = _recoupling(k1, n1)                                    # [d, g, x]
    in2 = _recoupling(k2, n2)                                    # [e, h, y]
    return np.einsum("abc,dea,ghb,dgx,ehy->xyc", out_rc, left, right, in1, in2)


_CG_CACHE: Dict[int, dict] = {}


def cg_table(maxdim: int) -> dict:
    """{((k1,n1),(k2,n2)): {(k,n): tensor (d_out, d1*d2)}} for all k,n < maxdim, transposed as the
    reference stores it (cg_dict.py:96-114, 190-225)."""
    if maxdim in _CG_CACHE:
        return _CG_CACHE[maxdim]
    table = {}
    for k1, n1, k2, n2 in itertools.product(range(maxdim), repeat=4):
        entry = {}
        for k, n in itertools.product(range(abs(k1 - k2), k1 + k2 + 1, 2), range(abs(n1 - n2), n1 + n2 + 1, 2)):
            h = sl2c_cg((k1, n1), (k2, n2), (k, n))
            entry[(k, n)] = torch.from_numpy(h.reshape(-1, h.shape[-1]).T.copy())
        table[((k1, n1), (k2, n2))] = entry
    _CG_CACHE[maxdim] = table
    return table


# ----------------------------------------------------------------------------------------------
# Basis changes and Minkowski norms  (lgn/cg_lib/zonal_functions.py)
# ----------------------------------------------------------------------------------------------
def _cartesian4(dtype=torch.float64) -> torch.Tensor:
    r = 1.0 / SQRT2
    re = [[1, 0, 0, 0], [0, r, 0, 0], [0, 0, 0, 1], [0, -r, 0, 0]]
    im = [[0, 0, 0, 0], [0, 0, -r, 0], [0, 0, 0, 0], [0, 0, -r, 0]]
    return torch.tensor([re, im], dtype=dtype)


def normsq4(p: torch.Tensor) -> torch.Tensor:
    """Minkowski square of real Cartesian 4-vectors: 2*E^2 - sum(p^2) (zonal_functions.py:201-218)."""
    psq = torch.pow(p, 2)
    return 2 * psq[..., 0] - psq.sum(dim=-1)


def p_to_rep(p: torch.Tensor) -> torch.Tensor:
    """Real Cartesian (...,4) -> canonical (1,1) part (2,...,1,4) (zonal_functions.py:251-289)."""
    m = _cartesian4(p.dtype)
    pe = p.unsqueeze(-1)
    rep = torch.stack([torch.matmul(m[0], pe), torch.matmul(m[1], pe)], 0)
    return rep.squeeze(-1).unsqueeze(-2)


def p_cplx_to_rep(p: torch.Tensor) -> torch.Tensor:
    """Complex Cartesian (2,...,4) -> canonical (2,...,4) (zonal_functions.py:292-341)."""
    m = _cartesian4(p.dtype)
    pe = p.unsqueeze(-1)
    rep = torch.stack(
        (torch.matmul(m[0], pe[0]) - torch.matmul(m[1], pe[1]), torch.matmul(m[0], pe[1]) + torch.matmul(m[1], pe[0])), 0
    )
    return rep.squeeze(-1)


def rep_to_p(rep: torch.Tensor) -> torch.Tensor:
    """Canonical (2,...,4) -> complex Cartesian (2,...,4): Hermitian transpose of p_to_rep
    (zonal_functions.py:344-381)."""
    m = _cartesian4(rep.dtype)
    mh = torch.stack([m[0], -m[1]]).permute(0, 2, 1)
    r = rep.unsqueeze(-1)
    p = torch.stack((torch.matmul(mh[0], r[0]) - torch.matmul(mh[1], r[1]), torch.matmul(mh[0], r[1]) + torch.matmul(mh[1], r[0])), 0)
    return p.squeeze(-1)


def metric11(dtype=torch.float64) -> torch.Tensor:
    """Invariant bilinear form on (1,1) in the canonical basis (zonal_functions.py:396-411)."""
    g = torch.zeros(4, 4, dtype=dtype)
    g[0, 0] = 1.0
    g[1, 3] = g[3, 1] = 1.0
    g[2, 2] = -1.0
    return g


def repdot11(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Complex bilinear a.g.b on canonical 4-vectors, -> (2, ..., 1) (zonal_functions.py:414-438)."""
    g = metric11(a.dtype)
    e = lambda x, y: torch.einsum("...a,ab,...b->...", x, g, y)
    return torch.stack((e(a[0], b[0]) - e(a[1], b[1]), e(a[0], b[1]) + e(a[1], b[0])), 0).unsqueeze(-1)


EPS64 = 1e-16  # zonal_functions.py:441-446


def zonal_functions4(p: torch.Tensor, max_zf: int, cg: dict):
    """Real Cartesian input (zonal_functions.py:123-166). Returns (zf, norm, norm_sq)."""
    norm_sq = normsq4(p).unsqueeze(-1) + EPS64
    mask = norm_sq != 0
    norm = torch.where(mask, norm_sq / (norm_sq.abs().sqrt()), norm_sq)
    p11 = p_to_rep(p)
    zf = {(0, 0): torch.ones(p11.shape[:-1] + (1,), dtype=p11.dtype), (1, 1): p11}
    _higher_zonal(zf, p11, max_zf, cg)
    return zf, norm.squeeze(-1), norm_sq.squeeze(-1)


def zonal_functions_canonical(p: torch.Tensor, max_zf: int, cg: dict):
    """Complex canonical input (2,...,4) (zonal_functions.py:169-198)."""
    norm_sq = repdot11(p, p) + EPS64
    norm = norm_sq / (norm_sq.abs().sqrt())
    zf = {(0, 0): torch.ones(p.shape[:-1] + (1,), dtype=p.dtype), (1, 1): p}
    _higher_zonal(zf, p, max_zf, cg)
    zf = {key: val.unsqueeze(-2) for key, val in zf.items()}
    return zf, norm.squeeze(-1), norm_sq.squeeze(-1)


def _higher_zonal(zf: Rep, p11: torch.Tensor, max_zf: int, cg: dict) -> None:
    for l in range(2, max_zf + 1):  # zonal_functions.py:158-164
        new = cg_product(cg, {(l - 1, l - 1): zf[(l - 1, l - 1)]}, {(1, 1): p11}, maxdim=l + 1)[(l, l)]
        zf[(l, l)] = new * math.sqrt(2 * l / (l + 1))


def zonal_functions_rel(p1: torch.Tensor, p2: torch.Tensor, max_zf: int, cg: dict, basis: str):
    """Pairwise p_i - p_j (zonal_functions.py:221-248)."""
    rel = p1.unsqueeze(-2) - p2.unsqueeze(-3)
    if basis == "cartesian":
        return zonal_functions4(rel, max_zf, cg)
    return zonal_functions_canonical(rel, max_zf, cg)


# ----------------------------------------------------------------------------------------------
# Complex helpers  (lgn/g_lib/cplx_lib.py)
# ----------------------------------------------------------------------------------------------
def mix_zweight_zvec(w: torch.Tensor, part: torch.Tensor) -> torch.Tensor:
    """W (2,C',C) applied on the channel axis of part (2,...,C,d) (cplx_lib.py:7-25)."""
    return torch.stack([w[0] @ part[0] - w[1] @ part[1], w[1] @ part[0] + w[0] @ part[1]], 0)


def mul_zscalar_zirrep(scalar: torch.Tensor, part: torch.Tensor) -> torch.Tensor:
    """(2,...,C) x (2,...,C,d) (cplx_lib.py:54-72)."""
    s = scalar.unsqueeze(-1)
    return torch.stack([part[0] * s[0] - part[1] * s[1], part[0] * s[1] + part[1] * s[0]], 0)


# ----------------------------------------------------------------------------------------------
# CG product  (lgn/cg_lib/cg_ops.py:135-298)
# ----------------------------------------------------------------------------------------------
def complex_kron_product(z1: torch.Tensor, z2: torch.Tensor, aggregate: bool) -> torch.Tensor:
    """Channel-wise outer product of two complex parts, optionally summed over the neighbour
    axis (cg_ops.py:221-298)."""
    b1, b2 = z1.shape[1:-2], z2.shape[1:-2]
    c, d1, d2 = z1.shape[-2], z1.shape[-1], z2.shape[-1]
    if aggregate:
        if len(b1) == 3 and len(b2) == 2:
            z2 = z2.unsqueeze(2)
            b = b1
        elif len(b1) == 2 and len(b2) == 3:
            z1 = z1.unsqueeze(2)
            b = b2
        else:
            raise ValueError(f"Batch size error! {b1} {b2}")
        b1, b2 = z1.shape[1:-2], z2.shape[1:-2]
    else:
        assert b1 == b2
        b = b1
    z = z1.reshape((2, 1) + tuple(b1) + (c, d1, 1)) * z2.reshape((1, 2) + tuple(b2) + (c, 1, d2))
    z = z.contiguous().view((4,) + tuple(b) + (c, d1 * d2))
    if aggregate:
        z = z.sum(3)
    zrot = torch.tensor([[1.0, 0.0, 0.0, -1.0], [0.0, 1.0, 1.0, 0.0]], dtype=z.dtype)
    return torch.einsum("ab,b...->a...", zrot, z)


def cg_product(cg: dict, rep1: Rep, rep2: Rep, maxdim: int, aggregate: bool = False) -> Rep:
    """cg_ops.py:135-218. Output parts for the same irrep are concatenated on the channel axis in
    loop order (rep1 keys outer, rep2 keys inner)."""
    maxk1 = max(k for k, _ in rep1)
    maxn1 = max(n for _, n in rep1)
    maxk2 = max(k for k, _ in rep2)
    maxn2 = max(n for _, n in rep2)
    max_dim = min(max(maxk1 + maxk2, maxn1 + maxn2) + 1, maxdim)
    out: Dict[Key, List[torch.Tensor]] = {}
    for (k1, n1), part1 in rep1.items():
        for (k2, n2), part2 in rep2.items():
            if max(k1, n1, k2, n2) > max_dim - 1:
                continue
            keys = [
                (k, n)
                for k in range(abs(k1 - k2), min(maxdim, k1 + k2 + 1), 2)
                for n in range(abs(n1 - n2), min(maxdim, n1 + n2 + 1), 2)
            ]
            if not keys:
                continue
            cg_mat = torch.cat([cg[((k1, n1), (k2, n2))][key] for key in keys], -2)
            prod = complex_kron_product(part1, part2, aggregate)
            decomp = torch.matmul(cg_mat, prod.unsqueeze(-1)).squeeze(-1)
            pieces = torch.split(decomp, [(k + 1) * (n + 1) for k, n in keys], dim=-1)
            for key, piece in zip(keys, pieces):
                out.setdefault(key, []).append(piece)
    return {key: torch.cat(val, dim=-2) for key, val in out.items()}


def cg_product_tau(tau1: Dict[Key, int], tau2: Dict[Key, int], maxdim: int) -> Dict[Key, int]:
    """cg_ops_tau.py:6-44."""
    tau: Dict[Key, int] = {}
    for k1, n1 in tau1:
        for k2, n2 in tau2:
            if max(k1, n1, k2, n2) >= maxdim:
                continue
            for k in range(abs(k1 - k2), min(k1 + k2, maxdim - 1) + 1, 2):
                for n in range(abs(n1 - n2), min(n1 + n2, maxdim - 1) + 1, 2):
                    tau[(k, n)] = tau.get((k, n), 0) + tau1[(k1, n1)] * tau2[(k2, n2)]
    return tau


# ----------------------------------------------------------------------------------------------
# Mixing / concatenation with the reference's container iteration orders
# ----------------------------------------------------------------------------------------------
def weight_key_order(keys) -> List[Key]:
    """Order in which MixReps emits its output parts: ``ParameterDictNew.keys()`` is a *set* built
    from the registered names (g_lib/parameter_dict_new.py:14-15, g_torch.py:217-232)."""
    return list(set(map(eval, [str(k) for k in keys])))


def mix_reps(weights: Dict[Key, torch.Tensor], rep: Rep, key_order=None) -> Rep:
    """Per-irrep complex channel mix (nn/g_nn.py:95-117, g_torch.py:217-255)."""
    if key_order is None:
        key_order = weight_key_order(weights.keys())
    if set(rep.keys()) != set(weights.keys()):
        raise ValueError("Must have one mixing weight for each part of the rep")
    return {key: mix_zweight_zvec(weights[key], rep[key]) for key in key_order}


def cat_reps(reps: List[Rep], maxdim: int) -> Rep:
    """Truncate to max(key) < maxdim then concatenate on the channel axis; irreps are visited in
    the iteration order of a set union (nn/g_nn.py:169-189, g_torch.py:190-214)."""
    reps = [{key: val for key, val in rep.items() if max(key) < maxdim} for rep in reps]
    all_keys = set().union(*[rep.keys() for rep in reps])
    return {key: torch.cat([rep[key] for rep in reps if key in rep], dim=-2) for key in all_keys}


# ----------------------------------------------------------------------------------------------
# Radial functions  (lgn/nn/position_levels.py:118-209)
# ----------------------------------------------------------------------------------------------
def rad_poly_trig(norms, edge_mask, a, b, c, linears, num_channels: int, basis: str) -> Dict[Key, torch.Tensor]:
    """``linears`` = [(weight, bias)] for l = 0..max_zf. Cartesian basis: norms (B,N,N) ->
    parts (2,B,N,N,C) with output 2c -> re, 2c+1 -> im.  Canonical basis: norms (2,B,N,N), the same
    real Linear(2K -> C) applied to the re- and im-derived basis values."""
    s = tuple(norms.shape)
    mask = edge_mask.bool().unsqueeze(-1)
    x = norms.unsqueeze(-1)
    zero = torch.tensor(0, dtype=norms.dtype)
    trig = torch.where(mask, b * (torch.ones_like(b) + (c * x).pow(2) + 1e-16).pow(-1) + a, zero)
    trig = trig.view(s + (1, trig.shape[-1]))
    out = {}
    for l, (w, bias) in enumerate(linears):
        y = torch.nn.functional.linear(trig, w, bias)
        if basis == "canonical":
            assert len(s) == 4
            out[(l, l)] = y.view(s + (num_channels,))
        else:
            assert len(s) == 3
            out[(l, l)] = y.view(s + (num_channels, 2)).permute(4, 0, 1, 2, 3)
    return out


# ----------------------------------------------------------------------------------------------
# Levels  (lgn/models/lgn_levels.py, lgn/models/lgn_cg.py)
# ----------------------------------------------------------------------------------------------
def cgmlp(rep: Rep, linears, negative_slope: float = 0.01) -> Rep:
    """MLP on the (0,0) part, re/im interleaved on the feature axis; pops (0,0) and re-inserts it
    at the end of the dict (lgn_levels.py:191-227)."""
    x = rep.pop((0, 0)).squeeze(-1)
    s = x.shape
    x = x.permute(1, 2, 3, 0).contiguous().view(s[1:3] + (2 * s[3],))
    for w, b in linears[:-1]:
        x = torch.nn.functional.leaky_relu(torch.nn.functional.linear(x, w, b), negative_slope)
    w, b = linears[-1]
    x = torch.nn.functional.linear(x, w, b)
    rep[(0, 0)] = x.view(s[1:] + (2,)).permute(3, 0, 1, 2).unsqueeze(-1)
    return rep


def node_level(cg, node: Rep, edge: Rep, mix_w: Dict[Key, torch.Tensor], maxdim: int) -> Rep:
    """lgn_levels.py:96-121: aggregate over neighbours, self product, concat, mix."""
    ag = cg_product(cg, node, edge, maxdim=maxdim, aggregate=True)
    sq = cg_product(cg, node, node, maxdim=maxdim, aggregate=False)
    cat = cat_reps([ag, node, sq], maxdim)
    return mix_reps(mix_w, cat)


# ----------------------------------------------------------------------------------------------
# State-dict access helpers (key names: SURVEY.md appendix A.9)
# ----------------------------------------------------------------------------------------------
def _weights(sd: dict, prefix: str) -> Dict[Key, torch.Tensor]:
    """Collect ``prefix.weights.(k, n)`` entries in state-dict (= registration) order."""
    out = {}
    pre = prefix + ".weights."
    for name, val in sd.items():
        if name.startswith(pre):
            out[eval(name[len(pre):])] = val
    return out


def _linears(sd: dict, prefix: str):
    out = []
    i = 0
    while f"{prefix}.{i}.weight" in sd:
        out.append((sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"]))
        i += 1
    return out


def _num_levels(sd: dict) -> int:
    n = 0
    while f"lgn_cg.node_levels.{n}.cat_mix.mix_reps.weights.(0, 0)" in sd:
        n += 1
    return n


def _lgn_cg(sd: dict, cg, node: Rep, rad_levels, zonal: Rep, maxdim: List[int], nodes_all: List[Rep]) -> Rep:
    """lgn_cg.py:124-180 with mlp=True."""
    for lvl in range(len(rad_levels)):
        edge = {key: mul_zscalar_zirrep(rad_levels[lvl][key], zonal[key]) for key in rad_levels[lvl]}
        node = node_level(cg, node, edge, _weights(sd, f"lgn_cg.node_levels.{lvl}.cat_mix.mix_reps"), maxdim[lvl])
        if f"lgn_cg.mlp_levels.{lvl}.linear.0.weight" in sd:
            node = cgmlp(node, _linears(sd, f"lgn_cg.mlp_levels.{lvl}.linear"))
        nodes_all.append(dict(node))
    return node


def _radial_levels(sd: dict, norms, mask, num_channels: List[int], basis: str):
    levels = []
    lvl = 0
    while f"rad_funcs.rad_funcs.{lvl}.a" in sd:
        pre = f"rad_funcs.rad_funcs.{lvl}"
        levels.append(
            rad_poly_trig(norms, mask, sd[pre + ".a"], sd[pre + ".b"], sd[pre + ".c"], _linears(sd, pre + ".linear"),
                          num_channels[lvl], basis)
        )
        lvl += 1
    return levels


def _adapt(var, n):
    if isinstance(var, (int, float)):
        return [var] * n
    var = list(var)
    if len(var) < n:
        return var + [var[-1]] * (n - len(var))
    return var[:n] if len(var) == n else var[: n - 1]


# ----------------------------------------------------------------------------------------------
# Latent aggregation  (lgn/models/lgn_encoder.py:419-583)
# ----------------------------------------------------------------------------------------------
def get_msq(p4: torch.Tensor) -> torch.Tensor:
    e, p3 = p4[..., 0], p4[..., 1:]
    return e ** 2 - torch.norm(p3, dim=-1) ** 2  # lgn_encoder.py:499-505 (sqrt then square)


def _gather_particles(feature: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    # feature (2,B,N,tau,d); idx (2,B,tau) -> (2,B,1,tau,d)   (lgn_encoder.py:508-537)
    d = feature.shape[-1]
    index = idx.unsqueeze(2).unsqueeze(-1).expand(-1, -1, 1, -1, d)
    return torch.gather(feature, 2, index)


def get_min_features(feature: torch.Tensor) -> torch.Tensor:
    if feature.shape[-1] == 1:
        scalar = feature.min(dim=-1).values
    elif feature.shape[-1] == 4:
        scalar = get_msq(feature)
    else:
        raise NotImplementedError
    return _gather_particles(feature, torch.min(scalar, dim=-2).indices)


def get_max_features(feature: torch.Tensor) -> torch.Tensor:
    scalar = get_msq(feature)  # for d == 1 this is s**2 (lgn_encoder.py:568-569)
    if feature.shape[-1] not in (1, 4):
        raise NotImplementedError
    return _gather_particles(feature, torch.max(scalar, dim=-2).indices)


def aggregate(method: str, latent: Rep) -> Rep:
    m = method.lower()
    if m == "sum":
        return {k: torch.sum(v, dim=-3, keepdim=True).unsqueeze(dim=-3) for k, v in latent.items()}
    if m in ("mean", "average"):
        return {k: torch.mean(v, dim=-3, keepdim=True) for k, v in latent.items()}
    if m == "max":
        return {k: get_max_features(v) for k, v in latent.items()}
    if m == "min":
        return {k: get_min_features(v) for k, v in latent.items()}
    if m == "mix":
        return latent
    if "+" in m:
        parts = [aggregate(x, latent) for x in method.split("+")]
        return {k: sum(p[k] for p in parts) / len(parts) for k in latent}
    if "&" in method:
        parts = [aggregate(x, latent) for x in method.split("&")]
        return {k: torch.cat([p[k] for p in parts], dim=3) for k in latent}
    raise NotImplementedError(method)


# ----------------------------------------------------------------------------------------------
# Encoder / decoder forward  (lgn/models/lgn_encoder.py:255-412, lgn/models/lgn_decoder.py:218-345)
# ----------------------------------------------------------------------------------------------
def encoder_forward(sd: dict, cfg: dict, data, covariance_test: bool = False):
    """``cfg``: num_channels (list), maxdim, max_zf, map_to_latent.  ``data``: dict with 'p4'
    (B,N,4) and optionally 'labels', or a bare tensor."""
    if not isinstance(data, dict):
        data = {"p4": torch.as_tensor(data)}
    num_channels = list(cfg["num_channels"])
    n_lvl = len(num_channels) - 1
    maxdim = _adapt(cfg["maxdim"], n_lvl)
    max_zf = _adapt(cfg.get("max_zf", [1]), n_lvl)
    cg = cg_table(max(maxdim + max_zf))
    p = data["p4"].to(torch.float64) * cfg.get("scale", 1.0)
    scalars = normsq4(p).abs().sqrt().unsqueeze(-1)                       # lgn_encoder.py:376
    for key in ("labels", "masks", "mask"):                                # lgn_encoder.py:387-395
        if key in data:
            node_mask = data[key].to(torch.uint8)
            break
    else:
        node_mask = (data["p4"][..., 0] != 0).to(torch.uint8)
    edge_mask = node_mask.unsqueeze(1) * node_mask.unsqueeze(2)           # :403

    zf_in, _, _ = zonal_functions4(p, max(max_zf), cg)
    zf_in[(0, 0)] = torch.stack([scalars.unsqueeze(-1), torch.zeros_like(scalars.unsqueeze(-1))])  # :290-292
    zonal, norms, _ = zonal_functions_rel(p, p, max(max_zf), cg, "cartesian")
    rad = _radial_levels(sd, norms, edge_mask * (norms != 0).byte(), num_channels, "cartesian")
    node = mix_reps(_weights(sd, "input_func_node"), zf_in)
    nodes_all = [dict(node)]
    node = _lgn_cg(sd, cg, node, rad, zonal, maxdim, nodes_all)

    if cfg["map_to_latent"].lower() == "mix":
        node = {k: v.reshape(2, v.shape[1], 1, -1, v.shape[-1]) for k, v in node.items()}   # :313-319
    latent = mix_reps(_weights(sd, "mix_reps"), node)
    latent = {k: latent[k] for k in [(0, 0), (1, 1)]}
    latent[(1, 1)] = rep_to_p(latent[(1, 1)])
    latent = aggregate(cfg["map_to_latent"], latent)
    return (latent, nodes_all) if covariance_test else latent


def decoder_forward(sd: dict, cfg: dict, latent: Rep, covariance_test: bool = False, nodes_all=None):
    num_channels = list(cfg["num_channels"])
    n_lvl = len(num_channels) - 1
    maxdim = _adapt(cfg["maxdim"], n_lvl)
    max_zf = _adapt(cfg.get("max_zf", [1]), n_lvl)
    cg = cg_table(max(maxdim + max_zf))

    graph = mix_reps(_weights(sd, "latent_to_graph"), latent)
    graph = {k: v.squeeze(-3) for k, v in graph.items()}                   # lgn_decoder.py:327-330
    p = p_cplx_to_rep(graph[(1, 1)])                                       # (2,B,N,4) canonical
    b, n = p.shape[1], p.shape[2]
    edge_mask = torch.zeros(2, b, n, n, dtype=torch.float32)               # :335-340 (all zeros)

    zf_in, _, _ = zonal_functions_canonical(p, max(max_zf), cg)
    zonal, norms, _ = zonal_functions_rel(p, p, max(max_zf), cg, "canonical")
    rad = _radial_levels(sd, norms, edge_mask * (norms != 0).byte(), num_channels, "canonical")
    node = mix_reps(_weights(sd, "input_func_node"), zf_in)
    dec_nodes = [dict(node)]
    node = _lgn_cg(sd, cg, node, rad, zonal, maxdim, dec_nodes)
    gen = mix_reps(_weights(sd, "mix_to_output"), node)
    gen = {k: gen[k] for k in [(0, 0), (1, 1)]}
    ps = rep_to_p(gen[(1, 1)].clone()).squeeze(-2)
    if not covariance_test:
        return ps
    nodes_all = list(nodes_all) + dec_nodes + [gen]                        # :298-303 (input mix + levels + generated)
    return gen, nodes_all


# ----------------------------------------------------------------------------------------------
# Caller-side ops on the hot path (utils/normalize_p4.py, utils/utils.py, chamfer_loss)
# ----------------------------------------------------------------------------------------------
def normalize_p4_overall_max(p4: torch.Tensor):
    f = torch.abs(p4).amax(dim=-1, keepdim=True).amax(dim=-2, keepdim=True) + 1e-16   # normalize_p4.py:39-52
    return p4 / f, f


def get_real_sum(x: torch.Tensor) -> torch.Tensor:
    return x[0] + x[1]                                                      # utils/utils.py:201-202


def get_real(x: torch.Tensor, method: str = "real", eps: float = 1e-16) -> torch.Tensor:
    """utils/utils.py:194-207 (an unknown method means 'real')."""
    m = method.lower()
    if m == "imag":
        return x[1]
    if m == "norm":
        return torch.sqrt(torch.pow(x[0], 2) + torch.pow(x[1], 2) + eps)
    if m == "sum":
        return x[0] + x[1]
    if m == "mean":
        return (x[0] + x[1]) / 2
    return x[0]


def chamfer_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """chamfer_loss.py:16-31 over distance_sq.py:263-304 (p=2): squared Euclidean distance on all
    four components, **summed** over the batch."""
    diffs = -(x.unsqueeze(-2) - y.unsqueeze(-3))
    dist = torch.sum(diffs ** 2, dim=-1)
    return torch.sum((dist.min(dim=-1).values + dist.min(dim=-2).values) / 2)


def anomaly_scores_cartesian(recons: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The Cartesian-family scores of utils/jet_analysis/anomaly_detection.py:251-419 on real (B,N,4) jets: chamfer (:498-503,
    Euclidean norm, both directions added per particle index, mean over particles), mse (:473), their Minkowski versions
    (:523-527, :428-437) and the jet-level scores (:405, :419; jet = sum over particles, :653-654)."""
    def lorentz(x):
        return x[..., 0] ** 2 - x[..., 1] ** 2 - x[..., 2] ** 2 - x[..., 3] ** 2
    diffs = recons.unsqueeze(-2) - target.unsqueeze(-3)
    dist = torch.norm(diffs, dim=-1)
    dl = lorentz(diffs)
    jr, jt = recons.sum(-2), target.sum(-2)
    return {
        "chamfer_particle_cartesian": (dist.min(-1).values + dist.min(-2).values).mean(-1),
        "mse_particle_cartesian": ((recons - target) ** 2).sum(-1).mean(-1),
        "chamfer_particle_lorentz": (dl.min(-1).values + dl.min(-2).values).mean(-1),
        "mse_particle_lorentz": lorentz(recons - target).mean(-1),
        "jet_cartesian": ((jr - jt) ** 2).sum(-1),
        "jet_lorentz": lorentz(jr - jt),
    }


def l1_norm(sd: dict) -> torch.Tensor:
    return sum(p.abs().sum() for p in sd.values())                         # lgn_encoder.py:249-250


def training_step(enc_sd, dec_sd, enc_cfg, dec_cfg, data, l1_lambda: float = 1e-8, get_real_method: str = "sum"):
    """utils/train.py:283-327 minus optimizer: returns (loss, latent, recons)."""
    latent = encoder_forward(enc_sd, enc_cfg, data)
    recons_c = decoder_forward(dec_sd, dec_cfg, latent)
    recons = get_real(recons_c, get_real_method)
    p4 = data["p4"] if isinstance(data, dict) else data
    loss = chamfer_loss(recons, p4.to(torch.float64))
    if l1_lambda:
        loss = loss + l1_lambda * (l1_norm(enc_sd) + l1_norm(dec_sd))
    return loss, latent, recons_c


# ----------------------------------------------------------------------------------------------
# Synthetic jets -- shared by tests and bench
# ----------------------------------------------------------------------------------------------
def synthetic_jets(batch: int, n: int, seed: int = 0, mass_scale: float = 1e-6, pad: bool = False):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
    pt = 0.2 * u(batch, n) ** 3 + 1e-3
    eta = 0.8 * u(batch, n) - 0.4
    phi = 0.8 * u(batch, n) - 0.4
    m = mass_scale * u(batch, n)
    px, py, pz = pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta)
    e = torch.sqrt((pt * torch.cosh(eta)) ** 2 + m ** 2)
    p4 = torch.stack([e, px, py, pz], -1)
    out = {"p4": p4}
    if pad:
        nobj = torch.randint(max(1, n // 3), n + 1, (batch,), generator=g)
        labels = (torch.arange(n).unsqueeze(0) < nobj.unsqueeze(1))
        p4 = p4 * labels.unsqueeze(-1)
        out = {"p4": p4, "labels": labels.to(torch.float64), "Nobj": nobj}
    return out
