"""Layer-level kernels (csrc/lgae_cg.cu) behind CGProduct / cg_product / MixReps: term tables on CPU, and on the GPU
the product and its adjoint against the oracle's restatement of the reference (cg_ops.py:135-298, cplx_lib.py:7-25)."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_err

TOL = 1e-12   # fp64, sums of <= a few hundred products


def _cg(maxdim):
    from lgn_autoencoder_b200.cg_lib import CGDict
    return CGDict(maxdim=maxdim, transpose=True, dtype=torch.float64, device=torch.device("cpu"))


@pytest.mark.parametrize("swap", [False, True])
def test_term_tables_reproduce_the_cg_matrices(swap):
    from lgn_autoencoder_b200 import layer_ops
    cg = _cg(3)
    irreps = [(0, 0), (1, 1), (0, 2), (2, 0), (2, 2)]
    for key1 in irreps:
        for key2 in irreps:
            outs = [(k, n) for k in range(abs(key1[0] - key2[0]), min(3, key1[0] + key2[0] + 1), 2)
                    for n in range(abs(key1[1] - key2[1]), min(3, key1[1] + key2[1] + 1), 2)]
            tab, coef, n, n_comp, dk1, dk2 = layer_ops._term_tables(cg, key1, key2, outs, swap, torch.device("cpu"))
            d1, d2 = (key1[0] + 1) * (key1[1] + 1), (key2[0] + 1) * (key2[1] + 1)
            assert (dk1, dk2) == ((d2, d1) if swap else (d1, d2))
            dense = torch.cat([cg[(key1, key2)][o] for o in outs], 0).numpy()
            assert n == np.count_nonzero(dense) and n_comp == dense.shape[0]
            tab, coef = tab.numpy(), coef.numpy().reshape(3, n)
            trip = tab[:9 * n].reshape(3, n, 3)
            starts = np.split(tab[9 * n:], [n_comp + 1, n_comp + 1 + dk1 + 1])
            assert [len(s) for s in starts] == [n_comp + 1, dk1 + 1, dk2 + 1]
            for order in range(3):
                rebuilt = np.zeros_like(dense)
                for (comp, a, d), c in zip(trip[order], coef[order]):
                    if swap:
                        a, d = d, a
                    rebuilt[comp, a * d2 + d] += c
                assert np.array_equal(rebuilt, dense)
                col = trip[order][:, order]
                assert np.all(np.diff(col) >= 0)
                s = starts[order]
                assert s[0] == 0 and s[-1] == n
                for g in range(len(s) - 1):
                    assert np.all(col[s[g]:s[g + 1]] == g)


def _rand_rep(gen, batch, tau, requires_grad=True):
    rep = {k: torch.randn((2,) + batch + (c, (k[0] + 1) * (k[1] + 1)), generator=gen, dtype=torch.float64) for k, c in tau.items()}
    return {k: v.requires_grad_(requires_grad) for k, v in rep.items()}


def _run_both(rep1, rep2, maxdim, aggregate):
    """(outputs, input grads) of the oracle on CPU and of the kernels on cuda:0 for the same random cotangent."""
    from lgn_autoencoder_b200.cg_lib import cg_product
    from lgn_autoencoder_b200.g_lib import GVec
    from oracle import lgae_oracle as orc
    dev = torch.device("cuda:0")
    cg_o = orc.cg_table(maxdim)
    ref = orc.cg_product(cg_o, rep1, rep2, maxdim, aggregate)
    gen = torch.Generator().manual_seed(7)
    cot = {k: torch.randn(v.shape, generator=gen, dtype=torch.float64) for k, v in ref.items()}
    leaves = list(rep1.values()) + [v for v in rep2.values() if all(v is not u for u in rep1.values())]
    g_ref = torch.autograd.grad(sum((ref[k] * cot[k]).sum() for k in ref), leaves, allow_unused=True)
    d1 = {k: v.detach().to(dev).requires_grad_(True) for k, v in rep1.items()}
    d2 = d1 if rep2 is rep1 else {k: v.detach().to(dev).requires_grad_(True) for k, v in rep2.items()}
    out = cg_product(_cg(maxdim).to(device=dev), GVec(d1, ignore_check=True), GVec(d2, ignore_check=True), maxdim=maxdim, aggregate=aggregate)
    dleaves = list(d1.values()) + ([] if d2 is d1 else list(d2.values()))
    g_out = torch.autograd.grad(sum((out[k] * cot[k].to(dev)).sum() for k in out.keys()), dleaves, allow_unused=True)
    return ref, out, g_ref, g_out


def _check(ref, out, g_ref, g_out):
    assert list(out.keys()) == list(ref.keys())          # part order is part of the contract
    for k in ref:
        assert rel_err(out[k], ref[k]) < TOL, k
    for a, b in zip(g_out, g_ref):
        if b is None:
            assert a is None or a.abs().max().item() == 0.0
        else:
            assert rel_err(a, b) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("maxdim,C,B,N", [(2, 3, 2, 5), (3, 3, 2, 5), (3, 8, 2, 40), (3, 4, 1, 150)])   # N = 150: neighbour-tiled kernels
def test_cg_aggregate_matches_oracle(maxdim, C, B, N):
    gen = torch.Generator().manual_seed(maxdim * 100 + C)
    irreps = [(0, 0), (1, 1)] + ([(0, 2), (2, 0), (2, 2)] if maxdim == 3 else [])
    node = _rand_rep(gen, (B, N), {k: C for k in irreps})
    edge = _rand_rep(gen, (B, N, N), {(0, 0): C, (1, 1): C})
    _check(*_run_both(node, edge, maxdim, True))
    _check(*_run_both(edge, node, maxdim, True))        # the other orientation (edge (x) node)


@pytest.mark.gpu
@pytest.mark.parametrize("maxdim", [2, 3])
def test_cg_pointwise_and_power_match_oracle(maxdim):
    gen = torch.Generator().manual_seed(maxdim)
    irreps = [(0, 0), (1, 1)] + ([(0, 2), (2, 0), (2, 2)] if maxdim == 3 else [])
    a = _rand_rep(gen, (3, 7), {k: 4 for k in irreps})
    b = _rand_rep(gen, (3, 7), {k: 4 for k in irreps[:3]})
    _check(*_run_both(a, b, maxdim, False))
    _check(*_run_both(a, a, maxdim, False))              # cg_power: both operands are the same tensors


@pytest.mark.gpu
def test_cg_product_rejects_cpu_tensors_and_bad_shapes():
    from lgn_autoencoder_b200.cg_lib import cg_product
    from lgn_autoencoder_b200.g_lib import GVec
    gen = torch.Generator().manual_seed(0)
    a = GVec(_rand_rep(gen, (2, 4), {(0, 0): 2, (1, 1): 2}, False), ignore_check=True)
    with pytest.raises(RuntimeError):
        cg_product(_cg(2), a, a, maxdim=2)
    dev = torch.device("cuda:0")
    a3 = GVec({k: v.to(dev) for k, v in _rand_rep(gen, (2, 4), {(0, 0): 2, (1, 1): 2}, False).items()}, ignore_check=True)
    b3 = GVec({k: v.to(dev) for k, v in _rand_rep(gen, (2, 4), {(0, 0): 3, (1, 1): 3}, False).items()}, ignore_check=True)
    with pytest.raises(ValueError):
        cg_product(_cg(2).to(device=dev), a3, b3, maxdim=2)     # channel mismatch


@pytest.mark.gpu
@pytest.mark.parametrize("shape,cin,cout,d", [((2, 5), 15, 4, 4), ((3,), 7, 9, 1), ((2, 3, 3), 40, 8, 9), ((700,), 20, 4, 4),
                                                ((5,), 6, 20, 3), ((4,), 300, 3, 1), ((1,), 1, 1, 1)])
def test_mix_matches_oracle(shape, cin, cout, d):
    from lgn_autoencoder_b200 import layer_ops
    from oracle import lgae_oracle as orc
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(cin)
    w = torch.randn((2, cout, cin), generator=gen, dtype=torch.float64, requires_grad=True)
    x = torch.randn((2,) + shape + (cin, d), generator=gen, dtype=torch.float64, requires_grad=True)
    ref = orc.mix_zweight_zvec(w, x)
    cot = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    gw_ref, gx_ref = torch.autograd.grad((ref * cot).sum(), (w, x))
    wd, xd = w.detach().to(dev).requires_grad_(True), x.detach().to(dev).requires_grad_(True)
    out = layer_ops.mix(wd, xd)
    gw, gx = torch.autograd.grad((out * cot.to(dev)).sum(), (wd, xd))
    assert rel_err(out, ref) < TOL and rel_err(gx, gx_ref) < TOL and rel_err(gw, gw_ref) < TOL
    # reproducible: fixed summation order
    gw2, _ = torch.autograd.grad((layer_ops.mix(wd, xd) * cot.to(dev)).sum(), (wd, xd))
    assert torch.equal(gw, gw2)


@pytest.mark.gpu
@pytest.mark.parametrize("cs,cv", [(5, 1), (5, 5), (1, 5)])
def test_scalar_times_irrep_matches_oracle(cs, cv):
    from lgn_autoencoder_b200 import layer_ops
    from oracle import lgae_oracle as orc
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(cs * 10 + cv)
    s = torch.randn((2, 2, 6, 6, cs), generator=gen, dtype=torch.float64, requires_grad=True)
    v = torch.randn((2, 2, 6, 6, cv, 4), generator=gen, dtype=torch.float64, requires_grad=True)
    ref = orc.mul_zscalar_zirrep(s, v)
    cot = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    gs_ref, gv_ref = torch.autograd.grad((ref * cot).sum(), (s, v))
    sd, vd = s.detach().to(dev).requires_grad_(True), v.detach().to(dev).requires_grad_(True)
    out = layer_ops.scalar_irrep(sd, vd)
    gs, gv = torch.autograd.grad((out * cot.to(dev)).sum(), (sd, vd))
    assert rel_err(out, ref) < TOL and rel_err(gs, gs_ref) < TOL and rel_err(gv, gv_ref) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("basis,B,N,C", [("cartesian", 2, 7, 3), ("canonical", 2, 7, 4), ("cartesian", 3, 40, 8)])
def test_radial_functions_match_oracle(basis, B, N, C):
    """RadPolyTrig module (kernel) vs the oracle's restatement: outputs, d/dnorms and every parameter gradient."""
    from lgn_autoencoder_b200.nn import RadPolyTrig
    from oracle import lgae_oracle as orc
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(N)
    torch.manual_seed(5)
    mod = RadPolyTrig(1, 10, C, mix=True, input_basis=basis, device=dev, dtype=torch.float64)
    shape = (B, N, N) if basis == "cartesian" else (2, B, N, N)
    norms = torch.randn(shape, generator=gen, dtype=torch.float64)
    mask = (torch.rand((B, N, N), generator=gen) > 0.3).to(torch.uint8)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.named_parameters()}
    x_ref = norms.clone().requires_grad_(True)
    lin = [(params[f"linear.{l}.weight"], params[f"linear.{l}.bias"]) for l in range(2)]
    ref = orc.rad_poly_trig(x_ref, mask, params["a"], params["b"], params["c"], lin, C, basis)
    cot = {k: torch.randn(v.shape, generator=gen, dtype=torch.float64) for k, v in ref.items()}
    names = list(params)
    g_ref = torch.autograd.grad(sum((ref[k] * cot[k]).sum() for k in ref), [x_ref] + [params[n] for n in names])
    x = norms.to(dev).requires_grad_(True)
    out = mod(x, mask.to(dev))
    assert list(out.keys()) == list(ref.keys())
    for k in ref:
        assert out[k].shape == ref[k].shape and rel_err(out[k], ref[k]) < TOL
    g = torch.autograd.grad(sum((out[k] * cot[k].to(dev)).sum() for k in ref), [x] + [dict(mod.named_parameters())[n] for n in names])
    for a, b, n in zip(g, g_ref, ["norms"] + names):
        assert rel_err(a, b) < 1e-11, n


@pytest.mark.gpu
@pytest.mark.parametrize("rows,nin,nout,slope", [(37, 6, 36, 0.01), (700, 96, 96, None), (1000, 72, 12, 0.01), (5, 3, 130, 0.2),
                                                 (1003, 128, 128, 0.01), (19, 5, 7, None), (4100, 16, 96, 0.01), (333, 97, 33, 0.3), (1, 1, 1, None), (9, 2, 129, 0.01)])
def test_linear_matches_torch_fp64(rows, nin, nout, slope):
    """Floating-point GEMM kernel: reference = the same op in plain torch fp64 on the CPU; tolerance 1e-12 relative."""
    from lgn_autoencoder_b200 import layer_ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(rows)
    x = torch.randn((rows, nin), generator=gen, dtype=torch.float64, requires_grad=True)
    w = torch.randn((nout, nin), generator=gen, dtype=torch.float64, requires_grad=True)
    b = torch.randn((nout,), generator=gen, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.linear(x, w, b)
    if slope is not None:
        ref = torch.nn.functional.leaky_relu(ref, slope)
    cot = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    g_ref = torch.autograd.grad((ref * cot).sum(), (x, w, b))
    xd, wd, bd = (t.detach().to(dev).requires_grad_(True) for t in (x, w, b))
    out = layer_ops.linear(xd, wd, bd, leaky_slope=slope)
    g = torch.autograd.grad((out * cot.to(dev)).sum(), (xd, wd, bd))
    assert rel_err(out, ref) < TOL
    for a, r in zip(g, g_ref):
        assert rel_err(a, r) < TOL


@pytest.mark.parametrize("swap", [False, True])
def test_multi_pair_plan_reproduces_every_pair(swap):
    """The one-launch plan of an aggregated call (all node x edge irrep pairs): its global term list, component starts and
    output map give, for random operands, exactly what the dense CG matrices of every pair give, in the reference's channel order."""
    from lgn_autoencoder_b200 import layer_ops
    cg = _cg(3)
    node_keys, edge_keys = [(0, 0), (1, 1), (0, 2), (2, 0), (2, 2)], [(0, 0), (1, 1)]
    keys1, keys2 = (edge_keys, node_keys) if swap else (node_keys, edge_keys)
    C = 3
    pairs, out_keys, out_ch = layer_ops.plan_pairs(cg, keys1, keys2, [C] * len(keys1), [C] * len(keys2), 3, swap, torch.device("cpu"))
    dims1, dims2 = [(k + 1) * (n + 1) for k, n in keys1], [(k + 1) * (n + 1) for k, n in keys2]
    mp = layer_ops.MultiPlan(cg, pairs, dims1, dims2, C, out_keys, out_ch, swap, torch.device("cpu"))
    nt, nc = mp.desc.n_terms, mp.desc.n_comp
    tab = mp.tab.numpy()
    rows, start, cinfo = tab[:3 * nt].reshape(nt, 3), tab[3 * nt:3 * nt + nc + 1], tab[3 * nt + nc + 1:].reshape(nc, 3)
    assert np.all(np.diff(rows[:, 0]) >= 0) and start[0] == 0 and start[-1] == nt
    rng = np.random.default_rng(0)
    node_d, edge_d = (dims2, dims1) if swap else (dims1, dims2)
    x, y = rng.normal(size=sum(node_d)), rng.normal(size=sum(edge_d))
    noff, eoff = np.concatenate(([0], np.cumsum(node_d))), np.concatenate(([0], np.cumsum(edge_d)))
    got = {}
    for oc in range(nc):
        got[(int(cinfo[oc, 0]), int(cinfo[oc, 2]), int(cinfo[oc, 1]))] = sum(
            mp.coef[t].item() * x[rows[t, 1]] * y[rows[t, 2]] for t in range(start[oc], start[oc + 1]))
    ref, chan = {}, {k: 0 for k in out_keys}
    for i1, key1 in enumerate(keys1):
        for i2, key2 in enumerate(keys2):
            outs = [(k, n) for k in range(abs(key1[0] - key2[0]), min(3, key1[0] + key2[0] + 1), 2)
                    for n in range(abs(key1[1] - key2[1]), min(3, key1[1] + key2[1] + 1), 2)]
            z1, z2 = (y[eoff[i1]:eoff[i1 + 1]], x[noff[i2]:noff[i2 + 1]]) if swap else (x[noff[i1]:noff[i1 + 1]], y[eoff[i2]:eoff[i2 + 1]])
            kron = np.outer(z1, z2).reshape(-1)
            for ok in outs:
                for m, v in enumerate(cg[(key1, key2)][ok].numpy() @ kron):
                    ref[(out_keys.index(ok), chan[ok], m)] = v
                chan[ok] += C
    assert set(got) == set(ref)
    assert max(abs(got[k] - ref[k]) for k in ref) < 1e-14
    # adjoint order: the same terms sorted by cell (a_all, d_all) apply the transposed CG matrices
    ncell = mp.d1t * mp.d2t
    tab_b = mp.tab_b.numpy()
    oc_b, cstart = tab_b[:nt], tab_b[nt:nt + ncell + 1]
    assert cstart[0] == 0 and cstart[-1] == nt and np.array_equal(tab_b[nt + ncell + 1:].reshape(nc, 3), cinfo)
    g = rng.normal(size=nc)
    gk = np.array([sum(mp.coef_b[t].item() * g[oc_b[t]] for t in range(cstart[c], cstart[c + 1])) for c in range(ncell)])
    gk_ref = np.zeros(ncell)
    for t in range(nt):
        gk_ref[rows[t, 1] * mp.d2t + rows[t, 2]] += mp.coef[t].item() * g[rows[t, 0]]
    assert np.allclose(gk, gk_ref, atol=1e-15)
