"""Pin the CPU oracle (oracle/lgae_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import lgae_oracle as orc
from tests.helpers import dec_cfg, enc_cfg, load_golden, rel_err

TOL = 1e-12  # fp64, identical operation order up to BLAS/summation details
CASES = ["cfg1_b3", "pad_n8", "mean_n6", "md3_mix_n5", "n40_b2", "cfg4_b2", "minplusmax_n6", "real_n6", "norm_n6"]


def test_cg_coefficients_match_reference():
    ref = load_golden("cg_maxdim3")
    mine = orc.cg_table(3)
    assert {str(k) for k in mine} == set(ref)
    for k, entry in mine.items():
        assert {str(kk) for kk in entry} == set(ref[str(k)])
        for kk, mat in entry.items():
            assert torch.allclose(mat, ref[str(k)][str(kk)], atol=1e-14, rtol=0), (k, kk)


def test_cg_stacked_matrix_is_orthogonal():
    table = orc.cg_table(3)
    for (r1, r2), entry in table.items():
        h = torch.cat(list(entry.values()), 0)
        assert torch.allclose(h @ h.T, torch.eye(h.shape[0], dtype=h.dtype), atol=1e-13)


def test_basis_kat():
    kat = load_golden("basis_kat")
    rep = orc.p_to_rep(kat["p"])
    assert torch.equal(rep, kat["rep"])
    assert torch.allclose(orc.rep_to_p(rep.squeeze(-2)), kat["back"], atol=1e-15)
    assert torch.equal(orc.normsq4(kat["p"]), kat["normsq4"])
    # (1,1)x(1,1)->(0,0) row is half the Minkowski metric in the canonical basis
    h = orc.cg_table(2)[((1, 1), (1, 1))][(0, 0)].view(4, 4)
    assert torch.allclose(h, 0.5 * orc.metric11(), atol=1e-15)


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    with torch.no_grad():
        latent, nodes = orc.encoder_forward(g["enc_state"], enc_cfg(cfg), g["batch"], covariance_test=True)
        gen, nodes_all = orc.decoder_forward(g["dec_state"], dec_cfg(cfg), latent, covariance_test=True, nodes_all=nodes)
        recons = orc.decoder_forward(g["dec_state"], dec_cfg(cfg), latent)
    for key, val in g["latent"].items():
        assert rel_err(latent[eval(key)], val) < TOL, ("latent", key)
    assert rel_err(recons, g["recons"]) < TOL
    assert len(nodes_all) == len(g["nodes_all"])
    for i, (mine, ref) in enumerate(zip(nodes_all, g["nodes_all"])):
        assert [str(k) for k in mine.keys()] == list(ref.keys()), (i, list(mine.keys()), list(ref.keys()))
        for key, val in ref.items():
            assert rel_err(mine[eval(key)], val) < TOL, (i, key)


@pytest.mark.parametrize("name", CASES)
def test_loss_and_gradients_match_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    enc_sd = {k: v.clone().requires_grad_(True) for k, v in g["enc_state"].items()}
    dec_sd = {k: v.clone().requires_grad_(True) for k, v in g["dec_state"].items()}
    loss, _, _ = orc.training_step(enc_sd, dec_sd, enc_cfg(cfg), dec_cfg(cfg), g["batch"], l1_lambda=1e-8,
                                   get_real_method=cfg.get("get_real", "sum"))
    assert abs(loss.item() - g["loss"].item()) < 1e-12 * abs(g["loss"].item())
    loss.backward()
    for sd, grads in ((enc_sd, g["grads_enc"]), (dec_sd, g["grads_dec"])):
        for k, ref in grads.items():
            mine = sd[k].grad
            if ref is None:
                # the L1 term touches every parameter in the reference too, so None never happens
                assert mine is None, k
                continue
            assert mine is not None, k
            assert rel_err(mine, ref) < 1e-10, (k, rel_err(mine, ref))


def sum_latent_loss(latent):
    """The fixed quadratic form tests/golden/make_golden.py differentiates for the encoder-only 'sum' fixture."""
    return sum((v * v).sum() * (0.5 + i) for i, v in enumerate(latent.values()))


def test_sum_pooling_encoder_matches_reference():
    """'sum' pooling (lgn_encoder.py:421-427) keeps a spurious axis the reference's decoder cannot consume: encoder only."""
    g = load_golden("sum_n6")
    cfg = g["cfg"]
    enc_sd = {k: v.clone().requires_grad_(True) for k, v in g["enc_state"].items()}
    latent = orc.encoder_forward(enc_sd, enc_cfg(cfg), g["batch"])
    for key, val in g["latent"].items():
        assert latent[eval(key)].shape == val.shape
        assert rel_err(latent[eval(key)], val) < TOL, ("latent", key)
    loss = sum_latent_loss(latent)
    assert abs(loss.item() - g["loss"].item()) < 1e-12 * abs(g["loss"].item())
    loss.backward()
    for k, ref in g["grads_enc"].items():
        if ref is None:
            assert enc_sd[k].grad is None or enc_sd[k].grad.abs().max().item() == 0.0, k
        else:
            assert rel_err(enc_sd[k].grad, ref) < 1e-10, k


def test_anomaly_scores_match_reference():
    """The Cartesian-family anomaly scores against the unmodified reference's functions (tests/golden/make_scores_golden.py)."""
    g = load_golden("anomaly_scores")
    mine = orc.anomaly_scores_cartesian(g["recons"], g["target"])
    for short, name in (("chamfer_cartesian", "chamfer_particle_cartesian"), ("mse_cartesian", "mse_particle_cartesian"),
                        ("chamfer_lorentz", "chamfer_particle_lorentz"), ("mse_lorentz", "mse_particle_lorentz"),
                        ("jet_cartesian", "jet_cartesian"), ("jet_lorentz", "jet_lorentz")):
        assert rel_err(mine[name], g[short]) < TOL, name
