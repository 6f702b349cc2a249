"""The closed-form / neighbour-sum algebra the CUDA kernels implement (tests/stage_math.py), checked
on the CPU against the golden vectors of the unmodified reference."""
import pytest
import torch

from tests import stage_math as sm
from tests.helpers import assert_grads_close, load_golden, rel_err


@pytest.mark.parametrize("name", ["cfg1_b3", "pad_n8", "mean_n6"])
def test_stage_algebra_matches_reference(name):
    g = load_golden(name)
    cfg = g["cfg"]
    out = sm.model_step(g["enc_state"], g["dec_state"], g["batch"], map_to_latent=cfg["map_to_latent"], l1_lambda=1e-8)
    lat00 = sm.planar_to_c(g["latent"]["(0, 0)"])[:, 0, :, 0]
    lat11 = sm.planar_to_c(g["latent"]["(1, 1)"])[:, 0]
    assert rel_err(torch.view_as_real(out["lat00"]), torch.view_as_real(lat00)) < 1e-10
    assert rel_err(torch.view_as_real(out["lat11"]), torch.view_as_real(lat11)) < 1e-10
    assert rel_err(sm.c_to_planar(out["recons"]), g["recons"]) < 1e-10
    assert abs(out["loss"].item() - g["loss"].item()) < 1e-10 * abs(g["loss"].item())
    assert_grads_close(out["grads_enc"], g["grads_enc"])
    assert_grads_close(out["grads_dec"], g["grads_dec"])
