"""Edge cases of the fused path against the CPU oracle on seeded inputs (sizes the oracle finishes in seconds): a single
particle, the 32-particle boundary of the adjoint, fully padded jets, every pooling mode, odd channel counts, batch 1 and the
empty batch.  The oracle itself is pinned to the reference by tests/test_oracle_golden.py."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _build(cfg, dev):
    from lgn_autoencoder_b200.models import LGNDecoder, LGNEncoder
    torch.manual_seed(cfg["seed"])
    common = dict(maxdim=[2], num_basis_fn=10, max_zf=[1], weight_init="randn", level_gain=[1.0], activation="leakyrelu", mlp=True,
                  mlp_depth=cfg["mlp_depth"], mlp_width=cfg["mlp_width"], device=torch.device("cpu"), dtype=torch.float64)
    enc = LGNEncoder(num_input_particles=cfg["n"], tau_input_scalars=1, tau_input_vectors=1, tau_latent_scalars=cfg["tau_s"],
                     tau_latent_vectors=cfg["tau_v"], num_channels=cfg["enc_channels"], jet_features=False,
                     map_to_latent=cfg["map_to_latent"], **common)
    mult = 2 if cfg["map_to_latent"] == "min&max" else 1
    dec = LGNDecoder(tau_latent_scalars=cfg["tau_s"] * mult, tau_latent_vectors=cfg["tau_v"] * mult, num_output_particles=cfg["n"],
                     tau_output_scalars=1, tau_output_vectors=1, num_channels=cfg["dec_channels"], cg_dict=enc.cg_dict, **common)
    return enc.to(dev), dec.to(dev)


CASES = {
    "single_particle": dict(seed=11, batch=3, n=1, enc_channels=[2, 2, 2], dec_channels=[2, 2, 2], tau_s=1, tau_v=2, map_to_latent="min&max",
                            mlp_depth=2, mlp_width=2, pad=False),
    "n32_boundary": dict(seed=12, batch=2, n=32, enc_channels=[1, 2, 3], dec_channels=[3, 2, 1], tau_s=1, tau_v=3, map_to_latent="mean",
                         mlp_depth=2, mlp_width=3, pad=True),
    "max_pool_odd_channels": dict(seed=13, batch=4, n=9, enc_channels=[3, 1, 4, 2], dec_channels=[2, 4, 1, 3], tau_s=2, tau_v=1,
                                  map_to_latent="max", mlp_depth=3, mlp_width=2, pad=True),
    "min_pool": dict(seed=14, batch=2, n=6, enc_channels=[2, 3], dec_channels=[3, 2], tau_s=1, tau_v=2, map_to_latent="min", mlp_depth=1,
                     mlp_width=4, pad=False),
    "batch_one_wide_mlp": dict(seed=15, batch=1, n=12, enc_channels=[2, 2, 2, 2], dec_channels=[2, 2, 2, 2], tau_s=1, tau_v=2, map_to_latent="mean",
                          mlp_depth=6, mlp_width=6, pad=False),
}


def _jets(cfg):
    from oracle import lgae_oracle as orc
    data = orc.synthetic_jets(cfg["batch"], cfg["n"], seed=cfg["seed"] + 1, mass_scale=0.05, pad=cfg["pad"])
    if cfg["pad"]:
        # one jet padded completely: every particle masked (labels 0, p4 0)
        data["p4"][0] = 0.0
        data["labels"][0] = 0.0
    return data


@pytest.mark.parametrize("name", list(CASES))
def test_forward_backward_match_oracle(name):
    from lgn_autoencoder_b200 import fused
    from oracle import lgae_oracle as orc
    cfg = CASES[name]
    dev = torch.device("cuda:0")
    enc, dec = _build(cfg, dev)
    assert enc.fused and dec.fused
    data = _jets(cfg)
    p4n, _ = orc.normalize_p4_overall_max(data["p4"])
    batch_cpu = dict(data, p4=p4n)
    batch = {k: v.to(dev) for k, v in batch_cpu.items() if k in ("p4", "labels")}
    # product path: module API + autograd
    latent = enc(batch)
    recon = dec(latent)
    loss = fused.chamfer_loss(recon, batch["p4"], "sum") + 1e-8 * (enc.l1_norm() + dec.l1_norm())
    loss.backward()
    # oracle on the same weights
    enc_sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    dec_sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    ecfg = dict(num_channels=cfg["enc_channels"], maxdim=[2], max_zf=[1], map_to_latent=cfg["map_to_latent"])
    dcfg = dict(num_channels=cfg["dec_channels"], maxdim=[2], max_zf=[1])
    ref_loss, ref_latent, ref_recon = orc.training_step(enc_sd, dec_sd, ecfg, dcfg, batch_cpu, l1_lambda=1e-8)
    ref_loss.backward()
    for key in ((0, 0), (1, 1)):
        assert rel_err(latent[key], ref_latent[key]) < 1e-10, (name, key)
    assert rel_err(recon, ref_recon) < 1e-10, name
    assert abs(loss.item() - ref_loss.item()) <= 1e-10 * abs(ref_loss.item()), name
    gmax = max(v.grad.abs().max().item() for v in list(enc_sd.values()) + list(dec_sd.values()) if v.grad is not None)
    for model, sd in ((enc, enc_sd), (dec, dec_sd)):
        for k, p in model.named_parameters():
            ref = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
            err = (p.grad.detach().cpu() - ref).abs().max().item()
            assert err <= 1e-10 * max(ref.abs().max().item(), 1e-4 * gmax), (name, k, err)


def test_empty_batch_is_a_no_op():
    """Batch 0 through the C entry points: nothing is launched, the loss is 0 and the parameter gradient is the L1 term alone.
    (At module level an empty batch is not representable: like the reference's, GVec drops empty parts.)"""
    from lgn_autoencoder_b200 import fused
    dev = torch.device("cuda:0")
    enc, dec = _build(CASES["min_pool"], dev)
    theta, _ = enc._flat_params()
    p4 = torch.zeros((0, 6, 4), dtype=torch.float64, device=dev)
    lat00, lat11, ws, sel = fused.encoder_forward_raw(enc._plan, theta, p4, None)
    assert lat11.shape[1] == 0
    th_d, _ = dec._flat_params()
    recon, _, ws_d = fused.decoder_forward_raw(dec._plan, th_d, lat11)
    assert recon.shape == (2, 0, 6, 4)
    assert fused.chamfer_loss(recon, p4, "sum").item() == 0.0
    g = fused.encoder_backward_raw(enc._plan, theta, p4, None, ws, sel, None, lat11)
    assert torch.count_nonzero(g).item() == 0


def test_more_than_32_particles():
    """The fused adjoint holds one particle per lane: the one-call step refuses N > 32 loudly, the raw C entry point returns
    LGAE_E_UNSUPPORTED before launching anything, and the module API trains such models through the layer-level composite
    (as the reference can, lgn_encoder.py:255-336 has no particle limit) -- never through a silent CPU path."""
    from lgn_autoencoder_b200 import fused
    from lgn_autoencoder_b200.train import FusedTrainStep
    dev = torch.device("cuda:0")
    cfg = dict(CASES["min_pool"], n=40)
    enc, dec = _build(cfg, dev)
    with pytest.raises(NotImplementedError):
        FusedTrainStep(enc, dec, 2, get_real="sum")
    p4 = torch.rand((2, 40, 4), dtype=torch.float64, device=dev) + 0.1
    theta, _ = enc._flat_params()
    lat00, lat11, ws, sel = fused.encoder_forward_raw(enc._plan, theta, p4, None)
    with pytest.raises(NotImplementedError):
        fused.encoder_backward_raw(enc._plan, theta, p4, None, ws, sel, None, torch.ones_like(lat11))
    with torch.no_grad():
        rec_fused = dec(enc({"p4": p4}))            # forward only: the fused kernels (blocks of 32 particles)
    rec = dec(enc({"p4": p4}))                      # with autograd: the layer-level composite
    assert (rec - rec_fused).abs().max().item() <= 1e-10 * rec_fused.abs().max().item()
    rec.sum().backward()
    grads = [p.grad for p in list(enc.parameters()) + list(dec.parameters()) if p.grad is not None]
    assert grads and all(torch.isfinite(g).all() for g in grads) and any(g.abs().max().item() > 0 for g in grads)
