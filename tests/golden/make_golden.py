"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference is imported from /root/reference (read-only); matplotlib / jetnet, which only its
``utils`` package needs, are stubbed.  /root/repo must NOT be on sys.path while this runs: the
repo's own ``lgn`` shim package would shadow the reference's namespace package.  Nothing in the
test-suite reads /root/reference; the tests consume the .pt files written here.

Each fixture holds: the constructor config, inputs, both state_dicts, latent, reconstruction,
all ``nodes_all`` GVecs of a ``covariance_test=True`` pass, the training loss of
``utils/train.py:283-327`` (chamfer + 1e-8 * L1) and the gradient of every parameter.
"""
import os
import sys
import types

REF = os.environ.get("LGAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(HERE, "..", ".."))]
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot", "jetnet", "jetnet.losses"):
    sys.modules.setdefault(name, types.ModuleType(name))

import torch  # noqa: E402

import lgn.models  # noqa: E402,F401  (must come before lgn.g_lib: import-order quirk)
from lgn.cg_lib import CGDict  # noqa: E402
from lgn.models import LGNDecoder, LGNEncoder  # noqa: E402
from utils.losses.chamfer_loss.chamfer_loss import ChamferLoss  # noqa: E402
from utils.normalize_p4 import normalize_p4  # noqa: E402
from utils.utils import get_real  # noqa: E402

assert lgn.models.__file__.startswith(REF), lgn.models.__file__


def synthetic_jets(batch, n, seed=0, mass_scale=1e-6, pad=False):
    """Same generator as oracle.lgae_oracle.synthetic_jets; duplicated so that this
    script depends on nothing but the reference."""
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
        labels = torch.arange(n).unsqueeze(0) < nobj.unsqueeze(1)
        p4 = p4 * labels.unsqueeze(-1)
        out = {"p4": p4, "labels": labels.to(torch.float64), "Nobj": nobj}
    return out


def build(cfg):
    """Constructor kwargs as utils/initialize.py:91-141."""
    torch.manual_seed(cfg["seed"])
    dev = torch.device("cpu")
    common = dict(maxdim=[cfg["maxdim"]], num_basis_fn=10, max_zf=[1], weight_init="randn", level_gain=[1.0],
                  activation="leakyrelu", mlp=True, mlp_depth=cfg.get("mlp_depth", 6), mlp_width=cfg.get("mlp_width", 6),
                  device=dev, dtype=torch.float64)
    enc = LGNEncoder(num_input_particles=cfg["n"], tau_input_scalars=1, tau_input_vectors=1,
                     tau_latent_scalars=cfg["tau_s"], tau_latent_vectors=cfg["tau_v"], num_channels=cfg["enc_channels"],
                     jet_features=False, map_to_latent=cfg["map_to_latent"], **common)
    mult = len(cfg["map_to_latent"].split("&")) if "&" in cfg["map_to_latent"] else 1
    if cfg["map_to_latent"] == "mix":
        mult = 1
    dec = LGNDecoder(tau_latent_scalars=cfg["tau_s"] * mult, tau_latent_vectors=cfg["tau_v"] * mult,
                     num_output_particles=cfg["n"], tau_output_scalars=1, tau_output_vectors=1,
                     num_channels=cfg["dec_channels"], cg_dict=enc.cg_dict, **common)
    return enc, dec


def gvec_to_dict(g):
    return {str(k): v.detach().clone() for k, v in g.items()}


def run(cfg):
    enc, dec = build(cfg)
    data = synthetic_jets(cfg["batch"], cfg["n"], seed=cfg["seed"] + 1, mass_scale=cfg["mass_scale"], pad=cfg["pad"])
    p4n, _ = normalize_p4(data["p4"], "overall_max")
    batch = dict(data)
    batch["p4"] = p4n

    latent = enc(batch, covariance_test=False)
    if cfg.get("encoder_only"):
        # the reference's own decoder cannot consume a 'sum' latent (its spurious extra axis breaks RadPolyTrig.forward,
        # position_levels.py:171-207): pin the encoder alone, with the gradient of a fixed quadratic form of the latent
        loss = sum((v * v).sum() * (0.5 + i) for i, v in enumerate(latent.values()))
        loss.backward()
        return {"cfg": cfg, "batch": {k: v.clone() for k, v in batch.items()},
                "enc_state": {k: v.detach().clone() for k, v in enc.state_dict().items()},
                "dec_state": {k: v.detach().clone() for k, v in dec.state_dict().items()},
                "latent": gvec_to_dict(latent), "loss": loss.detach().clone(),
                "grads_enc": {k: (None if p.grad is None else p.grad.clone()) for k, p in enc.named_parameters()},
                "torch_version": torch.__version__}
    recons = dec(latent, covariance_test=False)
    p4_recons = get_real(recons, cfg.get("get_real", "sum"))
    loss = ChamferLoss(device=torch.device("cpu"))(p4_recons, p4n) + 1e-8 * (enc.l1_norm() + dec.l1_norm())
    loss.backward()
    grads_enc = {k: (None if p.grad is None else p.grad.clone()) for k, p in enc.named_parameters()}
    grads_dec = {k: (None if p.grad is None else p.grad.clone()) for k, p in dec.named_parameters()}

    with torch.no_grad():
        lat2, nodes_enc = enc(batch, covariance_test=True)
        gen, nodes_all = dec(lat2, covariance_test=True, nodes_all=nodes_enc)

    out = {
        "cfg": cfg,
        "batch": {k: v.clone() for k, v in batch.items()},
        "enc_state": {k: v.detach().clone() for k, v in enc.state_dict().items()},
        "dec_state": {k: v.detach().clone() for k, v in dec.state_dict().items()},
        "latent": gvec_to_dict(latent),
        "recons": recons.detach().clone(),
        "loss": loss.detach().clone(),
        "grads_enc": grads_enc,
        "grads_dec": grads_dec,
        "nodes_all": [gvec_to_dict(g) for g in nodes_all],
        "generated": gvec_to_dict(gen),
        "torch_version": torch.__version__,
    }
    return out


CONFIGS = {
    # cfg-1 of BASELINE.json at a tiny batch
    "cfg1_b3": dict(seed=0, batch=3, n=30, maxdim=2, enc_channels=[3, 3, 4, 4], dec_channels=[4, 4, 3, 3], tau_s=1, tau_v=8,
                    map_to_latent="min&max", mass_scale=1e-6, pad=False),
    # zero-padded jets with labels, massive particles, odd channel counts
    "pad_n8": dict(seed=3, batch=4, n=8, maxdim=2, enc_channels=[2, 3, 2, 3], dec_channels=[3, 2, 3, 2], tau_s=2, tau_v=3,
                   map_to_latent="min&max", mass_scale=0.1, pad=True, mlp_depth=3, mlp_width=2),
    # mean aggregation, two levels only
    "mean_n6": dict(seed=5, batch=2, n=6, maxdim=2, enc_channels=[2, 2, 3], dec_channels=[3, 2, 2], tau_s=1, tau_v=2,
                    map_to_latent="mean", mass_scale=0.1, pad=False, mlp_depth=2, mlp_width=2),
    # wide family: maxdim 3 + 'mix' latent map
    "md3_mix_n5": dict(seed=7, batch=2, n=5, maxdim=3, enc_channels=[2, 2, 3, 3], dec_channels=[3, 3, 2, 2], tau_s=1, tau_v=2,
                       map_to_latent="mix", mass_scale=0.1, pad=False, mlp_depth=2, mlp_width=2),
    # more than 32 particles per jet (two particle blocks per CTA in the forward kernels; cfg-5 has 150): forward parity
    "n40_b2": dict(seed=9, batch=2, n=40, maxdim=2, enc_channels=[2, 2, 3, 3], dec_channels=[3, 3, 2, 2], tau_s=1, tau_v=4,
                   map_to_latent="min&max", mass_scale=1e-6, pad=True, mlp_depth=2, mlp_width=2),
    # cfg-4 of BASELINE.json at its real shape (30 particles, maxdim 3, 6-6-8-8 / 8-8-6-6, 'mix' latent map, MLP 6 x 6), two jets
    "cfg4_b2": dict(seed=11, batch=2, n=30, maxdim=3, enc_channels=[6, 6, 8, 8], dec_channels=[8, 8, 6, 6], tau_s=1, tau_v=8,
                    map_to_latent="mix", mass_scale=1e-6, pad=False),
    # the remaining pooling modes (lgn_encoder.py:419-476): 'sum' (keeps a spurious axis) and a '+' combination
    "sum_n6": dict(seed=13, batch=2, n=6, maxdim=2, enc_channels=[2, 2, 3], dec_channels=[3, 2, 2], tau_s=1, tau_v=2,
                   map_to_latent="sum", mass_scale=0.1, pad=False, mlp_depth=2, mlp_width=2, encoder_only=True),
    "minplusmax_n6": dict(seed=15, batch=2, n=6, maxdim=2, enc_channels=[2, 2, 3], dec_channels=[3, 2, 2], tau_s=1, tau_v=2,
                          map_to_latent="min+max", mass_scale=0.1, pad=True, mlp_depth=2, mlp_width=2),
    # get_real 'real' (the reference's default, main.py:295-300) and 'norm'
    "real_n6": dict(seed=17, batch=3, n=6, maxdim=2, enc_channels=[2, 2, 3], dec_channels=[3, 2, 2], tau_s=1, tau_v=2,
                    map_to_latent="min&max", mass_scale=0.1, pad=False, mlp_depth=2, mlp_width=2, get_real="real"),
    "norm_n6": dict(seed=17, batch=3, n=6, maxdim=2, enc_channels=[2, 2, 3], dec_channels=[3, 2, 2], tau_s=1, tau_v=2,
                    map_to_latent="min&max", mass_scale=0.1, pad=False, mlp_depth=2, mlp_width=2, get_real="norm"),
}


def main():
    only = sys.argv[1:]
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        out = run(cfg)
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(out, path)
        print(name, "loss", float(out["loss"]), os.path.getsize(path) // 1024, "KiB")
    if only:
        return
    cg = CGDict(maxdim=3, dtype=torch.float64)
    torch.save({str(k): {str(kk): vv.clone() for kk, vv in v.items()} for k, v in cg.items()}, os.path.join(HERE, "cg_maxdim3.pt"))
    # scalar KATs on the basis changes
    from lgn.cg_lib.zonal_functions import p_to_rep, rep_to_p, normsq4
    p = torch.tensor([[2.0, 0.3, -0.4, 1.2], [1.0, 0.6, 0.0, 0.8]], dtype=torch.float64)
    rep = p_to_rep(p)[(1, 1)]
    torch.save({"p": p, "rep": rep.clone(), "back": rep_to_p(rep.squeeze(-2)).clone(), "normsq4": normsq4(p).clone()},
               os.path.join(HERE, "basis_kat.pt"))


if __name__ == "__main__":
    main()
