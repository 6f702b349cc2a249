"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Per-tensor max-norm relative error (SURVEY.md section 7 'hard parts'): max|a-b| / max|b|."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = b.abs().max().item()
    num = (a - b).abs().max().item()
    if denom == 0.0:
        return num
    return num / denom


def enc_cfg(cfg):
    return dict(num_channels=cfg["enc_channels"], maxdim=[cfg["maxdim"]], max_zf=[1], map_to_latent=cfg["map_to_latent"])


def dec_cfg(cfg):
    return dict(num_channels=cfg["dec_channels"], maxdim=[cfg["maxdim"]], max_zf=[1])


def grad_errors(mine: dict, ref: dict):
    """Per-tensor gradient errors.  Returns {name: (per_tensor_rel, rel_to_global_max)}.  The second
    number divides by the largest gradient entry of the whole model: a few scalar-path parameters
    of the near-massless jets have gradients ~1e-8 of that scale which are themselves the result of
    cancellations ('bit-level sensitivity'), so their per-tensor relative error
    measures the summation order, not the kernel."""
    gmax = max(v.abs().max().item() for v in ref.values() if v is not None)
    out = {}
    for k, v in ref.items():
        if v is None:
            continue
        a = mine[k].detach().double().cpu()
        b = v.detach().double().cpu()
        out[k] = (rel_err(a, b), (a - b).abs().max().item() / gmax)
    return out


def assert_grads_close(mine: dict, ref: dict, tol=1e-10, tol_global=1e-14):
    """A tensor passes if it is within ``tol`` of the reference in per-tensor max-norm, or -- for the
    tiny cancellation-dominated gradients -- within ``tol_global`` of the model-wide gradient scale."""
    errs = grad_errors(mine, ref)
    bad = {k: e for k, e in errs.items() if not (e[0] < tol or e[1] < tol_global)}
    assert not bad, bad
