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
