"""The training step of utils/train.py:283-327 (the reference's hot loop body) on the fused path, plus the
data-parallel gradient exchange: jets are independent, so the batch is sharded over ranks
with no data-path collective; the only exchange is one NCCL all-reduce(SUM) of the flat gradient per model."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import fused


def training_step(encoder, decoder, p4, labels=None, l1_lambda: float = 1e-8, normalize: bool = True, l1_scale: float = 1.0):
    """normalize_p4('overall_max') -> encoder -> decoder -> get_real('sum') + ChamferLoss (sum over the batch)
    + l1_lambda * (|theta_enc|_1 + |theta_dec|_1).  Returns (loss, reconstruction (2,B,N,4), normalised input)."""
    if normalize:
        p4, _ = fused.normalize_p4(p4)
    batch = {"p4": p4}
    if labels is not None:
        batch["labels"] = labels
    latent = encoder(batch, covariance_test=False)
    recon = decoder(latent, covariance_test=False)
    loss = fused.chamfer_loss(recon, p4)
    if l1_lambda:
        loss = loss + (l1_lambda * l1_scale) * (encoder.l1_norm() + decoder.l1_norm())
    return loss, recon, p4


def allreduce_gradients(*models, group=None):
    """SUM-all-reduce the gradients of every model as one flat fp64 bucket each (chamfer is a sum over jets, so no
    division by the world size; scale the L1 term by 1/world on every rank instead -- ``l1_scale`` above)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for model in models:
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        if not grads:
            continue
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
