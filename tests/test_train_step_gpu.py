"""FusedTrainStep (static buffers + CUDA graph) against the module/autograd path and the reference's golden vectors."""
import pytest
import torch

from tests.helpers import assert_grads_close, load_golden, rel_err
from tests.test_models_gpu import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_train_step_matches_reference(use_graph):
    from lgn_autoencoder_b200.train import FusedTrainStep
    dev = torch.device("cuda:0")
    g, enc, dec, batch = load("cfg1_b3", dev)
    step = FusedTrainStep(enc, dec, batch["p4"].shape[0], l1_lambda=1e-8, normalize=False, use_graph=use_graph)
    for _ in range(3):   # replays must be idempotent
        loss = step.step(batch["p4"])
    torch.cuda.synchronize()
    assert abs(loss.item() - g["loss"].item()) < 1e-10 * abs(g["loss"].item())
    assert rel_err(step.recon, g["recons"]) < 1e-10
    assert rel_err(step.latent11, g["latent"]["(1, 1)"]) < 1e-10
    assert_grads_close({k: p.grad for k, p in enc.named_parameters()}, g["grads_enc"])
    assert_grads_close({k: p.grad for k, p in dec.named_parameters()}, g["grads_dec"])


def test_fused_train_step_padded_jets_and_optimizer():
    """labels mask, normalisation, and an optimizer step between replays (weights are read from the live buffers)."""
    from lgn_autoencoder_b200.train import FusedTrainStep, training_step
    dev = torch.device("cuda:0")
    g, enc, dec, batch = load("pad_n8", dev)
    b = batch["p4"].shape[0]
    step = FusedTrainStep(enc, dec, b, l1_lambda=1e-8, normalize=True, use_labels=True, use_graph=True)
    opt = torch.optim.SGD(list(enc.parameters()) + list(dec.parameters()), lr=1e-4)
    l0 = step.step(batch["p4"], batch["labels"]).item()
    # the same step through the module API + autograd
    g_fused = {k: p.grad.clone() for k, p in list(enc.named_parameters()) + list(dec.named_parameters())}
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    loss, _, _ = training_step(enc, dec, batch["p4"], labels=batch["labels"], l1_lambda=1e-8)
    loss.backward()
    assert abs(loss.item() - l0) < 1e-12 * abs(l0)
    g_mod = {k: p.grad.clone() for k, p in list(enc.named_parameters()) + list(dec.named_parameters())}
    gmax = max(v.abs().max().item() for v in g_mod.values())
    for k in g_mod:
        assert (g_mod[k] - g_fused[k]).abs().max().item() <= 1e-12 * gmax, k
    step._bind_grads()
    l1 = step.step(batch["p4"], batch["labels"]).item()
    assert l1 == l0
    opt.step()
    l2 = step.step(batch["p4"], batch["labels"]).item()
    assert l2 != l0


def test_step_host_matches_step():
    """End-to-end entry (pinned staging + in-graph copies) gives the same loss and gradients as the device-resident step."""
    from lgn_autoencoder_b200.train import FusedTrainStep
    dev = torch.device("cuda:0")
    g, enc, dec, batch = load("cfg1_b3", dev)
    step = FusedTrainStep(enc, dec, batch["p4"].shape[0], l1_lambda=1e-8, normalize=False, use_graph=True)
    l_dev = step.step(batch["p4"]).item()
    g_dev = step.g_e.clone()
    for _ in range(2):
        l_host = step.step_host(batch["p4"].cpu())
    assert l_host == l_dev
    assert torch.equal(step.g_e, g_dev)
    assert abs(l_host - g["loss"].item()) < 1e-10 * abs(g["loss"].item())
