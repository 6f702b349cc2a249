"""Host logic of the data-parallel path on CPU: two gloo ranks, each with half the batch, must
end up with the single-process full-batch gradient after allreduce_gradients -- SUM without division (the loss is a sum
over jets) and the L1 term scaled by 1/world on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4)).double()


def _loss(model, x, l1_scale):
    """sum-over-samples loss + l1 (the shape of utils/train.py:416-494 with ChamferLoss)."""
    rec = ((model(x) - x) ** 2).sum()
    return rec + 1e-3 * l1_scale * sum(p.abs().sum() for p in model.parameters())


def _worker(rank, world, port, x, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lgn_autoencoder_b200.train import allreduce_gradients
        model = _model()
        shard = x[rank::world]
        _loss(model, shard, 1.0 / world).backward()
        allreduce_gradients(model)
        out[rank] = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch_gradient():
    torch.manual_seed(1)
    x = torch.randn(16, 4, dtype=torch.float64)
    ref_model = _model()
    _loss(ref_model, x, 1.0).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, x, out), nprocs=2, join=True)
    assert torch.allclose(out[0], ref, rtol=1e-13, atol=1e-15)
    assert torch.equal(out[0], out[1])    # replicas stay bit-identical


def test_allreduce_is_a_noop_without_process_group():
    from lgn_autoencoder_b200.train import allreduce_gradients
    m = _model()
    _loss(m, torch.randn(4, 4, dtype=torch.float64), 1.0).backward()
    g0 = [p.grad.clone() for p in m.parameters()]
    allreduce_gradients(m)
    assert all(torch.equal(a, p.grad) for a, p in zip(g0, m.parameters()))
