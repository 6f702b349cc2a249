"""Developer: forward-only (anomaly-scoring) throughput with FusedInference, e.g. BASELINE.json configs[4] (cfg-5): 150 particles
per jet, 8192 jets sharded over 8 GPUs = 1024 per GPU, no communication on the data path.
Usage: python tools/infer_bench.py [N] [B_per_gpu]            (one GPU)
       python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/infer_bench.py 150 1024"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from bench import synthetic_jets
from lgn_autoencoder_b200.flop_model import step_flops_per_jet
from lgn_autoencoder_b200.train import FusedInference
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bench.CFG["n"] = N
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
enc, dec = bench.build_models(dev)
inf = FusedInference(enc, dec, B, get_real="sum")
p4 = synthetic_jets(B, N, seed=1 + rank).to(dev)     # this rank's shard of the global batch
for _ in range(3):
    s = inf.score(p4)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    s = inf.run()
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the slowest rank sets the job's time
ms = t.item()
fl = step_flops_per_jet(N, bench.CFG["enc_channels"], bench.CFG["dec_channels"], backward=False)
if rank == 0:
    print(f"N={N} B={B}/GPU x {world} GPU: {ms:.3f} ms per forward (max over ranks) -> {B * world / ms * 1e3:.0f} jets/s = "
          f"{B * world * fl / (ms * 1e-3) / 1e12:.1f} TFLOP/s fp64 (reference-faithful count); mean score {s.mean().item():.6g}; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if world > 1:
    del inf
    dist.destroy_process_group()
