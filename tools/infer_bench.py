"""Developer: forward-only (anomaly-scoring) throughput with FusedInference, e.g. cfg-5 shapes: 150 particles per jet,
8192 jets sharded over 8 GPUs = 1024 per GPU.  Usage: python tools/infer_bench.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import synthetic_jets
from lgn_autoencoder_b200.flop_model import step_flops_per_jet
from lgn_autoencoder_b200.train import FusedInference
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
bench.CFG["n"] = N
dev = torch.device("cuda:0")
enc, dec = bench.build_models(dev)
inf = FusedInference(enc, dec, B)
p4 = synthetic_jets(B, N, seed=1).to(dev)
for _ in range(3):
    s = inf.score(p4)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    s = inf.run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
fl = step_flops_per_jet(N, bench.CFG["enc_channels"], bench.CFG["dec_channels"], backward=False)
print(f"N={N} B={B}: {ms:.3f} ms per forward -> {B / ms * 1e3:.0f} jets/s = {B * fl / (ms * 1e-3) / 1e12:.1f} TFLOP/s fp64 (reference-faithful count); "
      f"mean score {s.mean().item():.6g}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
