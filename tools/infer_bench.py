"""Developer: forward-only throughput at cfg-5 shapes (150 particles per jet) through the module API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import synthetic_jets
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
bench.CFG["n"] = N
dev = torch.device("cuda:0")
enc, dec = bench.build_models(dev)
from lgn_autoencoder_b200 import fused
p4 = synthetic_jets(B, N, seed=1).to(dev)
p4, _ = fused.normalize_p4(p4)
def fwd():
    with torch.no_grad():
        lat = enc({"p4": p4})
        rec = dec(lat)
        return fused.chamfer_per_jet(rec, p4)
for _ in range(3): s = fwd()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): s = fwd()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"N={N} B={B}: {ms:.3f} ms per forward -> {B / ms * 1e3:.0f} jets/s; score mean {s.mean().item():.6g}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
