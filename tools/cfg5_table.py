"""Developer: per-kernel table of the forward-only scoring pass at cfg-5 shapes (150 particles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lgn_autoencoder_b200 import _lib
from lgn_autoencoder_b200.train import FusedInference
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bench.CFG["n"] = 150
dev = torch.device("cuda:0")
enc, dec = bench.build_models(dev)
inf = FusedInference(enc, dec, B, get_real="sum", use_graph=False)
p4 = bench.synthetic_jets(B, 150, seed=1).to(dev)
for _ in range(2):
    inf.score(p4)
torch.cuda.synchronize()
k = _lib.kernel_timings(inf.run, reps=3)
tot = sum(n * t for n, t in k.values()) / 3
for name, (n, t) in sorted(k.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
    print(f"{name:20s} {n/3:4.0f} x {t*1e3:9.1f} us  {n*t/3/tot*100:5.1f}%")
print(f"total {tot:.2f} ms for {B} jets -> {B/tot*1e3:.0f} jets/s")
