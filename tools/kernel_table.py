"""Developer tool: per-kernel table of one FusedTrainStep (eager, event-timed).  LGAE_B200_LIB selects a variant build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CFG, build_models, synthetic_jets
from lgn_autoencoder_b200 import _lib
from lgn_autoencoder_b200.train import FusedTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
enc, dec = build_models(dev)
st = FusedTrainStep(enc, dec, B, l1_lambda=1e-8, use_graph=True, get_real="sum")
st.load(synthetic_jets(B, 30, seed=3))
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
def eager():
    flush.zero_(); st._launch()
for _ in range(3): eager()
k = _lib.kernel_timings(eager, reps=10)
tot = sum(n * ms for n, ms in k.values()) / 10
for name, (n, ms) in sorted(k.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
    print(f"{name:18s} {n/10:4.0f} x {ms*1e3:7.1f} us  {n*ms/10/tot*100:5.1f}%")
for _ in range(5): st.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): st.run()
b.record(); torch.cuda.synchronize()
print(f"{os.environ.get('LGAE_B200_LIB','default')}: kernels {tot*1e3:.0f} us; graph replay {a.elapsed_time(b)/50*1e3:.0f} us/step -> {B/(a.elapsed_time(b)/50)*1e3:.0f} jets/s")
