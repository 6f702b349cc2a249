"""Developer: training-step throughput of BASELINE.json configs[3] (cfg-4: maxdim 3, enc 6 6 8 8 / dec 8 8 6 6, 'mix' latent map,
bs 1024) on the layer-level kernels (csrc/lgae_cg.cu, lgae_layers.cu) through the module API + autograd, with a per-kernel table.
Usage: python tools/cfg4_bench.py [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_jets
from lgn_autoencoder_b200 import _lib
from lgn_autoencoder_b200.flop_model import step_flops_per_jet
from lgn_autoencoder_b200.models import LGNDecoder, LGNEncoder
from lgn_autoencoder_b200.train import training_step

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(_pos[0]) if len(_pos) > 0 else 1024
steps = int(_pos[1]) if len(_pos) > 1 else 5
N, ENC, DEC = 30, [6, 6, 8, 8], [8, 8, 6, 6]
dev = torch.device("cuda:0")
torch.manual_seed(0)
common = dict(maxdim=[3], num_basis_fn=10, max_zf=[1], weight_init="randn", level_gain=[1.0], activation="leakyrelu", mlp=True, mlp_depth=6,
              mlp_width=6, device=dev, dtype=torch.float64)
enc = LGNEncoder(num_input_particles=N, tau_input_scalars=1, tau_input_vectors=1, tau_latent_scalars=1, tau_latent_vectors=8,
                 num_channels=ENC, jet_features=False, map_to_latent="mix", **common)
dec = LGNDecoder(tau_latent_scalars=1, tau_latent_vectors=8, num_output_particles=N, tau_output_scalars=1, tau_output_vectors=1,
                 num_channels=DEC, cg_dict=enc.cg_dict, **common)
p4 = synthetic_jets(B, N, seed=2).to(dev)


def step():
    for m in (enc, dec):
        for p in m.parameters():
            p.grad = None
    loss, _, _ = training_step(enc, dec, p4, get_real="sum")
    loss.backward()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    loss = step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
k = _lib.kernel_timings(step, reps=2)
tot = sum(n * t for n, t in k.values()) / 2
for name, (n, t) in sorted(k.items(), key=lambda kv: -kv[1][0] * kv[1][1])[:12]:
    print(f"{name:28s} {n/2:5.0f} x {t*1e3:8.1f} us  {n*t/2/tot*100:5.1f}% of library time")
# the same step (forward + autograd backward of the module path, ~700 launches) captured in one CUDA graph
# (--no-graph skips it; no reference to an autograd graph built on the default stream may be alive during the capture)
graph_ms = float("nan")
loss_eager = loss.item()
try:
    if "--no-graph" in sys.argv:
        raise RuntimeError("skipped")
    params = [p for m in (enc, dec) for p in m.parameters()]
    loss_eager = loss.item()
    del loss          # no reference to an autograd graph made on the default stream may survive into the capture
    import gc
    gc.collect()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            for p in params:
                p.grad = None
            l, _, _ = training_step(enc, dec, p4, get_real="sum")
            l.backward()
            del l
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    for p in params:
        p.grad = None
    with torch.cuda.graph(g):
        static_loss, _, _ = training_step(enc, dec, p4, get_real="sum")
        static_loss.backward()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    graph_ms = a.elapsed_time(b) / steps
    print(f"CUDA-graph replay of the same step: {graph_ms:.2f} ms/step -> {B / graph_ms * 1e3:.0f} jets/s; loss {static_loss.item():.6g}")
except Exception as e:   # noqa: BLE001
    if "--no-graph" not in sys.argv:
        import traceback
        print("graph capture of the module path failed:", repr(e)[:300])
        print("".join(traceback.format_exc().splitlines(True)[-14:]))
try:
    fl = step_flops_per_jet(N, ENC, DEC)
except Exception:
    fl = float("nan")
print(f"cfg-4 B={B}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} jets/s; library kernels {tot:.2f} ms/step; loss {loss_eager:.6g}; "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; maxdim-2 flop model would give {fl/1e6:.1f} MFLOP/jet (not the maxdim-3 count)")
