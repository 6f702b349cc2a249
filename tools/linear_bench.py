"""Developer: time of the generic Linear entry points at the cfg-4 MLP shapes (30720 rows, widths 16 / 96)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lgn_autoencoder_b200 import layer_ops
dev = torch.device("cuda:0")
rows = 30720
for nin, nout in ((96, 96), (16, 96), (96, 16), (72, 72), (12, 72)):
    x = torch.randn(rows, nin, dtype=torch.float64, device=dev, requires_grad=True)
    w = torch.randn(nout, nin, dtype=torch.float64, device=dev, requires_grad=True)
    b = torch.randn(nout, dtype=torch.float64, device=dev, requires_grad=True)
    g = torch.randn(rows, nout, dtype=torch.float64, device=dev)
    def run():
        y = layer_ops.linear(x, w, b, leaky_slope=0.01)
        return torch.autograd.grad(y, (x, w, b), g)
    for _ in range(3):
        run()
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.linear(x, w, b), 0.01)
    gref = torch.autograd.grad(ref, (x, w, b), g)
    got = run()
    err = max(((a - r).abs().max() / r.abs().max()).item() for a, r in zip(got, gref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 3 * 2.0 * rows * nin * nout
    print(f"linear {nin:3d}->{nout:3d}: fwd+bwd {us:7.1f} us  {fl / us * 1e-6:6.2f} TFLOP/s  grad err vs torch {err:.1e}")
    from lgn_autoencoder_b200 import _lib
    k = _lib.kernel_timings(run, reps=5)
    print("   ", "  ".join(f"{n} {t*1e3:.1f}us" for n, (c, t) in k.items()))
