"""Developer: the optimizer step next to the training step ('optimizer step excluded and reported
separately'): two torch.optim.Adam (the reference's setup) vs FlatAdam, eager and as a node of the step's CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models, synthetic_jets
from lgn_autoencoder_b200.train import FlatAdam, FusedTrainStep

B = 512
dev = torch.device("cuda:0")
enc, dec = build_models(dev)
fs = FusedTrainStep(enc, dec, B, l1_lambda=1e-8, use_graph=True, get_real="sum")
fs.load(synthetic_jets(B, 30, seed=3))
fs.run()


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


t_step = timeit(fs.run)
opts = [torch.optim.Adam(enc.parameters(), 1e-3), torch.optim.Adam(dec.parameters(), 1e-3)]
t_torch = timeit(lambda: [o.step() for o in opts])
flat = FlatAdam(fs, lr=1e-3)
t_flat = timeit(flat.step)
t_step_torch = timeit(lambda: (fs.run(), [o.step() for o in opts]))
fs.attach_optimizer(flat)
t_step_flat = timeit(fs.run)
print(f"training step {t_step:.0f} us; 2 x torch.optim.Adam.step {t_torch:.0f} us; FlatAdam.step {t_flat:.1f} us; "
      f"step + torch Adam {t_step_torch:.0f} us ({B / t_step_torch * 1e6:.0f} jets/s); step with FlatAdam in the graph {t_step_flat:.0f} us "
      f"({B / t_step_flat * 1e6:.0f} jets/s)")
