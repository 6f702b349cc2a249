"""Developer: where the end-to-end step's time goes (wall clock vs events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CFG, build_models, synthetic_jets
from lgn_autoencoder_b200.train import FusedTrainStep
dev = torch.device("cuda:0")
enc, dec = build_models(dev)
st = FusedTrainStep(enc, dec, 512, l1_lambda=1e-8, use_graph=True, get_real="sum")
hp = synthetic_jets(512, 30, seed=3).pin_memory()
st.host_p4.copy_(hp)
st.load(hp)
for _ in range(5): st.run(); st.step_host()
torch.cuda.synchronize()
def wall(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def run_sync(): st.run(); torch.cuda.synchronize()
print("graph replay back-to-back (async)  %.0f us" % wall(st.run))
print("graph replay + sync each step      %.0f us" % wall(run_sync))
print("step_host (in-graph H2D/D2H, sync) %.0f us" % wall(st.step_host))
def item_step(): return st.step(hp).item()
print("step(load)+item()                   %.0f us" % wall(item_step))
st2 = FusedTrainStep(enc, dec, 512, l1_lambda=1e-8, use_graph=False, get_real="sum")
st2.host_p4.copy_(hp)
for _ in range(3): st2.step_host()
print("step_host EAGER (C launches, sync)  %.0f us" % wall(st2.step_host))
def eager_async(): st2._launch()
print("eager launches back-to-back (async) %.0f us" % wall(eager_async))
# events inside a synced loop
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(20):
    a.record(); st.step_host(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
print("step_host event time               %.0f us" % (tot / 20 * 1e3))
# CPU cost of the replay call itself
torch.cuda.synchronize()
ts = []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st.graph.replay(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append(((t1 - t0) * 1e6, (t2 - t0) * 1e6))
ts.sort()
print("replay() call CPU time median %.0f us ; call+sync %.0f us" % (ts[15][0], sorted(t[1] for t in ts)[15]))
