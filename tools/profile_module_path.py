"""Developer: where the host time of the module-API training step goes (cProfile of 50 eager steps at cfg-1, bs 512)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import CFG, build_models, synthetic_jets
from lgn_autoencoder_b200.train import training_step

dev = torch.device("cuda:0")
enc, dec = build_models(dev)
p4 = synthetic_jets(512, CFG["n"], seed=1).to(dev)
params = [p for m in (enc, dec) for p in m.parameters()]


def step():
    for p in params:
        p.grad = None
    loss, _, _ = training_step(enc, dec, p4, l1_lambda=1e-8, get_real="sum")
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
