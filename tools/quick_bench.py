"""Developer timing of the raw fused path (not the bench contract): cfg-1 shapes, random-init weights."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_golden
from tests.test_fused_gpu import _plans
from lgn_autoencoder_b200 import fused, _lib
from oracle.lgae_oracle import synthetic_jets

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
g = load_golden("cfg1_b3")
enc, dec = _plans(g)
th_e, th_d = enc.flatten(g["enc_state"], dev), dec.flatten(g["dec_state"], dev)
p4 = synthetic_jets(B, 30, seed=1)["p4"]
p4 = (p4 / p4.abs().amax(dim=(1, 2), keepdim=True)).to(dev).contiguous()

def step():
    lat00, lat11, ws_e, sel = fused.encoder_forward_raw(enc, th_e, p4, None)
    recon, _, ws_d = fused.decoder_forward_raw(dec, th_d, lat11)
    rg = recon.clone().requires_grad_(True)
    loss = fused.chamfer_loss(rg, p4, "sum")
    loss.backward()
    g_lat11, gd = fused.decoder_backward_raw(dec, th_d, lat11, ws_d, rg.grad, None)
    ge = fused.encoder_backward_raw(enc, th_e, p4, None, ws_e, sel, None, g_lat11)
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"B={B} {ms:.3f} ms/step  {B / ms * 1e3:.0f} jets/s  wall {(time.time() - t0) / steps * 1e3:.3f} ms  launches/step {(_lib.launch_count() - n0) / steps}  loss {loss.item():.6f}")
