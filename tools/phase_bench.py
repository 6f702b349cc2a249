"""Developer timing: per-phase GPU (events) and CPU (perf_counter) time of the raw fused path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_golden
from tests.test_fused_gpu import _plans
from lgn_autoencoder_b200 import fused, _lib
from oracle.lgae_oracle import synthetic_jets

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
g = load_golden("cfg1_b3")
enc, dec = _plans(g)
th_e, th_d = enc.flatten(g["enc_state"], dev), dec.flatten(g["dec_state"], dev)
p4 = synthetic_jets(B, 30, seed=1)["p4"]
p4 = (p4 / p4.abs().amax(dim=(1, 2), keepdim=True)).to(dev).contiguous()
names = ["enc_fwd", "dec_fwd", "chamfer", "dec_bwd", "enc_bwd"]
gpu = {n: 0.0 for n in names}; cpu = {n: 0.0 for n in names}
def phase(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.record(); out = fn(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    gpu[name] += a.elapsed_time(b); cpu[name] += (t1 - t0) * 1e3
    return out
steps = 10
for it in range(3 + steps):
    if it == 3:
        gpu = {n: 0.0 for n in names}; cpu = {n: 0.0 for n in names}
    lat00, lat11, ws_e, sel = phase("enc_fwd", lambda: fused.encoder_forward_raw(enc, th_e, p4, None))
    recon, _, ws_d = phase("dec_fwd", lambda: fused.decoder_forward_raw(dec, th_d, lat11))
    def ch():
        rg = recon.clone().requires_grad_(True); loss = fused.chamfer_loss(rg, p4, "sum"); loss.backward(); return rg.grad
    gr = phase("chamfer", ch)
    g_lat11, gd = phase("dec_bwd", lambda: fused.decoder_backward_raw(dec, th_d, lat11, ws_d, gr, None))
    ge = phase("enc_bwd", lambda: fused.encoder_backward_raw(enc, th_e, p4, None, ws_e, sel, None, g_lat11))
print("B", B, {n: f"gpu {gpu[n]/steps:.3f} ms cpu {cpu[n]/steps:.3f} ms" for n in names}, "total gpu", sum(gpu.values())/steps)
