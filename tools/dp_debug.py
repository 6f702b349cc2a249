"""Developer: stage-by-stage check of the 2-rank FusedTrainStep (prints after every stage)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import CFG, build_models, synthetic_jets
from lgn_autoencoder_b200.train import FusedTrainStep
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[r{rank} {time.time() % 1000:.2f}]", *a, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
say("pg up")
enc, dec = build_models(dev)
st = FusedTrainStep(enc, dec, 512, l1_lambda=1e-8, l1_scale=1.0 / world, use_graph=(os.environ.get("GRAPH", "1") == "1"), get_real="sum")
st.load(synthetic_jets(512, 30, seed=100 + rank))
say("built")
st._launch(); torch.cuda.synchronize(); say("eager 1", st.loss.item())
st._launch(); torch.cuda.synchronize(); say("eager 2", st.loss.item())
g0 = st.g_e.clone()
t = st.g_e.double().sum().clone(); dist.all_reduce(t); say("grad checksum", t.item())
st.run(); torch.cuda.synchronize(); say("run 1 (capture)", st.loss.item())
for _ in range(5): st.run()
torch.cuda.synchronize(); say("replays ok", (st.g_e - g0).abs().max().item())
dist.barrier(); say("barrier ok")
import threading
threading.Timer(15.0, lambda: (say("destroy hung -> hard exit"), os._exit(0))).start()
st.graph = None
import gc; gc.collect(); torch.cuda.synchronize(); say("graph released")
dist.destroy_process_group(); say("destroyed")
os._exit(0)
