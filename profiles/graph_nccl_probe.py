"""Where does CUDA-graph capture of the sharded step stop working? (run under torchrun, 2 ranks)"""
import faulthandler
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(50, exit=True)
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    print("[rank %d]" % rank, *a, flush=True)


t = torch.ones(1000, device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
say("eager all_reduce ok", t[0].item())

# stage 1: plain captured all_reduce on the main thread
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_reduce(t)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
t.fill_(1.0)
with torch.cuda.graph(g):
    dist.all_reduce(t)
g.replay()
torch.cuda.synchronize()
say("stage 1 captured all_reduce ok", t[0].item())

from pinthememory_b200 import sharding, synth
from pinthememory_b200.graphed import GraphedStep
from pinthememory_b200.memory import Memory_sup

torch.manual_seed(1)
mem = Memory_sup(19, 64, 64, 0.8, 1.0, False).to(dev)
for p in mem.parameters():
    dist.broadcast(p.data, 0)
sharding.broadcast_memory(mem)
sharding.enable_sharded_update(mem)
x = synth.make_features(2, 64, 12, 16, seed=3 + rank, device=dev)
lab = synth.make_labels(2, 48, 64, 19, "blocky", seed=5 + rank).to(dev)
G = synth.make_upstream_grad((2, 64, 12, 16), seed=9, device=dev)

step = GraphedStep(mem, x, lab, None, memory_writing=True)
step.replay()
torch.cuda.synchronize()
say("stage 2 forward-only graph with all_reduce ok", float(step.outputs["writeloss"][0]))
step.release()

mode = os.environ.get("PROBE_MODE", "global")
import pinthememory_b200.graphed as gr
if mode != "global":
    _orig = torch.cuda.graph
    gr.torch.cuda.graph = lambda g_, **k: _orig(g_, capture_error_mode=mode, **k)
step = GraphedStep(mem, x, lab, G, memory_writing=True, writing_detach=False)
step.replay()
torch.cuda.synchronize()
say("stage 3 forward+backward graph ok (mode %s)" % mode, float(step.outputs["writeloss"][0]), float(step.query_grad.abs().sum()))
step.release()
del g
torch.cuda.synchronize()
dist.barrier()
say("graphs released")
dist.destroy_process_group()
say("process group destroyed")
faulthandler.cancel_dump_traceback_later()
