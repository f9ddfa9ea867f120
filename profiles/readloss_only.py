"""Runs pack + readloss8 a few times on the cfg-2 shape (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi, synth

dev, K, KP = "cuda", 19, 20
B, h, w, Hm, Wm = 8, 96, 96, 768, 768
N = B * h * w
q = torch.nn.functional.normalize(torch.randn(N, 256, device=dev), dim=1)
M = synth.make_memory(K, 256, device=dev)
s = torch.zeros(N, KP, device=dev)
s[:, :K] = q @ M.t()
labels = synth.make_labels(B, Hm, Wm, K, "blocky", device=dev)
for _ in range(4):
    buf = torch.zeros(N * KP + 2 * capi.WS_WORDS + 4, dtype=torch.float32, device=dev)
    ds, ws, out = buf[: N * KP], buf[N * KP: N * KP + 2 * capi.WS_WORDS], buf[N * KP + 2 * capi.WS_WORDS:]
    lab8 = capi.labels_pack(labels, K, ws)
    capi.readloss_fwd8(s, lab8, 1.0, B, h, w, K, ds, ws, out)
torch.cuda.synchronize()
