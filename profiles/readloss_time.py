"""pack + read loss on the cfg-2 / cfg-4 shapes: CUDA-event time of pm_readloss_fwd8 and a checksum (A/B across env switches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi, synth

dev, K, KP = "cuda", 19, 20
for (B, h, w, Hm, Wm, kind) in [(8, 96, 96, 768, 768, "blocky"), (8, 96, 96, 768, 768, "iid"), (8, 48, 48, 768, 768, "blocky"), (8, 192, 192, 768, 768, "blocky")]:
    N = B * h * w
    torch.manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(N, 256, device=dev), dim=1)
    M = synth.make_memory(K, 256, device=dev)
    s = torch.zeros(N, KP, device=dev)
    s[:, :K] = q @ M.t()
    labels = synth.make_labels(B, Hm, Wm, K, kind, device=dev)
    ts = []
    for it in range(30):
        buf = torch.zeros(N * KP + 2 * capi.WS_WORDS + 4, dtype=torch.float32, device=dev)
        ds, ws, out = buf[: N * KP], buf[N * KP: N * KP + 2 * capi.WS_WORDS], buf[N * KP + 2 * capi.WS_WORDS:]
        lab8 = capi.labels_pack(labels, K, ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.readloss_fwd8(s, lab8, 1.0, B, h, w, K, ds, ws, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[5:])
    print("%dx%d %s: readloss %.1f us (min %.1f)  loss %.7f  |ds| %.6f  ds.sum %.6e" % (h, w, kind, ts[len(ts) // 2], ts[0], out[0].item(), ds.abs().sum().item(), ds.double().sum().item()))
