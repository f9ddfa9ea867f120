"""Timing / agreement probe of the two read-loss kernels (pm_readloss_fwd vs pm_labels_pack + pm_readloss_fwd8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi, synth

dev = "cuda"
K, KP = 19, 20


def run(B, h, w, Hm, Wm, kind, v8, reps=20):
    N = B * h * w
    torch.manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(N, 256, device=dev), dim=1)
    M = synth.make_memory(K, 256, device=dev)
    s = torch.zeros(N, KP, device=dev)
    s[:, :K] = q @ M.t()
    labels = synth.make_labels(B, Hm, Wm, K, kind, device=dev)

    def once():
        buf = torch.zeros(N * KP + 2 * capi.WS_WORDS + 4, dtype=torch.float32, device=dev)
        ds, ws, out = buf[: N * KP], buf[N * KP: N * KP + 2 * capi.WS_WORDS], buf[N * KP + 2 * capi.WS_WORDS:]
        if v8:
            lab8 = capi.labels_pack(labels, K, ws)
            capi.readloss_fwd8(s, lab8, 1.0, B, h, w, K, ds, ws, out)
        else:
            capi.readloss_fwd(s, labels, 1.0, B, h, w, K, ds, ws, out)
        return ds, ws, out

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    capi.enable_kernel_timing(True)
    capi.reset_counters()
    for _ in range(reps):
        r = once()
    t = {k: sum(v) / len(v) * 1e3 for k, v in capi.kernel_timings_ms().items()}
    capi.enable_kernel_timing(False)
    return r, t


for shape in [(8, 96, 96, 768, 768), (8, 48, 48, 768, 768), (2, 48, 48, 768, 768), (8, 192, 192, 768, 768)]:
    for kind in ("blocky", "iid"):
        (d1, w1, o1), t1 = run(*shape, kind, False)
        (d8, w8, o8), t8 = run(*shape, kind, True)
        rel = float((d8 - d1).norm() / d1.norm())
        print(shape, kind, "v1 us", {k: round(v, 1) for k, v in t1.items()}, "v8 us", {k: round(v, 1) for k, v in t8.items()},
              "loss", float(o1[0]), float(o8[0]), "ds rel", "%.2e" % rel, flush=True)
