"""Small all-kernel run for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_check.py
"""
import sys, torch
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from test_gpu_parity import _oracle_case, _run_core
from golden_util import assert_close
for case in [(1, 256, 32, 32, 128, 128, 19, "blocky", False), (2, 64, 12, 20, 48, 80, 19, "iid", True), (1, 64, 7, 9, 29, 41, 19, "iid", False), (1, 256, 16, 16, 256, 256, 19, "blocky", False)]:
    B, C, h, w, Hm, Wm, K, kind, g = case
    for dt in (torch.float32, torch.bfloat16):
        o = _oracle_case(B, C, h, w, Hm, Wm, K, kind, seed=11, gumbel=g, dtype=dt)
        r = _run_core(o, K, dtype=dt)
        torch.cuda.synchronize()
        assert_close(r["dx"], o["dx"], 1e-5 if dt == torch.float32 else 2e-2, "dx")
        print("ok", case, dt)

# the whole module (tcgen05 GEMMs, BatchNorm passes, fused gradient sum, column-softmax combine in the read kernel)
from pinthememory_b200 import synth
from pinthememory_b200.memory import Memory_sup

for (B, C, h, w, Hm, Wm) in [(2, 256, 16, 16, 64, 64), (1, 64, 12, 20, 48, 80), (2, 128, 8, 24, 64, 96)]:
    for dt in (torch.float32, torch.bfloat16):
        torch.manual_seed(3)
        mem = Memory_sup(19, C, C, 0.8, 1.0, False).cuda()
        mem.fold_min_pixels = 0
        mem.overlap_write = (B == 2)   # the two-stream mode too (label pass / column softmax on side streams)
        x = synth.make_features(B, C, h, w, seed=1, device="cuda").requires_grad_(True)
        lab = synth.make_labels(B, Hm, Wm, 19, "blocky", seed=2).cuda()
        G = synth.make_upstream_grad((B, C, h, w), seed=3, device="cuda")
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
            uq, _, _, rl, wl = mem(x, lab, True, False)
        ((uq.float() * G).sum() + 0.02 * rl + 0.4 * wl[0] + 0.2 * wl[1]).backward()
        mem.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
            mem(x.detach(), None, False)
        torch.cuda.synchronize()
        assert torch.isfinite(x.grad).all() and torch.isfinite(mem.m_items).all()
        print("ok module", (B, C, h, w), dt)
