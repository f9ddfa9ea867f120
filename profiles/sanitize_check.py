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
