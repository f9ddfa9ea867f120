"""Small runs for compute-sanitizer --tool initcheck (which kernels read memory the tool saw no write to):
    python profiles/initcheck_probe.py core | module"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
if sys.argv[1:] == ["core"]:
    from test_gpu_parity import _oracle_case, _run_core
    o = _oracle_case(1, 64, 12, 20, 48, 80, 19, "blocky", seed=11, gumbel=False, dtype=torch.float32)
    _run_core(o, 19, dtype=torch.float32)
else:
    from pinthememory_b200 import synth
    from pinthememory_b200.memory import Memory_sup
    B, C, h, w, Hm, Wm = 1, 64, 12, 20, 48, 80
    torch.manual_seed(3)
    mem = Memory_sup(19, C, C, 0.8, 1.0, False).cuda()
    mem.fold_min_pixels = 0
    x = synth.make_features(B, C, h, w, seed=1, device="cuda").requires_grad_(True)
    lab = synth.make_labels(B, Hm, Wm, 19, "blocky", seed=2).cuda()
    G = synth.make_upstream_grad((B, C, h, w), seed=3, device="cuda")
    uq, _, _, rl, wl = mem(x, lab, True, False)
    ((uq.float() * G).sum() + 0.02 * rl + 0.4 * wl[0] + 0.2 * wl[1]).backward()
torch.cuda.synchronize()
print("done")
