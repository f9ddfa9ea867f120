import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from golden_util import rel_l2, max_abs_over_scale
from pinthememory_b200 import synth
from pinthememory_b200.memory import Memory_sup

torch.backends.cudnn.allow_tf32 = False
B, C, h, w, Hm, Wm, K = 2, 64, 12, 16, 48, 64, 19
x0 = synth.make_features(B, C, h, w, seed=21, device="cuda").bfloat16().float()
lab = synth.make_labels(B, Hm, Wm, K, "blocky", seed=22).cuda()
G = synth.make_upstream_grad((B, C, h, w), seed=23, device="cuda")


def run(fold, dtype, writing=True):
    torch.manual_seed(5)
    mem = Memory_sup(K, C, C, 0.8, 1.0, False).cuda()
    with torch.no_grad():
        mem.clsfier.weight.normal_(0, 0.2)
    mem.fold_memory_into_conv = fold
    mem.fold_min_pixels = 0
    x = x0.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        uq, sq, sm, rl, wl = mem(x.to(dtype), lab, writing, False)
    torch.autograd.backward([uq, rl, wl[0], wl[1]], [G.to(uq.dtype)] + [torch.tensor(v, device="cuda") for v in (0.02, 0.4, 0.2)])
    return dict(uq=uq.detach().float(), dx=x.grad, **{n: p.grad for n, p in mem.named_parameters()})


ref = run(False, torch.float32)
for name, r in (("fp32 fold", run(True, torch.float32)), ("bf16 plain", run(False, torch.bfloat16)), ("bf16 fold", run(True, torch.bfloat16))):
    print(name, {k: "%.1e/%.1e" % (rel_l2(r[k], ref[k]), max_abs_over_scale(r[k], ref[k])) for k in ref})
