"""Diagnostic: conv_bn_act backward vs torch modules; locates wrong elements and checks intermediates."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi
from pinthememory_b200.memory import conv_bn_act

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
captured = []
orig_fwd = capi.conv1x1_fwd


def spy(x, hi, lo, M, y=None, stats=None, accumulate=False):
    pre = None if y is None else y.clone()
    out = orig_fwd(x, hi, lo, M, y=y, stats=stats, accumulate=accumulate)
    captured.append((x.clone(), hi.clone(), lo.clone(), M, pre, out.clone(), accumulate))
    return out


capi.conv1x1_fwd = spy


def where(bad, hw):
    idx = bad.nonzero()
    if idx.numel() == 0:
        return "none"
    b, m, p = idx[:, 0], idx[:, 1], idx[:, 2]
    return "n=%d images %s rowblk32 %s pxblk32 %s" % (idx.shape[0], sorted(set(b.tolist())), sorted(set((m // 32).tolist()))[:12],
                                                     sorted(set((p // 32).tolist()))[:24])


def run(shape, residual, training):
    B, C, h, w = shape
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(C, C, 1, bias=False).cuda()
    bn = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
    conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
    for m in (bn, bn_r):
        m.train(training)
    x = torch.randn(B, C, h, w, device="cuda")
    G = torch.randn(B, C, h, w, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    captured.clear()
    ya = conv_bn_act(conv, bn, xa, xa if residual else None, True)
    cr = conv_r(xb)
    cr.retain_grad()
    t = bn_r(cr)
    yb = torch.relu(xb + t if residual else t)
    (ya * G).sum().backward()
    (yb * G).sum().backward()
    hw = h * w
    print(f"shape {shape} residual={residual} training={training}")
    e = (xa.grad - xb.grad).abs().view(B, C, hw)
    print("  dx bad:", where(e > 1e-4 * xb.grad.abs().max(), hw))
    # the dgrad call: captured[-1]
    xin, hi, lo, M, pre, out, acc = captured[-1]
    e2 = (xin - cr.grad).abs().view(B, C, hw)
    print("  dxc (input of dgrad) vs torch d(conv out) bad:", where(e2 > 1e-4 * cr.grad.abs().max(), hw), "acc", acc)
    Wt = (hi + lo)[:M].double()
    ref = torch.einsum("mk,bkp->bmp", Wt, xin.double().view(B, C, hw))
    if acc:
        ref = ref + pre.double().view(B, M, hw)
    e3 = (out.double().view(B, M, hw) - ref).abs()
    print("  dgrad GEMM vs fp64 of its own inputs bad:", where(e3 > 1e-4 * ref.abs().max(), hw))
    dWe = (conv.weight.grad - conv_r.weight.grad).abs()
    print("  dW max err / scale", float(dWe.max() / conv_r.weight.grad.abs().max()))


run((8, 256, 48, 48), False, False)
run((8, 256, 48, 48), False, True)
run((8, 256, 48, 48), True, True)
run((8, 256, 48, 48), True, False)
run((4, 64, 12, 20), False, True)
