"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference module (container only).

TEST INFRASTRUCTURE. Run in the build container, where ``/root/reference`` is mounted:

    python -m oracle.make_golden

Each fixture holds the inputs, the module state (state_dict + m_items) and what the
reference's ``Memory_sup`` produced for them on CPU fp32 (torch version recorded in the
file): the five forward outputs, the tensor entering the 1x1 output conv (``u``) and the
write feature (``f``) captured with forward hooks, the new ``m_items``, and -- for the
training cases -- the gradients of the scalar
``<G, updated_query> + 0.02*readloss + 0.4*div + 0.2*cls`` (loss weights train.py:1213-1215)
w.r.t. the query, ``m_items`` (when it carries grad) and the 8 parameters.
The reference hard-codes 19 slots (memory.py:336), so every case uses K=19.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.ref_loader import build_reference_memory  # noqa: E402
from pinthememory_b200 import synth  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
K = 19
LOSS_WEIGHTS = dict(read=0.02, div=0.4, cls=0.2)

CASES = [
    # name, B, C, h, w, Hm, Wm, label kind, dict(options)
    ("train_write_c64_blocky", 2, 64, 12, 12, 48, 48, "blocky", dict(writing=True, detach=False, train=True, backward=True)),
    ("train_write_c64_iid", 2, 64, 12, 12, 48, 48, "iid", dict(writing=True, detach=False, train=True, backward=True)),
    ("train_write_c256_ragged", 1, 256, 7, 9, 29, 41, "iid", dict(writing=True, detach=False, train=True, backward=True)),
    ("metatest_read_dM_c64", 2, 64, 12, 12, 48, 48, "blocky", dict(writing=False, detach=True, train=True, backward=True, mem_grad=True)),
    ("final_update_eval_c64", 2, 64, 12, 12, 48, 48, "blocky", dict(writing=True, detach=True, train=False, backward=False)),
    ("eval_read_nomask_c64", 1, 64, 10, 20, 0, 0, None, dict(writing=False, detach=True, train=False, backward=False)),
    ("labels_at_feature_res_c64", 2, 64, 12, 12, 12, 12, "iid", dict(writing=True, detach=False, train=True, backward=True)),
    ("absent_and_all_ignore_c64", 2, 64, 12, 12, 48, 48, "blocky", dict(writing=True, detach=False, train=True, backward=True, all_ignore_image=True)),
    ("gumbel_train_c64", 2, 64, 12, 12, 48, 48, "blocky", dict(writing=True, detach=False, train=True, backward=True, gumbel=True)),
    ("temperature_momentum_c64", 2, 64, 12, 12, 48, 48, "iid", dict(writing=True, detach=False, train=True, backward=True, temperature=0.5, momentum=0.3)),
]


def _np(t):
    return t.detach().cpu().numpy()


def run_case(name, B, C, h, w, Hm, Wm, kind, opt, seed):
    torch.manual_seed(seed)
    gumbel = bool(opt.get("gumbel"))
    mem = build_reference_memory(K, C, opt.get("momentum", 0.8), opt.get("temperature", 1.0), gumbel)
    # BN / classifier away from their degenerate init so every gradient path is exercised
    with torch.no_grad():
        for p in mem.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        mem.clsfier.weight.normal_(0.0, 0.2)
    mem.train(opt["train"])
    if not opt["train"]:
        for m in mem.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0.0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    state = {k: _np(v) for k, v in mem.state_dict().items()}
    m0 = mem.m_items.clone()
    if opt.get("mem_grad"):
        mem.m_items = m0.clone().requires_grad_(True)

    x = synth.make_features(B, C, h, w, seed=seed + 10).requires_grad_(opt["backward"])
    labels = None
    if kind is not None:
        labels = synth.make_labels(B, Hm, Wm, K, kind, seed=seed + 11, block=8)
        if opt.get("all_ignore_image"):
            labels[0] = 255
    G = synth.make_upstream_grad((B, C, h, w), seed=seed + 12)

    captured = {}
    hooks = [mem.output.register_forward_hook(lambda m, i, o: captured.__setitem__("u", i[0].detach().clone())),
             mem.writenet.register_forward_hook(lambda m, i, o: captured.__setitem__("f", o.detach().clone()))]

    fx = dict(x=_np(x), m_items_in=_np(m0), G=_np(G))
    if labels is not None:
        fx["labels"] = _np(labels).astype(np.int64)
    if gumbel:
        # replay the two draws the reference is about to make (memory.py:183-184)
        rng = torch.get_rng_state()
        probe = torch.empty(B * h * w, K)
        fx["g_query"] = _np(-torch.empty_like(probe).exponential_().log())
        fx["g_memory"] = _np(-torch.empty_like(probe).exponential_().log())
        torch.set_rng_state(rng)

    mem_in = mem.m_items
    uq, sq, sm, readloss, writeloss = mem(x, labels, opt["writing"], opt["detach"])
    for hk in hooks:
        hk.remove()
    fx.update(updated_query=_np(uq), score_query=_np(sq), score_memory=_np(sm), u=_np(captured["u"]),
              m_items_out=_np(mem.m_items))
    fx["readloss"] = np.float32(float(torch.as_tensor(readloss).detach()))
    fx["div_loss"] = np.float32(float(torch.as_tensor(writeloss[0]).detach()))
    fx["cls_loss"] = np.float32(float(torch.as_tensor(writeloss[1]).detach()))
    if "f" in captured:
        fx["f"] = _np(captured["f"])

    if opt["backward"]:
        total = (uq * G).sum()
        if labels is not None:
            total = total + LOSS_WEIGHTS["read"] * readloss
        if opt["writing"]:
            total = total + LOSS_WEIGHTS["div"] * writeloss[0] + LOSS_WEIGHTS["cls"] * writeloss[1]
        total.backward()
        fx["grad_x"] = _np(x.grad)
        if opt.get("mem_grad"):
            fx["grad_m_items"] = _np(mem_in.grad)
        for n, p in mem.named_parameters():
            if p.grad is not None:
                fx["grad_param." + n] = _np(p.grad)

    for k, v in state.items():
        fx["state." + k] = v
    meta = dict(opt)
    meta.update(name=name, B=B, C=C, h=h, w=w, Hm=Hm, Wm=Wm, K=K, kind=kind, torch=torch.__version__,
                momentum=mem.momentum, temperature=mem.temperature, loss_weights=LOSS_WEIGHTS)
    fx["meta"] = np.array(repr(meta))
    return fx


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for i, (name, B, C, h, w, Hm, Wm, kind, opt) in enumerate(CASES):
        fx = run_case(name, B, C, h, w, Hm, Wm, kind, opt, seed=1000 + 17 * i)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **fx)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  readloss={fx['readloss']:.6f} "
              f"div={fx['div_loss']:.6f} cls={fx['cls_loss']:.6f}")


if __name__ == "__main__":
    main()
