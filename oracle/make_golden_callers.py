"""Generate tests/golden/callers/{init_prototypes,main_loss}.npz by running the REFERENCE's own code on CPU.

Container-only (needs /root/reference). Nothing is copied: ``Trainer.memory_initalize`` is extracted
from /root/reference/train.py with ``ast`` and executed with a stand-in ``self`` (train.py itself cannot
be imported offline: it pulls the datasets/tensorboard stack); ``Upsample`` and ``CrossEntropyLoss2d``
are executed from network/mynn.py and loss.py the same way.

    python -m oracle.make_golden_callers
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from pinthememory_b200 import synth

REF = os.environ.get("PINMEM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "callers")


def _extract(path, names):
    """Source segments of the named top-level functions / classes / methods of a reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            found[node.name] = ast.get_source_segment(src, node)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError("not found in %s: %s" % (path, missing))
    return found


def reference_memory_initalize(batches, K, C):
    import textwrap

    body = textwrap.dedent(_extract(os.path.join(REF, "train.py"), ["memory_initalize"])["memory_initalize"])
    ns = {"torch": torch, "F": F, "tqdm": lambda it, **k: it, "enumerate": enumerate}
    exec(compile(body, "reference:train.py:memory_initalize", "exec"), ns)

    class Net:
        def __init__(self):
            self.module = types.SimpleNamespace(memory=types.SimpleNamespace(m_items=torch.zeros(K, C)))
            self._feat = {}

        def eval(self):
            pass

        def train(self):
            pass

        def __call__(self, input, gts=None, aux_gts=None):
            return [self._feat[int(input.reshape(-1)[0].item())]]

    net = Net()
    loader = []
    for i, (feat, lab) in enumerate(batches):
        B, Hm, Wm = lab.shape
        img = torch.full((B, 3, Hm, Wm), float(i))  # the stand-in network looks the features up by image id
        net._feat[i] = feat
        loader.append((img, lab.clone(), None, lab.clone()))
    fake_self = types.SimpleNamespace(net=net, args=types.SimpleNamespace(mem_slot=K, test_mode=False), train_loader=loader)
    t_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        ns["memory_initalize"](fake_self)
    finally:
        torch.Tensor.cuda = t_cuda
    return net.module.memory.m_items


def reference_main_loss(logits, labels):
    ns = {"nn": nn, "torch": torch, "F": F, "logging": types.SimpleNamespace(info=lambda *a, **k: None)}
    exec(_extract(os.path.join(REF, "network", "mynn.py"), ["Upsample"])["Upsample"], ns)
    exec(_extract(os.path.join(REF, "loss.py"), ["CrossEntropyLoss2d"])["CrossEntropyLoss2d"], ns)
    logits = logits.clone().requires_grad_(True)
    main_out = ns["Upsample"](logits, labels.shape[-2:])       # deepv3plus.py:575
    loss = ns["CrossEntropyLoss2d"](ignore_index=255)(main_out, labels)   # deepv3plus.py:578, loss.py:46
    loss.backward()
    return loss.detach(), logits.grad


def reference_input2basket(features, labels, K):
    """tsnelib.py:48-74 executed from the reference source on a stand-in ``self``; returns (vectors, class ids)."""
    import textwrap

    body = textwrap.dedent(_extract(os.path.join(REF, "tsnelib.py"), ["input2basket"])["input2basket"])
    ns = {"torch": torch, "F": F}
    exec(compile(body, "reference:tsnelib.py:input2basket", "exec"), ns)
    fake = types.SimpleNamespace(num_class=K, selected_clsid=list(range(K)), name2domId={"d": 0},
                                 feat_vecs=torch.tensor([]), feat_vec_labels=torch.tensor([]),
                                 feat_vec_domlabels=torch.tensor([]))
    t_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        ns["input2basket"](fake, features, labels, "d")
    finally:
        torch.Tensor.cuda = t_cuda
    return fake.feat_vecs, fake.feat_vec_labels.reshape(-1).long()


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(304)
    K, C = 19, 64
    # row 2: two "epochs" over three batches of different sizes (the reference loops the loader twice)
    batches = []
    for i, (B, h, w, Hm, Wm, kind) in enumerate([(2, 12, 16, 48, 64, "blocky"), (1, 12, 16, 48, 64, "iid"),
                                                 (2, 12, 16, 48, 64, "blocky")]):
        batches.append((synth.make_features(B, C, h, w, seed=50 + i), synth.make_labels(B, Hm, Wm, K, kind, seed=60 + i)))
    batches[2][1][batches[2][1] == 3] = 255  # class 3 rarer; class 7 absent everywhere -> count 0 -> "1 for nan" path
    for _, lab in batches:
        lab[lab == 7] = 255
    m = reference_memory_initalize(batches, K, C)
    arrays = {"memory": m.numpy(), "meta": np.array(repr({"K": K, "C": C, "n_batches": len(batches), "epochs": 2,
                                                           "torch": torch.__version__}))}
    for i, (f, l) in enumerate(batches):
        arrays["features%d" % i] = f.numpy()
        arrays["labels%d" % i] = l.numpy()
    np.savez_compressed(os.path.join(OUT, "init_prototypes.npz"), **arrays)

    # row 5: main loss at OS4 of a 96x128 crop, ragged size, with ignore pixels
    cases = {}
    for name, (B, h, w, Hm, Wm, kind) in {"os4": (2, 24, 32, 96, 128, "blocky"), "ragged": (1, 13, 17, 50, 67, "iid")}.items():
        logits = torch.randn(B, K, h, w) * 3.0
        labels = synth.make_labels(B, Hm, Wm, K, kind, seed=71)
        loss, grad = reference_main_loss(logits, labels)
        cases[name + ".logits"], cases[name + ".labels"] = logits.numpy(), labels.numpy()
        cases[name + ".loss"], cases[name + ".grad"] = loss.numpy(), grad.numpy()
    cases["meta"] = np.array(repr({"K": K, "cases": ["os4", "ragged"], "torch": torch.__version__}))
    np.savez_compressed(os.path.join(OUT, "main_loss.npz"), **cases)
    # row 3: t-SNE class-mean vectors (one class absent, an ignore band, non-integer up-sampling ratio)
    # (the reference function only works for batch size 1: it flattens the labels of the whole batch into one image,
    # tsnelib.py:57, and is called from the batch-1 evaluation loop)
    feat = synth.make_features(1, C, 12, 16, seed=81)
    lab = synth.make_labels(1, 45, 61, K, "blocky", seed=82, block=8)
    lab[lab == 5] = 255
    vecs, ids = reference_input2basket(feat, lab, K)
    np.savez_compressed(os.path.join(OUT, "tsne_basket.npz"), features=feat.numpy(), labels=lab.numpy(), vectors=vecs.numpy(),
                        class_ids=ids.numpy(), meta=np.array(repr({"K": K, "C": C, "torch": torch.__version__})))
    print("wrote init_prototypes.npz, main_loss.npz, tsne_basket.npz")


if __name__ == "__main__":
    main()
