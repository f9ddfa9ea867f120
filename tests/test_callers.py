"""Callers either side of the path: oracle pinned by reference-generated fixtures
(CPU), product against both (GPU)."""
import ast
import os

import numpy as np
import pytest
import torch

from golden_util import assert_close
from oracle import callers_oracle as co

GOLD = os.path.join(os.path.dirname(__file__), "golden", "callers")


def _load(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}


def _proto_batches(meta, fx, device="cpu"):
    one = [(fx["features%d" % i].to(device), fx["labels%d" % i].to(device)) for i in range(meta["n_batches"])]
    return one * meta["epochs"]  # the reference walks the loader twice (train.py:1006)


def test_oracle_prototypes_match_reference_fixture():
    meta, fx = _load("init_prototypes.npz")
    memory, sums, counts = co.initial_prototypes(_proto_batches(meta, fx), meta["K"])
    assert_close(memory, fx["memory"], 2e-6, "prototypes")
    assert counts[7] == 0 and torch.all(memory[7] == 0)  # class absent everywhere: zero row, no NaN


def test_oracle_main_loss_matches_reference_fixture():
    meta, fx = _load("main_loss.npz")
    for c in meta["cases"]:
        logits = fx[c + ".logits"].clone().requires_grad_(True)
        loss = co.upsampled_cross_entropy(logits, fx[c + ".labels"])
        loss.backward()
        assert_close(loss.detach(), fx[c + ".loss"], 2e-6, c + " loss")
        assert_close(logits.grad, fx[c + ".grad"], 2e-6, c + " grad")


@pytest.mark.gpu
def test_prototype_pool_matches_reference_fixture():
    from pinthememory_b200.callers import PrototypePool, initialize_memory
    from pinthememory_b200.memory import Memory_sup

    meta, fx = _load("init_prototypes.npz")
    batches = _proto_batches(meta, fx, "cuda")
    pool = PrototypePool(meta["K"], meta["C"])
    for f, l in batches:
        pool.add(f, l)
    memory = pool.finalize()
    assert_close(memory.cpu(), fx["memory"], 1e-5, "prototypes")
    _, sums, counts = co.initial_prototypes(_proto_batches(meta, fx), meta["K"])
    assert_close(pool.sums_counts[: meta["K"], : meta["C"]].cpu(), sums, 1e-5, "sums")
    assert_close(pool.sums_counts[: meta["K"], meta["C"]].cpu(), counts, 1e-6, "counts")
    assert torch.all(memory[7] == 0)
    mem = Memory_sup(meta["K"], meta["C"], meta["C"], 0.8, 1.0, False).cuda()
    initialize_memory(mem, batches)
    assert torch.equal(mem.m_items, memory) or torch.allclose(mem.m_items, memory, atol=1e-6)


@pytest.mark.gpu
def test_prototype_pool_bf16_and_full_size_counts():
    from pinthememory_b200 import synth
    from pinthememory_b200.callers import PrototypePool

    K, C = 19, 256
    f = synth.make_features(4, C, 96, 96, seed=3, device="cuda")
    lab = synth.make_labels(4, 96, 96, K, "blocky", seed=4).cuda()  # labels at feature resolution: integer counts
    p32, p16 = PrototypePool(K, C), PrototypePool(K, C)
    p32.add(f, lab).add(f, lab)
    p16.add(f.bfloat16(), lab).add(f.bfloat16(), lab)
    hist = torch.bincount(torch.where(lab == 255, torch.full_like(lab, K), lab).flatten(), minlength=K + 1).float() * 2
    assert torch.equal(p32.sums_counts[:, C].cpu(), hist.cpu())  # bit-exact counts
    assert torch.equal(p16.sums_counts[:, C].cpu(), hist.cpu())
    assert_close(p16.finalize(), p32.finalize(), 2e-2, "bf16 prototypes")


@pytest.mark.gpu
def test_upsampled_cross_entropy_matches_reference_fixture_and_oracle():
    from pinthememory_b200 import synth
    from pinthememory_b200.callers import upsampled_cross_entropy

    meta, fx = _load("main_loss.npz")
    for c in meta["cases"]:
        logits = fx[c + ".logits"].cuda().requires_grad_(True)
        loss = upsampled_cross_entropy(logits, fx[c + ".labels"].cuda())
        (loss * 3.0).backward()
        assert_close(loss.detach().cpu(), fx[c + ".loss"], 1e-5, c + " loss")
        assert_close(logits.grad.cpu() / 3.0, fx[c + ".grad"], 1e-5, c + " grad")
    # full size (decoder output at OS4 of a 768x768 crop), against the eager torch ops on the same GPU
    K = 19
    logits = (torch.randn(2, K, 192, 192, device="cuda") * 4).requires_grad_(True)
    labels = synth.make_labels(2, 768, 768, K, "blocky", seed=9).cuda()
    loss = upsampled_cross_entropy(logits, labels)
    loss.backward()
    ref_in = logits.detach().double().requires_grad_(True)  # fp64: eager fp32 atomics carry their own 5e-6 of noise
    ref = co.upsampled_cross_entropy(ref_in, labels)
    ref.backward()
    assert_close(loss.detach().double(), ref.detach(), 1e-5, "loss 768")
    assert_close(logits.grad.double(), ref_in.grad, 1e-5, "grad 768")
    # all-ignore -> NaN like torch
    nan = upsampled_cross_entropy(logits.detach(), torch.full_like(labels, 255))
    assert torch.isnan(nan)


def test_oracle_class_means_match_reference_fixture():
    """tsnelib.py:48-74 executed from the reference source (oracle/make_golden_callers.py) vs the oracle restatement."""
    meta, fx = _load("tsne_basket.npz")
    means, counts = co.class_mean_vectors(fx["features"], fx["labels"], meta["K"])
    present = (counts != 0).nonzero().flatten()
    assert torch.equal(present, fx["class_ids"]), "the reference appends exactly the classes that have pixels"
    assert_close(means[present], fx["vectors"], 2e-6, "class-mean vectors")
    assert int(counts[5]) == 0 and torch.all(means[5] == 0)


@pytest.mark.gpu
def test_class_mean_vectors_match_reference_fixture_and_oracle():
    from pinthememory_b200.callers import class_mean_vectors
    from pinthememory_b200 import synth

    meta, fx = _load("tsne_basket.npz")
    means, counts = class_mean_vectors(fx["features"].cuda(), fx["labels"].cuda(), meta["K"])
    present = (counts != 0).nonzero().flatten().cpu()
    assert torch.equal(present, fx["class_ids"])
    assert_close(means.cpu()[present], fx["vectors"], 1e-5, "class-mean vectors vs the reference")
    lab = fx["labels"].reshape(-1).clone()
    lab[lab == 255] = meta["K"]
    assert torch.equal(counts.cpu().long(), torch.bincount(lab, minlength=meta["K"] + 1)[: meta["K"]]), "counts are exact"
    # a batch of full-size maps against the oracle on the same device (the reference itself only takes batch 1)
    f = synth.make_features(2, 256, 48, 48, seed=5, device="cuda")
    l = synth.make_labels(2, 384, 384, 19, "blocky", seed=6, device="cuda")
    m1, c1 = class_mean_vectors(f, l, 19)
    m2, c2 = co.class_mean_vectors(f, l, 19)
    assert_close(m1, m2, 1e-5, "batch of maps vs oracle")
    assert torch.equal(c1, c2)
