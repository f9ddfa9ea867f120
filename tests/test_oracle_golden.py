"""Pin the oracle: restatement vs. the committed outputs of the unmodified reference.

CPU-only. ``tests/golden/*.npz`` were produced by ``oracle/make_golden.py`` from
``/root/reference/network/memory.py`` (CPU fp32); the oracle must reproduce them, and --
when the reference tree is present (build container) -- the live module as well.
"""
import pytest
import torch

from oracle import memory_oracle as mo
from oracle.ref_loader import build_reference_memory, reference_available
from golden_util import LOSS_WEIGHTS, assert_close, golden_names, load_golden, load_state

TOL = 2e-6  # same ops on the same CPU, only association order may differ


def _run_oracle(meta, fx, dtype=torch.float32):
    mem = mo.OracleMemorySup(meta["K"], meta["C"], meta["C"], meta["momentum"], meta["temperature"],
                             bool(meta.get("gumbel")))
    load_state(mem, fx)
    mem.train(meta["train"])
    if dtype == torch.float64:
        mem.double()
        mem.m_items = mem.m_items.double()
    if meta.get("mem_grad"):
        mem.m_items = mem.m_items.clone().requires_grad_(True)
    x = fx["x"].to(dtype).clone().requires_grad_(meta["backward"])
    labels = fx.get("labels")
    noise = (fx["g_query"].to(dtype), fx["g_memory"].to(dtype)) if meta.get("gumbel") else None
    mem_in = mem.m_items
    uq, sq, sm, rl, wl = mem(x, labels, meta["writing"], meta["detach"], noise=noise)
    out = dict(updated_query=uq, score_query=sq, score_memory=sm, m_items_out=mem.m_items,
               readloss=torch.as_tensor(rl, dtype=dtype), div_loss=torch.as_tensor(wl[0], dtype=dtype),
               cls_loss=torch.as_tensor(wl[1], dtype=dtype))
    if meta["backward"]:
        total = (uq * fx["G"].to(dtype)).sum()
        if labels is not None:
            total = total + LOSS_WEIGHTS["read"] * rl
        if meta["writing"]:
            total = total + LOSS_WEIGHTS["div"] * wl[0] + LOSS_WEIGHTS["cls"] * wl[1]
        total.backward()
        out["grad_x"] = x.grad
        if meta.get("mem_grad"):
            out["grad_m_items"] = mem_in.grad
        for n, p in mem.named_parameters():
            if p.grad is not None:
                out["grad_param." + n] = p.grad
    return out


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_fixture(name):
    meta, fx = load_golden(name)
    out = _run_oracle(meta, fx)
    checked = 0
    for key, got in out.items():
        if key not in fx:
            continue
        assert_close(got.detach().float(), fx[key].float(), TOL, f"{name}:{key}")
        checked += 1
    assert checked >= 5
    grads = [k for k in fx if k.startswith("grad_")]
    for k in grads:
        assert k in out, f"{name}: oracle produced no {k}"


@pytest.mark.parametrize("name", ["train_write_c64_blocky", "metatest_read_dM_c64"])
def test_oracle_fp64_agrees_with_fp32_reference(name):
    """The fp64 run of the oracle (used for error attribution) stays within fp32 noise."""
    meta, fx = load_golden(name)
    out = _run_oracle(meta, fx, torch.float64)
    for key in ("updated_query", "m_items_out", "grad_x"):
        if key in fx:
            assert_close(out[key].detach().float(), fx[key], 2e-5, f"{name}:{key}")


def test_fixture_structure():
    names = golden_names()
    assert len(names) >= 10
    meta, fx = load_golden("metatest_read_dM_c64")
    assert "grad_m_items" in fx and fx["grad_m_items"].abs().sum() > 0
    assert float(fx["div_loss"]) == 0.0 and float(fx["cls_loss"]) == 0.0  # writeloss == [0, 0]
    meta, fx = load_golden("eval_read_nomask_c64")
    assert float(fx["readloss"]) == 0.0 and "labels" not in fx


def test_soft_counts_are_integers_at_feature_resolution():
    """Labels at feature resolution: the resample is the identity, so counts are integers."""
    meta, fx = load_golden("labels_at_feature_res_c64")
    S, D = mo.class_sums(fx["f"], fx["labels"], meta["K"])
    hist = mo.label_histogram(fx["labels"], meta["K"])
    assert torch.equal(D, hist.to(D.dtype))


def test_soft_label_rows_sum_to_one():
    meta, fx = load_golden("train_write_c256_ragged")
    om = mo.soft_label_weights(fx["labels"], meta["K"], meta["h"], meta["w"])
    assert torch.allclose(om.sum(-1), torch.ones_like(om[..., 0]), atol=1e-6)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("gumbel", [False, True])
def test_oracle_matches_live_reference(gumbel):
    """Fresh random case against the live reference module (build container only)."""
    from pinthememory_b200 import synth

    torch.manual_seed(77)
    B, C, h, w, K = 2, 32, 9, 11, 19
    ref = build_reference_memory(K, C, 0.8, 1.0, gumbel)
    ora = mo.OracleMemorySup(K, C, C, 0.8, 1.0, gumbel)
    ora.load_state_dict(ref.state_dict())
    ora.m_items = ref.m_items.clone()
    x = synth.make_features(B, C, h, w, seed=5)
    labels = synth.make_labels(B, 40, 52, K, "iid", seed=6)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    torch.manual_seed(3)
    r = ref(xr, labels, True, False)
    torch.manual_seed(3)
    o = ora(xo, labels, True, False)
    for a, b, n in zip(o[:4], r[:4], ("updated_query", "score_query", "score_memory", "readloss")):
        assert_close(a.detach(), b.detach(), TOL, n)
    assert_close(ora.m_items.detach(), ref.m_items.detach(), TOL, "m_items")
    lr = r[0].square().sum() + r[3] + r[4][0] + r[4][1]
    lo = o[0].square().sum() + o[3] + o[4][0] + o[4][1]
    lr.backward()
    lo.backward()
    assert_close(xo.grad, xr.grad, TOL, "grad_x")
    for (n, po), (_, pr) in zip(ora.named_parameters(), ref.named_parameters()):
        assert_close(po.grad, pr.grad, 5e-6, n)
