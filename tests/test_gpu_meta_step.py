"""BASELINE config 3: the memory module's part of one GS meta-train step (train.py:530-583) on the CUDA path.

The sequence (pinthememory_b200/metastep.py) is the reference's: A = forward with write + ``backward(retain_graph=True)``,
theta' = theta - lr*grad installed in two module copies through ``_parameters`` (``put_theta``, train.py:262-277),
B = write on the saved memory with graph, C = read of the graph-carrying memory on the meta-test batch + backward,
D = eval-mode no-grad write on the saved memory. The checker is the oracle run through the very same function in fp64
(these gradients cross BatchNorm statistics several times, so fp32-vs-fp32 would compare two rounding errors).
"""
import pytest
import torch

from gate_util import check_ties, record_gates
from golden_util import assert_close

pytestmark = pytest.mark.gpu


def _trio(cls, K, C, state=None, memory=None, dtype=torch.float32):
    mods = []
    for i in range(3):
        torch.manual_seed(7)
        m = cls(K, C, C, 0.8, 1.0, False).cuda().to(dtype)
        if state is not None:
            m.load_state_dict(state)
        m.m_items = (memory if memory is not None else m.m_items).clone().to(dtype)
        m.train()
        mods.append(m)
    return mods


@pytest.mark.parametrize("shape", [(2, 64, 12, 12, 48, 48), (4, 256, 48, 48, 192, 192)], ids=["c64", "cfg3_os16"])
def test_meta_train_step_matches_fp64_oracle(shape):
    from oracle import memory_oracle as mo
    from pinthememory_b200 import synth
    from pinthememory_b200.memory import Memory_sup
    from pinthememory_b200.metastep import meta_step

    B, C, h, w, Hm, Wm = shape
    K = 19
    net, upd, upd2 = _trio(Memory_sup, K, C)
    with torch.no_grad():
        net.clsfier.weight.normal_(0, 0.2)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    mem0 = net.m_items.clone()
    o_net, o_upd, o_upd2 = _trio(mo.OracleMemorySup, K, C, {k: v.double() for k, v in state.items()}, mem0, torch.float64)
    x_tr = synth.make_features(B, C, h, w, seed=11, device="cuda")
    x_te = synth.make_features(B, C, h, w, seed=12, device="cuda")
    lab_tr = synth.make_labels(B, Hm, Wm, K, "blocky", seed=13, device="cuda", block=16)
    lab_te = synth.make_labels(B, Hm, Wm, K, "blocky", seed=14, device="cuda", block=16)
    G_tr = synth.make_upstream_grad((B, C, h, w), seed=15, device="cuda") / (B * h * w)
    G_te = synth.make_upstream_grad((B, C, h, w), seed=16, device="cuda") / (B * h * w)

    with record_gates() as gates:
        got = meta_step(net, upd, upd2, x_tr, lab_tr, x_te, lab_te, G_tr, G_te, inner_lr=0.01)
    # rounding-level ReLU ties are replayed the module's way; the three oracle instances share one FIFO per ReLU
    shared = mo.ReluGates(gates)
    for o in (o_net, o_upd, o_upd2):
        o.relu_gates = shared
    ref = meta_step(o_net, o_upd, o_upd2, x_tr.double(), lab_tr, x_te.double(), lab_te, G_tr.double(), G_te.double(),
                    inner_lr=0.01)
    # yardstick for the ill-conditioned gradients: the same oracle in fp32 (eager torch / cuDNN fp32) against its fp64 self
    f_net, f_upd, f_upd2 = _trio(mo.OracleMemorySup, K, C, state, mem0)
    shared32 = mo.ReluGates(gates)
    for o in (f_net, f_upd, f_upd2):
        o.relu_gates = shared32
    ref32 = meta_step(f_net, f_upd, f_upd2, x_tr, lab_tr, x_te, lab_te, G_tr, G_te, inner_lr=0.01)

    check_ties(shared)
    for k in ("inner_loss", "outer_loss", "readloss_a", "div_a", "cls_a", "readloss_c"):
        assert_close(got[k].reshape(1), ref[k].reshape(1), 1e-5, k)
    assert_close(got["uq_c"], ref["uq_c"], 2e-5, "meta-test read output")
    assert_close(got["memory_b"], ref["memory_b"], 1e-5, "memory after the meta-test write (B)")
    assert_close(got["memory_final"], ref["memory_final"], 1e-5, "final memory (D)")
    assert set(got["grads"]) == set(ref["grads"])
    assert "writenet.writefeat.0.weight" in got["grads"]
    # gradients that crossed the BatchNorm statistics of up to three passes: as close to fp64 as fp32 arithmetic gets,
    # i.e. within 5e-5 or 3x the deviation of eager fp32 torch from its own fp64 evaluation, whichever is larger
    from golden_util import max_abs_over_scale, rel_l2

    for name in ("inner_grads", "grads"):
        for k in ref[name]:
            yard = max(rel_l2(ref32[name][k], ref[name][k]), max_abs_over_scale(ref32[name][k], ref[name][k]))
            assert_close(got[name][k], ref[name][k], max(5e-5, 3.0 * yard), name + " " + k)
    # aliasing rules of the step: D rebinds m_items to a fresh detached tensor; the functional copies hold non-leaf
    # parameters that point back at net's
    assert not net.m_items.requires_grad
    assert upd.output[0].weight.grad_fn is not None and not isinstance(upd.output[0].weight, torch.nn.Parameter)
