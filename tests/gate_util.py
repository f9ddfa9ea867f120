"""ReLU tie handling shared by the GPU parity tests (see oracle/memory_oracle.py::ReluGates for the why).

``record_gates()`` switches on the product's debug hook and collects the gates of every ReLU the module evaluates;
``oracle_outputs`` runs the oracle on the same inputs with those gates replayed; ``check_ties`` accepts a replayed gate
only where the oracle's own pre-activation is zero at rounding level."""
import contextlib

import torch

from golden_util import LOSS_WEIGHTS

TIE_RATIO = 2e-5      # |z| / max|z| below which a ReLU input counts as a rounding-level tie
TIE_FRACTION = 1e-4   # at most this fraction of a map may be ties (a handful in practice)


@contextlib.contextmanager
def record_gates():
    from pinthememory_b200 import memory as pm_memory

    log, old = {}, pm_memory.GATE_LOG
    pm_memory.GATE_LOG = log
    try:
        yield log
    finally:
        pm_memory.GATE_LOG = old


def check_ties(relu_gates):
    """Every gate the oracle would have decided differently must be a rounding-level tie; returns their number."""
    total = 0
    for name, n, numel, ratio in relu_gates.mismatches:
        assert ratio <= TIE_RATIO, f"{name}: ReLU gate differs where |z|/max|z| = {ratio:.2e} (not a rounding tie)"
        assert n <= max(3, TIE_FRACTION * numel), f"{name}: {n} of {numel} ReLU gates differ"
        total += n
    return total


def oracle_module_like(mem, gates=None, dtype=torch.float32):
    from oracle import memory_oracle as mo

    o = mo.OracleMemorySup(mem.memory_size, mem.feature_dim, mem.feature_dim, mem.momentum, mem.temperature,
                           mem.gumbel_read).to(next(mem.parameters()).device).to(dtype)
    o.load_state_dict({k: v.to(dtype) if v.is_floating_point() else v for k, v in mem.state_dict().items()})
    o.m_items = mem.m_items.detach().clone().to(dtype)
    o.train(mem.training)
    if gates is not None:
        o.relu_gates = mo.ReluGates(gates)
    return o


def oracle_fixture_outputs(ora, meta, fx, mem_grad):
    """Run the oracle module on a fixture's inputs; returns a dict with the fixture's key names."""
    if mem_grad:
        ora.m_items = ora.m_items.clone().requires_grad_(True)
    mem_in = ora.m_items
    x = fx["x"].clone().requires_grad_(meta["backward"])
    labels = fx.get("labels")
    noise = (fx["g_query"], fx["g_memory"]) if meta.get("gumbel") else None
    uq, sq, sm, rl, wl = ora(x, labels, meta["writing"], meta["detach"], noise=noise)
    out = {"updated_query": uq.detach(), "score_query": sq.detach(), "score_memory": sm.detach(),
           "m_items_out": ora.m_items.detach()}
    if labels is not None:
        out["readloss"] = rl.detach()
    if meta["writing"]:
        out["div_loss"], out["cls_loss"] = wl[0].detach(), wl[1].detach()
    if meta["backward"]:
        total = (uq * fx["G"]).sum()
        if labels is not None:
            total = total + LOSS_WEIGHTS["read"] * rl
        if meta["writing"]:
            total = total + LOSS_WEIGHTS["div"] * wl[0] + LOSS_WEIGHTS["cls"] * wl[1]
        total.backward()
        out["grad_x"] = x.grad
        if mem_grad:
            out["grad_m_items"] = mem_in.grad
        for n, p in ora.named_parameters():
            if p.grad is not None:
                out["grad_param." + n] = p.grad
    return out
