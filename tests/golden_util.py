"""Helpers shared by the parity tests: golden fixtures and comparison metrics."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_WEIGHTS = dict(read=0.02, div=0.4, cls=0.2)  # train.py:1213-1215 in the reference


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name, device="cpu"):
    """Return (meta dict, {key: tensor})."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    fx = {k: torch.from_numpy(z[k]).to(device) for k in z.files if k != "meta"}
    return meta, fx


def load_state(module, fx):
    """Copy the fixture's state_dict + m_items into a module with the reference's key names."""
    sd = {k[len("state."):]: v for k, v in fx.items() if k.startswith("state.")}
    dev = next(module.parameters()).device
    module.load_state_dict({k: v.to(dev) for k, v in sd.items()})
    module.m_items = fx["m_items_in"].to(dev).clone()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()


def max_abs_over_scale(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    scale = b.abs().max().item()
    return (a - b).abs().max().item() / scale if scale > 0 else (a - b).abs().max().item()


def assert_close(a, b, tol, what=""):
    """Parity metric used everywhere: rel-L2 and max-abs/scale both within ``tol``."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    r, m = rel_l2(a, b), max_abs_over_scale(a, b)
    assert r <= tol and m <= tol, f"{what}: rel_l2={r:.3e} max_abs/scale={m:.3e} > tol={tol:.1e}"
