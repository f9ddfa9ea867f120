"""CPU, build container only: the UNCHANGED reference head picks up this implementation by module name.

Builds the reference's DeepR50V3PlusD (network/deepv3plus.py) with `--memory` after
`pinthememory_b200.install()`; the recipe for importing the reference offline is SURVEY.md appendix A.
Skipped on the GPU box (no /root/reference there).
"""
import argparse
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("PINMEM_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "network")), reason="reference tree not mounted")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def test_reference_head_constructs_and_calls_our_module(monkeypatch):
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        # third-party packages the reference imports but this image lacks (not on the memory path)
        noop = lambda *a, **k: None
        sk = _stub("skimage")
        sk.__path__ = []
        for sub, attrs in (("color", {}), ("filters", {"gaussian": noop}), ("restoration", {"denoise_bilateral": noop}),
                           ("segmentation", {"find_boundaries": noop}), ("util", {"random_noise": noop}),
                           ("io", {}), ("transform", {})):
            setattr(sk, sub, _stub("skimage." + sub, **attrs))
        _stub("imageio", imread=noop, imwrite=noop)
        _stub("tensorboardX", SummaryWriter=type("SummaryWriter", (), {"__init__": noop}))
        _stub("kmeans1d", cluster=noop)
        import torch.utils.model_zoo as model_zoo

        monkeypatch.setattr(model_zoo, "load_url", lambda *a, **k: {})  # no network: random-init trunk
        monkeypatch.setenv("PINMEM_B200_DEVICE", "cpu")
        for k in [k for k in sys.modules if k == "network" or k.startswith("network.") or k in ("config", "datasets")]:
            del sys.modules[k]
        sys.path.insert(0, REF)

        import pinthememory_b200

        mod = pinthememory_b200.install()  # before `import network`, as INTEGRATION.md says
        from config import assert_and_infer_cfg

        args = argparse.Namespace(syncbn=False, wt_layer=[0] * 7, relax_denom=0.0, clusters=50, memory=True, mem_slot=19,
                                  mem_dim=256, mem_momentum=0.8, mem_temp=1.0, gumbel_off=True, use_wtloss=False,
                                  arch="network.deepv3plus.DeepR50V3PlusD", batch_weighting=False, strict_bdr_cls=None,
                                  rlx_off_iter=None, rlx_off_epoch=-1, cov_stat_epoch=0, jointwtborder=False,
                                  dataset=["gtav"], exp="t", tb_tag="t", ckpt="/tmp", tb_path="/tmp", date="0", snapshot=None)
        assert_and_infer_cfg(args, train_mode=False)
        import network

        assert sys.modules["network.memory"] is mod
        crit = torch.nn.CrossEntropyLoss(ignore_index=255)
        net = network.get_model(args, 19, crit, crit)
        assert type(net.memory) is pinthememory_b200.Memory_sup
        assert type(net.memory.writenet) is mod.Writingnet
        keys = [k for k in net.state_dict() if k.startswith("memory.")]
        assert "memory.output.0.weight" in keys and "memory.writenet.writefeat.1.running_var" in keys
        assert "memory.clsfier.bias" in keys and not any("m_items" in k for k in keys)
        assert tuple(net.memory.m_items.shape) == (19, 256)
        # the call site deepv3plus.py:561 reaches our forward (which refuses to run without a GPU)
        net.eval()
        with pytest.raises(RuntimeError, match="no CPU path"):
            with torch.no_grad():
                net(torch.randn(1, 3, 64, 64))
    finally:
        sys.path[:] = saved_path
        # drop only what this test brought in (reference packages and the stubs); never torch's own lazily
        # imported sub-modules, which must not be imported twice
        stubs = ("skimage", "imageio", "tensorboardX", "kmeans1d")
        for k in list(sys.modules):
            if k in saved:
                continue
            f = getattr(sys.modules[k], "__file__", None) or ""
            if f.startswith(REF) or k.split(".")[0] in stubs or k == "network.memory":
                del sys.modules[k]
        for k in stubs + ("network", "network.memory", "config", "datasets", "transforms", "transforms.transforms"):
            if k in saved:
                sys.modules[k] = saved[k]
