"""Deterministic synthetic inputs for the memory path.

Shapes follow the Cityscapes / GTAV crops the reference trains on: features are the
ASPP bottleneck map (C=256) at output stride 8 or 16 of a 768x768 crop, labels are the
full-resolution trainId map (int64, 255 = ignore; transforms.py:95-97,
joint_transforms.py:109-112 in the reference). Generated on the CPU generator so the
same seed gives the same tensors on every host, then moved to ``device``.
"""
import torch

SEED = 304  # the reference's fixed seed (config.py:52)
IGNORE_LABEL = 255

# (B, h, w, Hm, Wm) per BASELINE.json config
WORKLOADS = {
    "cfg1_dr50v3p_os16_b2": dict(B=2, h=48, w=48, Hm=768, Wm=768),
    "cfg2_module_os8_b8": dict(B=8, h=96, w=96, Hm=768, Wm=768),
    "cfg3_meta_os16_b4": dict(B=4, h=48, w=48, Hm=768, Wm=768),
    "cfg4_dp_os16_b8": dict(B=8, h=48, w=48, Hm=768, Wm=768),
    "cfg5_dr101v2_eval_b1": dict(B=1, h=128, w=256, Hm=0, Wm=0),
}


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_features(B, C, h, w, seed=SEED, dtype=torch.float32, device="cpu"):
    """x ~ N(0,1), NCHW contiguous."""
    x = torch.randn(B, C, h, w, generator=_gen(seed), dtype=torch.float32)
    return x.to(device=device, dtype=dtype).contiguous()


def make_memory(K, C, seed=SEED + 1, device="cpu"):
    """Unit-norm rows of a uniform random matrix, as memory.py:120 initialises m_items."""
    m = torch.rand(K, C, generator=_gen(seed), dtype=torch.float32)
    return torch.nn.functional.normalize(m, dim=1).to(device)


def make_labels(B, Hm, Wm, K=19, kind="blocky", seed=SEED + 2, device="cpu", block=32):
    """int64 label maps with 255 = ignore.

    ``iid``: uniform over the K classes per pixel plus a 1/16-height band of 255 at the
    top (worst case for the soft-label taps: every tap a different class).
    ``blocky``: constant ``block`` x ``block`` regions, classes drawn from a fixed
    Zipf-like distribution with the last class forced absent, 5% of blocks = 255.
    """
    g = _gen(seed)
    if kind == "iid":
        lab = torch.randint(0, K, (B, Hm, Wm), generator=g, dtype=torch.int64)
        lab[:, : max(Hm // 16, 1)] = IGNORE_LABEL
    elif kind == "blocky":
        by, bx = -(-Hm // block), -(-Wm // block)
        p = 1.0 / torch.arange(1, K + 1, dtype=torch.float64)
        p[K - 1] = 0.0  # one class absent from the batch
        p = p / p.sum()
        cls = torch.multinomial(p, B * by * bx, replacement=True, generator=g).view(B, by, bx)
        ign = torch.rand(B, by, bx, generator=g) < 0.05
        cls = torch.where(ign, torch.full_like(cls, IGNORE_LABEL), cls)
        lab = cls.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :Hm, :Wm].contiguous()
    else:
        raise ValueError(kind)
    return lab.to(device)


def make_upstream_grad(shape, seed=SEED + 3, dtype=torch.float32, device="cpu"):
    """Upstream gradient G ~ N(0,1) for the read output."""
    g = torch.randn(*shape, generator=_gen(seed), dtype=torch.float32)
    return g.to(device=device, dtype=dtype).contiguous()


def make_gumbel_noise(N, K, seed=SEED + 4, device="cpu"):
    """A (g_query, g_memory) pair distributed like F.gumbel_softmax's noise."""
    gen = _gen(seed)
    e = torch.empty(2, N, K, dtype=torch.float32).exponential_(generator=gen)
    g = -e.log()
    return g[0].contiguous().to(device), g[1].contiguous().to(device)
