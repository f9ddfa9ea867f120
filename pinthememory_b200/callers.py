"""The two callers either side of the memory path, on the same kernels.

* ``PrototypePool`` / ``initialize_memory`` -- the reference's ``Trainer.memory_initalize``
  (train.py:1000-1042): class prototypes pooled over the training set. One ``pm_write_reduce_fwd`` launch
  per batch accumulates the packed class sums|counts in place (the kernel only ever adds into its output);
  with a shard group the sums are all-reduced once at the end, so every rank starts from the same memory.
* ``upsampled_cross_entropy`` -- the head's main loss ``criterion(Upsample(logits, size), gts)``
  (network/deepv3plus.py:575-578, mynn.py:57-62, loss.py:167-180) through ``pm_readloss_fwd``: the
  ``[B,K,Hm,Wm]`` up-sampled logits and their log-softmax never exist in memory (for K=19 at 768x768 and
  batch 8 that is 358 MB written and re-read several times by the eager path).

CUDA only, like the rest of the package.
"""
import torch
import torch.nn.functional as F

from . import capi, sharding


class PrototypePool:
    """Accumulates per-class sums of L2-normalised features weighted by the down-sampled one-hot labels."""

    def __init__(self, memory_size, feature_dim, device="cuda", group=None):
        self.K, self.C, self.group = int(memory_size), int(feature_dim), group
        self.sums_counts = torch.zeros(self.K + 1, self.C + 4, dtype=torch.float32, device=device)  # [K+1, C | count,0,0,0]
        self.batches = 0

    def add(self, features, labels):
        """features [B,C,h,w] fp32|bf16 (un-normalised: train.py:1019 normalises, so does the kernel),
        labels [B,Hm,Wm] int64 with 255 = ignore."""
        capi.require_cuda(features, labels)
        if features.dim() != 4 or features.shape[1] != self.C:
            raise RuntimeError(f"pinmem_b200: features must be [B,{self.C},h,w], got {tuple(features.shape)}")
        if labels.dtype != torch.int64 or labels.dim() != 3 or labels.shape[0] != features.shape[0]:
            raise RuntimeError("pinmem_b200: labels must be int64 [B,Hm,Wm]")
        capi.write_reduce_fwd(features.detach().contiguous(), labels.contiguous(), self.sums_counts, self.K)
        self.batches += 1
        return self

    def finalize(self):
        """train.py:1037-1040: ``normalize(basket / count)`` with empty classes' counts set to 1 (zero rows)."""
        sd = self.sums_counts
        if self.group is not None:
            sd = sharding.all_reduce_sum_(sd.clone(), self.group)
        sums, counts = sd[: self.K, : self.C], sd[: self.K, self.C]
        counts = torch.where(counts == 0, torch.ones_like(counts), counts)
        return F.normalize(sums / counts.unsqueeze(1), dim=1)


def initialize_memory(module, batches, group=None):
    """``module.m_items <- prototypes`` from an iterable of (features, labels); returns the pool."""
    pool = PrototypePool(module.memory_size, module.feature_dim, module.m_items.device,
                         group if group is not None else getattr(module, "shard_group", None))
    for features, labels in batches:
        pool.add(features, labels)
    module.m_items = pool.finalize()
    return pool


class _UpsampledCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        B, K, h, w = logits.shape
        N = B * h * w
        KP = capi.score_stride(K)
        dev = logits.device
        s = torch.zeros(N, KP, dtype=torch.float32, device=dev)
        s[:, :K] = logits.detach().permute(0, 2, 3, 1).reshape(N, K)
        buf = torch.zeros(N * KP + 2 * capi.WS_WORDS + 4, dtype=torch.float32, device=dev)
        ds = buf[: N * KP]
        ws = buf[N * KP: N * KP + 2 * capi.WS_WORDS]
        out = buf[N * KP + 2 * capi.WS_WORDS:]
        capi.readloss(s, labels, 1.0, B, h, w, K, ds, ws, out)
        ctx.save_for_backward(ds, out)
        ctx.shape, ctx.dtype = (B, K, h, w), logits.dtype
        return out[0].to(logits.dtype)

    @staticmethod
    def backward(ctx, g):
        ds, out = ctx.saved_tensors
        B, K, h, w = ctx.shape
        grad = (ds.view(B, h, w, -1)[..., :K] * (g.to(torch.float32) * out[1])).permute(0, 3, 1, 2)
        return grad.to(ctx.dtype).contiguous(), None


def upsampled_cross_entropy(logits, labels):
    """``CrossEntropyLoss2d(ignore_index=255)(Upsample(logits, labels.shape[-2:]), labels)`` fused.

    logits [B,K,h,w] (K <= 31) fp32|bf16 CUDA, labels [B,Hm,Wm] int64. Mean over the non-ignored pixels
    (NaN if there is none, like torch). Differentiable w.r.t. ``logits``.
    """
    capi.require_cuda(logits, labels)
    if logits.dim() != 4 or labels.dim() != 3 or labels.shape[0] != logits.shape[0]:
        raise RuntimeError("pinmem_b200: logits must be [B,K,h,w] and labels [B,Hm,Wm]")
    if labels.dtype not in (torch.int64, torch.uint8):
        raise RuntimeError("pinmem_b200: labels must be int64 (or uint8 class ids)")
    if not 1 <= logits.shape[1] <= 31:
        raise RuntimeError("pinmem_b200: upsampled_cross_entropy supports 1..31 classes")
    return _UpsampledCE.apply(logits, labels.contiguous())
