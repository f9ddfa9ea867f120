"""Batch sharding of the memory path across the GPUs of one box.

Read (including the read loss) is per-pixel independent given the memory, so ranks just process their own
images. The write has ONE exchange step: the packed per-class sums|counts ``[K+1, C+4]`` fp32 (~21 KB)
are all-reduced (sum) over NCCL/NVLink before the momentum update, and -- the autograd mirror -- the
gradient w.r.t. the class sums ``[K, C]`` is all-reduced in the backward. Every rank then holds a
bit-identical ``m_items`` equal to the single-process update on the concatenated global batch.
The reference has no such collective (each rank silently keeps its own memory).
"""
import torch
import torch.distributed as dist


class ShardGroup:
    """Handle for the process group the class sums are reduced over (``None`` group = WORLD)."""

    def __init__(self, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("pinmem_b200.sharding: torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)


def all_reduce_sum_(t, shard):
    """In-place sum all-reduce on the caller's current stream; a no-op for a single rank."""
    if shard is None or shard.world_size == 1:
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=shard.group)
    return t


class AllReduceSum(torch.autograd.Function):
    """Autograd-aware sum all-reduce: forward sums over ranks, backward sums the gradients over ranks.

    This is the exchange step as a differentiable op; ``Memory_sup`` issues the same two collectives
    around its update kernel. Exposed so host-side logic (and the gloo CPU tests) can compose it.
    """

    @staticmethod
    def forward(ctx, t, shard):
        ctx.shard = shard
        return all_reduce_sum_(t.clone(), shard)

    @staticmethod
    def backward(ctx, g):
        return all_reduce_sum_(g.contiguous().clone(), ctx.shard), None


def enable_sharded_update(module, group=None):
    """Turn on the all-reduce of class sums for every ``Memory_sup`` inside ``module``.

    Call after ``init_process_group`` (and after DDP wrapping, or on the bare net). Returns the
    ``ShardGroup``. Without this call the module keeps the reference's rank-local update.
    """
    from .memory import Memory_sup

    shard = ShardGroup(group)
    found = 0
    for m in module.modules():
        if isinstance(m, Memory_sup):
            m.shard_group = shard
            found += 1
    if found == 0:
        raise RuntimeError("enable_sharded_update: no Memory_sup inside the given module")
    return shard


def broadcast_memory(module, src=0, group=None):
    """Make ``m_items`` identical on all ranks (e.g. right after construction / checkpoint restore)."""
    from .memory import Memory_sup

    for m in module.modules():
        if isinstance(m, Memory_sup):
            t = m.m_items.detach().clone()
            dist.broadcast(t, src=src, group=group)
            m.m_items = t
