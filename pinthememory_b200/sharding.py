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


class PeerExchange:
    """Symmetric (peer-mapped) buffer of one module for the exchange fused into the update kernels (pm_update_fwd_peer /
    pm_update_bwd_peer, include/pinmem_b200.h): every rank allocates ``PM_PEER_BYTES`` through
    ``torch.distributed._symmetric_memory``, the rendezvous hands every rank the others' addresses, and the kernels read
    the class sums / their gradient straight out of the peers' buffers over NVLink -- no NCCL launch on the step."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm

        from . import capi

        pg = group if group is not None else dist.group.WORLD
        self.buffer = symm.empty(capi.PEER_BYTES // 4, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self.handle = symm.rendezvous(self.buffer, pg.group_name)
        self.rank, self.world_size = self.handle.rank, self.handle.world_size
        self.bufs_dev = int(self.handle.buffer_ptrs_dev)      # device array of all ranks' buffer addresses
        self.epochs = torch.zeros(2, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                              # every rank's flags are zero before anyone signals

    def sums_view(self, K, C):
        """This rank's [K+1, C+4] sums|counts region (pm_write_reduce_fwd accumulates into it; zero it first)."""
        return self.buffer[: (K + 1) * (C + 4)].view(K + 1, C + 4)


class ShardGroup:
    """Handle for the process group the class sums are reduced over (``None`` group = WORLD)."""

    def __init__(self, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("pinmem_b200.sharding: torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.peer = None          # a PeerExchange when the fused exchange is available (CUDA, NVLink peers, <= 16 ranks)
        self.peer_error = None    # why it is not (the NCCL all-reduce is used then)

    def enable_peer_exchange(self, device):
        """Try to set up the fused peer-memory exchange; on any failure keep the NCCL all-reduces (and remember why).
        Collective: every rank of the group must call it."""
        import os

        ok = 0
        if self.world_size > 1 and self.world_size <= 16 and not os.environ.get("PINMEM_B200_NCCL_EXCHANGE"):
            try:
                self.peer = PeerExchange(self.group, device)
                ok = 1
            except Exception as e:  # no symmetric-memory support / no peer access: fall back, identically on all ranks
                self.peer, self.peer_error = None, repr(e)[:300]
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.peer = None
        return self.peer is not None


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


def enable_sharded_update(module, group=None, peer_exchange=True):
    """Turn on the all-reduce of class sums for every ``Memory_sup`` inside ``module``.

    Call after ``init_process_group`` (and after DDP wrapping, or on the bare net) on EVERY rank (collective). Returns
    the (last) ``ShardGroup``. Without this call the module keeps the reference's rank-local update. With
    ``peer_exchange`` (default) the exchange is fused into the update kernels over NVLink peer memory when the platform
    supports it (``ShardGroup.peer``), else two NCCL all-reduces per step are issued.
    """
    from .memory import Memory_sup

    mems = [m for m in module.modules() if isinstance(m, Memory_sup)]
    if not mems:
        raise RuntimeError("enable_sharded_update: no Memory_sup inside the given module")
    shard = None
    for m in mems:   # one exchange buffer per memory module (each has its own sums and step order)
        shard = ShardGroup(group)
        if peer_exchange and m.m_items.is_cuda:
            shard.enable_peer_exchange(m.m_items.device)
        m.shard_group = shard
    return shard


def broadcast_memory(module, src=0, group=None):
    """Make ``m_items`` identical on all ranks (e.g. right after construction / checkpoint restore)."""
    from .memory import Memory_sup

    for m in module.modules():
        if isinstance(m, Memory_sup):
            t = m.m_items.detach().clone()
            dist.broadcast(t, src=src, group=group)
            m.m_items = t
