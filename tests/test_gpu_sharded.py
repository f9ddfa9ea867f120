"""GPU, 2 ranks over NCCL: batch-sharded Memory_sup == single-process global batch.

Needs >= 2 CUDA devices (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`).
Every rank runs the product module on its half of the batch with `enable_sharded_update`; the checker is the
oracle on the concatenated batch in one process. BatchNorm is put in eval mode so that the only cross-image
coupling is the class-sum exchange under test.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

K, C, B, h, w, Hm, Wm = 19, 64, 4, 12, 16, 48, 64
LW = dict(read=0.02, div=0.4, cls=0.2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(device):
    from pinthememory_b200 import synth

    x = synth.make_features(B, C, h, w, seed=21, device=device)
    labels = synth.make_labels(B, Hm, Wm, K, "blocky", seed=22, device=device, block=8)
    G = synth.make_upstream_grad((B, C, h, w), seed=23, device=device)
    return x, labels, G


def _state(device):
    """One module state shared by every rank and the checker (built on the CPU generator)."""
    from oracle import memory_oracle as mo

    torch.manual_seed(7)
    ora = mo.OracleMemorySup(K, C, C, 0.8, 1.0, False)
    with torch.no_grad():
        ora.clsfier.weight.normal_(0, 0.2)
        for m in ora.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    ora = ora.to(device).eval()
    ora.m_items = ora.m_items.to(device)  # a plain attribute, like the reference's: .to() does not move it
    return ora


def _worker(rank, world, port, out, exchange):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if exchange == "nccl":
        os.environ["PINMEM_B200_NCCL_EXCHANGE"] = "1"
    else:
        os.environ.pop("PINMEM_B200_NCCL_EXCHANGE", None)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from pinthememory_b200 import sharding
        from pinthememory_b200.memory import Memory_sup

        ora = _state(dev)
        mem = Memory_sup(K, C, C, 0.8, 1.0, False).to(dev).eval()
        mem.load_state_dict(ora.state_dict())
        mem.m_items = ora.m_items.clone()
        shard = sharding.enable_sharded_update(mem)
        x, labels, G = _inputs(dev)
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        xr = x[sl].clone().requires_grad_(True)
        uq, _, _, rl, wl = mem(xr, labels[sl].contiguous(), True, False)
        ((uq * G[sl]).sum() + LW["read"] * rl + LW["div"] * wl[0] + LW["cls"] * wl[1]).backward()
        out[rank] = dict(m_items=mem.m_items.detach().cpu(), dx=xr.grad.cpu(), rl=rl.detach().cpu(),
                         grads={k: p.grad.cpu() for k, p in mem.named_parameters() if p.grad is not None},
                         D=mem.last_class_sums[:, C].cpu(), peer=shard.peer is not None, peer_error=shard.peer_error)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_rank_nccl_sharded_module_equals_global_batch(exchange):
    """exchange = "peer": the all-reduce of the class sums (and of their gradient) fused into the update kernels over
    NVLink peer memory (pm_update_fwd_peer / pm_update_bwd_peer); "nccl": two NCCL all-reduces around the kernels."""
    import torch.multiprocessing as mp

    from golden_util import assert_close
    from oracle import memory_oracle as mo

    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, exchange), nprocs=world, join=True)
    if exchange == "peer" and not out[0]["peer"]:
        pytest.skip("no symmetric-memory peer exchange on this box: %s" % out[0]["peer_error"])
    assert out[0]["peer"] == (exchange == "peer")

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    ora = _state(dev)
    x, labels, G = _inputs(dev)
    xa = x.clone().requires_grad_(True)
    M0 = ora.m_items.clone()
    wr = mo.write(ora.writenet(xa), labels, M0, 0.8, ora.clsfier.weight, ora.clsfier.bias)
    n = B // world
    total = 0
    rls = []
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        rd = mo.read(xa[sl], M0, labels[sl], 1.0)
        uq = ora.output(rd["u"])
        rls.append(rd["readloss"].detach().cpu())
        total = total + ((uq * G[sl]).sum() + LW["read"] * rd["readloss"] + LW["div"] * wr["div_loss"] +
                         LW["cls"] * wr["cls_loss"]) / world
    total.backward()

    assert torch.equal(out[0]["m_items"], out[1]["m_items"]), "ranks must hold bit-identical memory"
    for r in range(world):
        assert_close(out[r]["m_items"], wr["memory_new"].detach().cpu(), 1e-5, "m_items rank %d" % r)
        assert_close(out[r]["D"][:K], wr["D"][:K].detach().cpu(), 1e-6, "global counts")
        assert_close(out[r]["rl"], rls[r], 1e-5, "rank-local readloss")
        assert_close(out[r]["dx"] / world, xa.grad[r * n:(r + 1) * n].cpu(), 2e-5, "dx rank %d" % r)
    ref_grads = {k: p.grad.cpu() for k, p in ora.named_parameters() if p.grad is not None}
    for k, g in ref_grads.items():
        mean = sum(out[r]["grads"][k] for r in range(world)) / world  # what DDP's all-reduce(mean) produces
        assert_close(mean, g, 5e-5, k)


# ------------------------------------------------------------------ SyncBatchNorm (the reference under --syncbn)


def _worker_syncbn(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.pop("PINMEM_B200_NCCL_EXCHANGE", None)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from pinthememory_b200 import memory as pm_memory
        from pinthememory_b200 import sharding
        from pinthememory_b200.memory import Memory_sup

        ora = _state(dev)
        mem = Memory_sup(K, C, C, 0.8, 1.0, False).to(dev)
        mem.load_state_dict(ora.state_dict())
        mem.m_items = ora.m_items.clone()
        mem = torch.nn.SyncBatchNorm.convert_sync_batchnorm(mem).train()      # train.py:95
        assert type(mem.output[1]) is torch.nn.SyncBatchNorm and type(mem.writenet.writefeat[1]) is torch.nn.SyncBatchNorm
        mem.fold_min_pixels = 0
        sharding.enable_sharded_update(mem)
        x, labels, G = _inputs(dev)
        n = B // world
        sl = slice(rank * n, (rank + 1) * n)
        xr = x[sl].clone().requires_grad_(True)
        pm_memory._warned.clear()
        uq, _, _, rl, wl = mem(xr, labels[sl].contiguous(), True, False)
        ((uq * G[sl]).sum() + LW["div"] * wl[0] + LW["cls"] * wl[1]).backward()
        out[rank] = dict(m_items=mem.m_items.detach().cpu(), dx=xr.grad.cpu(), uq=uq.detach().cpu(),
                         grads={k: p.grad.cpu() for k, p in mem.named_parameters() if p.grad is not None},
                         rmean=mem.output[1].running_mean.cpu(), rvar=mem.writenet.writefeat[1].running_var.cpu(),
                         fell_back="libconv" in pm_memory._warned)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_syncbatchnorm_module_equals_global_batch_batchnorm():
    """``convert_sync_batchnorm`` turns the module's two BatchNorm2d into SyncBatchNorm (the reference's multi-GPU scripts
    pass --syncbn). The fused conv+BN path must then normalise with the statistics of the GLOBAL batch: the GEMM
    epilogue's fp64 sums (+ the pixel count) are all-reduced before the normalise pass, the two backward sums before the
    input-gradient pass. Checker: the oracle with plain BatchNorm2d in training mode on the concatenated batch."""
    import torch.multiprocessing as mp

    from golden_util import assert_close
    from oracle import memory_oracle as mo

    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_syncbn, args=(world, _free_port(), out), nprocs=world, join=True)
    assert not out[0]["fell_back"], "the SyncBatchNorm module must stay on the fused tcgen05 conv + BatchNorm path"

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    ora = _state(dev).train()
    x, labels, G = _inputs(dev)
    xa = x.clone().requires_grad_(True)
    M0 = ora.m_items.clone()
    wr = mo.write(ora.writenet(xa), labels, M0, 0.8, ora.clsfier.weight, ora.clsfier.bias)   # BN over the whole batch
    n = B // world
    us = [mo.read(xa[r * n:(r + 1) * n], M0, None, 1.0)["u"] for r in range(world)]     # the read is per pixel
    uq = ora.output(torch.cat(us))                                                        # BN over the whole batch
    ((uq * G).sum() + LW["div"] * wr["div_loss"] + LW["cls"] * wr["cls_loss"]).backward()

    assert torch.equal(out[0]["m_items"], out[1]["m_items"])
    assert_close(out[0]["m_items"], wr["memory_new"].detach().cpu(), 1e-5, "m_items")
    for r in range(world):
        assert_close(out[r]["uq"], uq[r * n:(r + 1) * n].detach().cpu(), 1e-5, "updated_query rank %d" % r)
        assert_close(out[r]["rmean"], ora.output[1].running_mean.cpu(), 1e-5, "running_mean (global statistics)")
        assert_close(out[r]["rvar"], ora.writenet.writefeat[1].running_var.cpu(), 1e-5, "running_var (unbiased, global count)")
    ref = {k: p.grad.cpu() for k, p in ora.named_parameters() if p.grad is not None}
    for k, g in ref.items():
        total = sum(out[r]["grads"][k] for r in range(world))
        if k.startswith("output."):
            # the read branch: <uq, G> is a sum over the rank's own pixels, so the ranks' gradients ADD UP to the global one
            assert_close(total, g, 5e-5, k)
        else:
            # the write branch: div / cls are functions of the (global) memory, identical on every rank, and the backward
            # all-reduces dS -- every rank holds its pixels' share of W times the global gradient: DDP's mean restores it
            assert_close(total / world, g, 5e-5, k)
