"""CPU, world_size 2 over gloo: the one exchange step of the sharded write.

Each rank holds half of the batch; the product's autograd-aware all-reduce (pinthememory_b200.sharding)
is composed with the oracle's write. Every rank must end with the single-process global-batch memory, and
the gradient reaching its shard of the write feature must be W x the global-batch gradient (DDP's 1/W
mean over ranks then reproduces the global-batch parameter gradient exactly).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    from pinthememory_b200 import synth

    B, C, h, w, K = 4, 32, 6, 7, 19
    f = synth.make_features(B, C, h, w, seed=1).abs() + 0.05
    labels = synth.make_labels(B, 24, 28, K, "iid", seed=2)
    M = synth.make_memory(K, C, seed=3)
    Wc = 0.2 * synth.make_features(1, 1, K, C, seed=4).view(K, C)
    bc = torch.zeros(K)
    G = [synth.make_upstream_grad((K, C), seed=10 + r) for r in range(2)]
    return f, labels, M, Wc, bc, G, K


def _loss(wr, G):
    return 0.4 * wr["div_loss"] + 0.2 * wr["cls_loss"] + (wr["memory_new"] * G).sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import memory_oracle as mo
        from pinthememory_b200 import sharding

        torch.set_num_threads(1)
        f, labels, M, Wc, bc, G, K = _inputs()
        shard = sharding.ShardGroup()
        sl = slice(rank * 2, rank * 2 + 2)
        fr = f[sl].clone().requires_grad_(True)
        Wr = Wc.clone().requires_grad_(True)
        wr = mo.write(fr, labels[sl], M, 0.8, Wr, bc, reduce_fn=lambda t: sharding.AllReduceSum.apply(t, shard))
        _loss(wr, G[rank]).backward()
        out[rank] = dict(M_new=wr["memory_new"].detach(), df=fr.grad, dW=Wr.grad, D=wr["D"].detach())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_write_equals_global_batch():
    from oracle import memory_oracle as mo

    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)

    f, labels, M, Wc, bc, G, K = _inputs()
    fa = f.clone().requires_grad_(True)
    Wa = Wc.clone().requires_grad_(True)
    wr = mo.write(fa, labels, M, 0.8, Wa, bc)
    # DDP objective = mean over ranks of the per-rank loss
    (0.5 * (_loss(wr, G[0]) + _loss(wr, G[1]))).backward()

    for r in range(world):
        res = out[r]
        assert torch.allclose(res["M_new"], wr["memory_new"].detach(), atol=1e-6), "rank %d memory" % r
        assert torch.allclose(res["D"], wr["D"].detach(), atol=1e-5)
        assert torch.allclose(res["df"] / world, fa.grad[r * 2: r * 2 + 2], atol=1e-6, rtol=1e-4), "rank %d df" % r
    assert torch.equal(out[0]["M_new"], out[1]["M_new"]), "ranks must hold bit-identical memory"
    # classifier gradient: DDP averages the per-rank grads
    dW_mean = 0.5 * (out[0]["dW"] + out[1]["dW"])
    assert torch.allclose(dW_mean, Wa.grad, atol=1e-6, rtol=1e-4)


def test_single_rank_is_a_no_op():
    from pinthememory_b200 import sharding

    t = torch.arange(6.0)
    assert sharding.all_reduce_sum_(t, None) is t
    with pytest.raises(RuntimeError):
        sharding.ShardGroup()  # torch.distributed not initialised
