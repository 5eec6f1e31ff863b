"""not gpu: the N>1 host logic (crop sharding + the single gather of decoded keypoints) on two gloo
processes.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from buctd_b200.distributed import gather_keypoints, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 17 * 3, dtype=torch.float32).view(n_total, 17, 3)
        lo, hi = shard_range(n_total, rank, world)
        got = gather_keypoints(full[lo:hi].clone(), n_total)
        q.put((rank, bool(torch.equal(got, full)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])      # even and ragged shards
def test_gather_keypoints_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (n_total, 17, 3) for _, _, shape in res)


def _grad_worker(rank, world, port, q):
    """data-parallel training step: per-rank JointsMSELoss gradients (oracle closed form) averaged by the ONE
    all-reduce == the gradient of the loss over the concatenated global batch (equal shards)"""
    import numpy as np
    from buctd_b200.distributed import allreduce_mean_
    from oracle import loop_oracle as LO
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        o = rng.standard_normal((4 * world, 17, 8, 6)).astype(np.float32)
        t = rng.uniform(0, 1, o.shape).astype(np.float32)
        w = (rng.uniform(size=(4 * world, 17, 1)) < 0.8).astype(np.float32)
        lin = rng.standard_normal((5, 8 * 6)).astype(np.float32)          # a toy "model": param grad = lin @ dL/dout rows
        lo, hi = shard_range(4 * world, rank, world)
        g_local = LO.joints_mse_grad(o[lo:hi], t[lo:hi], w[lo:hi], True)
        flat = torch.from_numpy((lin @ g_local.reshape(-1, 48).T).sum(1).copy())
        allreduce_mean_(flat)
        g_global = LO.joints_mse_grad(o, t, w, True)
        want = (lin @ g_global.reshape(-1, 48).T).sum(1)
        q.put((rank, float(np.abs(flat.numpy() - want).max())))
    finally:
        dist.destroy_process_group()


def test_allreduce_mean_matches_global_batch_gradient_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err < 1e-6 for _, err in res), res


def test_longest_final_range():
    """the early bucket of the bucketed gradient all-reduce: longest contiguous run of parameters whose last writer precedes
    the cut (buctd_b200.train.longest_final_range; the GPU test checks it against the recorded step)"""
    from buctd_b200.train import longest_final_range
    poff = {"a": 0, "b": 10, "c": 30, "d": 35, "e": 60}
    n = 100
    last = {"a": 90, "b": 5, "c": 7, "d": 80, "e": 3}
    assert longest_final_range(poff, n, last, 50) == (60, 100)            # e alone (40) beats b + c (25)
    assert longest_final_range(poff, n, last, 85) == (10, 100)            # b .. e
    assert longest_final_range(poff, n, last, 4) == (60, 100)
    assert longest_final_range(poff, n, last, 3) == (0, 0)
    assert longest_final_range(poff, n, {"b": 1, "c": 1}, 50) == (10, 35)  # parameters without a writer never qualify
    assert longest_final_range(poff, n, last, 1000) == (0, 100)


def _bucket_worker(rank, world, port, q):
    """ONE logical all-reduce issued as buckets (an asynchronous early range + the rest) == the single call"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11 + rank)
        flat = torch.randn(1000, generator=g)
        whole = flat.clone()
        dist.all_reduce(whole, op=dist.ReduceOp.SUM)
        whole /= world
        lo, hi = 333, 801
        work = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)       # gloo has no AVG: sum, then scale
        dist.all_reduce(flat[:lo], op=dist.ReduceOp.SUM)
        dist.all_reduce(flat[hi:], op=dist.ReduceOp.SUM)
        work.wait()
        flat /= world
        q.put((rank, bool(torch.equal(flat, whole))))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_call_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
