"""Multi-GPU plumbing: one process per GPU, crops sharded over ranks, ONE
collective per step (the decoded keypoints, 204 B per crop) -- replaces the reference's single-process
``torch.nn.DataParallel`` (tools/test.py:134), which re-broadcasts 463 MB of parameters and gathers the
full heat-maps to GPU 0 on every forward.  torch.distributed does the plumbing (NCCL on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n crops: the first n % world ranks get one extra crop."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_keypoints(local: torch.Tensor, n_total: int | None = None) -> torch.Tensor:
    """all_gather of per-rank decoded keypoints ``[B_local, K, 3]`` (x, y, maxval) into ``[n_total, K, 3]``
    in rank order.  Shards may be ragged (see ``shard_range``): they are padded to the largest shard for the
    collective and trimmed afterwards.  Single-process groups return the input."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if n_total is None:
        n_total = local.shape[0] * world
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] < cap:
        pad = torch.zeros((cap - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous())
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """The ONE collective of the data-parallel training step: in-place mean over ranks of the flat
    gradient buffer (``buctd_b200.optim.Adam.flat_grad``: every parameter's gradient is a view into it).  Equal shard
    sizes make the mean of the per-rank JointsMSELoss gradients the global-batch gradient, which is what the
    reference's DataParallel reduce-add computes (tools/train.py:147).  No-op in a single-process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat
