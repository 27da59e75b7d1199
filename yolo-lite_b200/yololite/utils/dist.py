"""Multi-GPU plumbing of the inference path: one process per GPU, the image batch sharded contiguously, weights
replicated and NO collective on the hot path (images are independent).  The only exchanges are for the metric:
gathering the per-rank detections and taking the maximum over ranks of a device-measured time.

Works on any initialised `torch.distributed` backend: NCCL on the B200 box, gloo in the CPU tests
(tests/test_dist_gloo.py, world_size 2).  The reference has no distributed inference at all (its predictor is
single-device, engine/predictor.py:21-323); its validator only shards the dataloader under DDP training.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ("shard_range", "shard_batch", "gather_detections", "max_over_ranks", "sum_over_ranks")


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous share of `total` items; the first total % world ranks get one more."""
    if not (world >= 1 and 0 <= rank < world and total >= 0):
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch: torch.Tensor, group=None) -> torch.Tensor:
    """This rank's contiguous slice (a view) of a batch-major tensor."""
    world, rank = _world(group)
    s, e = shard_range(batch.shape[0], world, rank)
    return batch[s:e]


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, total: int | None = None, group=None):
    """All-gather the padded NMS output of every rank's shard, in global image order.

    dets (b_local, max_det, 6) fp32, counts (b_local,) int32 -> (dets (total, max_det, 6), counts (total,)) on
    every rank.  Shards may differ by one image (see shard_range): they are padded to the largest for the
    fixed-size collective and trimmed afterwards.  7.2 KB per image: metric plumbing, not data path."""
    world, rank = _world(group)
    if world == 1:
        return dets, counts
    sizes = [shard_range(total, world, r) for r in range(world)] if total is not None else None
    if sizes is None:
        n = torch.tensor([dets.shape[0]], dtype=torch.int64, device=dets.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        lens = [int(t.item()) for t in ns]
    else:
        lens = [e - s for s, e in sizes]
        if lens[rank] != dets.shape[0]:
            raise ValueError(f"rank {rank} holds {dets.shape[0]} images, its shard of {total} is {lens[rank]}")
    m = max(lens)
    pd = dets.new_zeros((m,) + tuple(dets.shape[1:]))
    pc = counts.new_zeros((m,))
    pd[: dets.shape[0]] = dets
    pc[: counts.shape[0]] = counts
    gd = [torch.empty_like(pd) for _ in range(world)]
    gc = [torch.empty_like(pc) for _ in range(world)]
    dist.all_gather(gd, pd.contiguous(), group=group)
    dist.all_gather(gc, pc.contiguous(), group=group)
    return (torch.cat([g[:n] for g, n in zip(gd, lens)]), torch.cat([g[:n] for g, n in zip(gc, lens)]))


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Maximum of a per-rank scalar (the device-measured step time: a multi-GPU number is the slowest rank's)."""
    world, _ = _world(group)
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: int, device=None, group=None) -> int:
    world, _ = _world(group)
    if world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
