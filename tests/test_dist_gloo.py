"""N>1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding and the metric-only
collectives of yololite.utils.dist (shards are independent, no collective on the data path)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]

from yololite.utils.dist import (gather_detections, max_over_ranks, shard_batch, shard_range,  # noqa: E402
                                 sum_over_ranks)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            lens = [e - s for s, e in spans]
            assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _fake_nms(batch, max_det=300):
    """Deterministic stand-in for the per-image result of the GPU path: depends on the image content only."""
    b = batch.shape[0]
    dets = torch.zeros((b, max_det, 6), dtype=torch.float32)
    counts = torch.zeros((b,), dtype=torch.int32)
    for i in range(b):
        n = int(batch[i].sum().item() * 1000) % (max_det + 1)
        counts[i] = n
        dets[i, :n] = batch[i].flatten()[:6] + torch.arange(n, dtype=torch.float32)[:, None]
    return dets, counts


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        batch = torch.rand(total, 3, 4, 4, generator=g)          # every rank sees the same global batch
        mine = shard_batch(batch)
        s, e = shard_range(total, world, rank)
        assert mine.shape[0] == e - s and torch.equal(mine, batch[s:e])
        dets, counts = _fake_nms(mine)
        gd, gc = gather_detections(dets, counts, total=total)
        gd2, gc2 = gather_detections(dets, counts)                # sizes discovered by an all-gather
        assert torch.equal(gd, gd2) and torch.equal(gc, gc2)
        t = max_over_ranks(1.0 + rank)                            # the slowest rank's time is the job's
        n = sum_over_ranks(int(counts.sum()))
        q.put((rank, gd.numpy(), gc.numpy(), t, n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_run_equals_single_process(total):
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    ref_d, ref_c = _fake_nms(torch.rand(total, 3, 4, 4, generator=g))
    for rank, gd, gc, t, n in got:
        assert np.array_equal(gd, ref_d.numpy()) and np.array_equal(gc, ref_c.numpy()), f"rank {rank}"
        assert t == float(world) and n == int(ref_c.sum())
