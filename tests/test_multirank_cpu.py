"""N > 1 host logic on CPU: two gloo ranks shard the streams and reduce timing/work the way bench.py
does under torchrun (the data path itself has no collective: streams are independent)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_streams(2050, world, rank)
    # every rank "decodes" its own streams: 8 frames per stream and step, rank 1 is slower
    ms = [10.0 + 5.0 * rank, 20.0 - rank]
    frames = [8 * (hi - lo), 4 * (hi - lo)]
    t, f = bench.reduce_over_ranks(ms, frames, torch.device("cpu"), world)
    q.put((rank, lo, hi, t, f))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, f0), (r1, lo1, hi1, t1, f1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 1025, 1025, 2050)          # contiguous, disjoint, complete
    assert t0 == t1 == [15.0, 20.0]                               # max over ranks
    assert f0 == f1 == [8 * 2050.0, 4 * 2050.0]                   # whole-job work


def test_shards_cover_everything():
    sys.path.insert(0, ROOT)
    import bench
    for total, world in [(8192, 8), (1000, 3), (7, 8)]:
        spans = [bench.shard_streams(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
