"""CPU suite: host-side multi-GPU logic with world_size-2 gloo processes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from realtime_stereo_matcher_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from realtime_stereo_matcher_b200.sharding import all_gather_metrics, shard_range
    b, e = shard_range(9, rank, world)
    got = all_gather_metrics({"pairs": e - b, "ms": 10.0 * (rank + 1)})
    # the bench's max-over-ranks timing + sum of units
    t = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, got, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, got, tmax in res:
        assert got["pairs"] == [5.0, 4.0] and got["ms"] == [10.0, 20.0]
        assert tmax == 20.0


def test_all_gather_without_process_group():
    from realtime_stereo_matcher_b200.sharding import all_gather_metrics
    assert all_gather_metrics({"a": 1, "b": 2.5}) == {"a": [1.0], "b": [2.5]}
