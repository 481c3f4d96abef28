"""Multi-GPU plumbing on CPU: contiguous frame sharding and the max-over-ranks timing reduction, run as a real
world_size-2 gloo job (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyorbslam_b200.sharding import owner_of, shard_range


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8), (0, 2), (4097, 8)])
def test_ranges_partition_the_frames(n, world):
    seen = []
    for r in range(world):
        a, b = shard_range(n, r, world)
        assert 0 <= a <= b <= n
        seen += list(range(a, b))
        assert all(owner_of(i, n, world) == r for i in range(a, b))
    assert seen == list(range(n))
    sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(101, rank, world)
    t = torch.tensor([float(10 + 5 * rank)], dtype=torch.float64)     # pretend per-rank elapsed ms
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([b - a], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, float(t), int(cnt)))
    dist.destroy_process_group()


def test_gloo_world2_max_time_and_total_units():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [o[1] for o in out] == [15.0, 15.0]        # max over ranks
    assert [o[2] for o in out] == [101, 101]          # every frame owned exactly once
