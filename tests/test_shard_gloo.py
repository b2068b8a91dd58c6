"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: window partitioning and the max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cdvslam_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.windows_of_rank(64, world, rank)
    ms = 10.0 + 5.0 * rank                                  # rank 1 is slower
    thr, worst = shard.aggregate_throughput(2 * len(mine), ms, "cpu", dist)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    if rank == 0:
        out.put((thr, worst, everyone))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_and_max_over_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    thr, worst, everyone = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst == 15.0                                    # max over ranks, not the mean
    assert abs(thr - (2 * 64) / 15e-3) < 1e-6               # all windows / slowest rank
    assert sorted(everyone[0] + everyone[1]) == list(range(64))
    assert not set(everyone[0]) & set(everyone[1])


def test_partition_covers_all_windows():
    for world in (1, 2, 4, 8):
        got = sorted(w for r in range(world) for w in shard.windows_of_rank(64, world, r))
        assert got == list(range(64))
        sizes = [len(shard.windows_of_rank(64, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
