"""N > 1 host logic on CPU: walker sharding and the accumulator all-reduce over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

import kagomedsl.jl_b200 as kd


def test_shard_walkers_partition():
    for total, world in ((32768, 8), (4097, 4), (10, 3), (1, 2)):
        spans = [kd.dist.shard_walkers(total, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == total
        pos = 0
        for first, c in spans:
            assert first == pos
            pos += c
    # disjoint RNG streams across ranks
    a = kd.walker_states(1234, 4, first_walker=0)
    b = kd.walker_states(1234, 4, first_walker=4)
    assert not np.intersect1d(a.ravel(), b.ravel()).size


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = kd.dist.shard_walkers(10, rank, world)
    vec = np.array([count * 216.0, 0.128 * count, -185.0 * count, 185.0 ** 2 * count, count, 0, 0, 0])
    out = kd.dist.allreduce_sum(vec)
    q.put((rank, out.tolist(), first, count))
    dist.destroy_process_group()


def test_accumulator_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1]                         # both ranks hold the same global sums
    assert res[0][1][0] == 10 * 216.0 and res[0][1][4] == 10
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 5, 5, 5)
