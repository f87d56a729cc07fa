"""CPU: the N>1 host logic (work partition + the single gather) under gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmd_gen_b200.sharding import gather_point_clouds, shard_pockets


def test_shard_pockets_partitions_contiguously():
    costs = [5, 1, 1, 1, 4, 4, 2, 2]
    for world in (1, 2, 3, 4, 8, 16):
        parts = shard_pockets(costs, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and parts[-1][1] == len(costs)
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c and a <= b
    two = shard_pockets(costs, 2)
    load = [sum(costs[a:b]) for a, b in two]
    assert abs(load[0] - load[1]) <= max(costs)
    assert shard_pockets([], 2) == [(0, 0), (0, 0)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_samples = 2 + rank                               # ragged across ranks
    counts = torch.arange(1, n_samples + 1) + rank
    x = torch.arange(int(counts.sum()) * 11, dtype=torch.float32).reshape(-1, 11) + 1000 * rank
    xs, cs = gather_point_clouds(x, counts)
    q.put((rank, xs.clone(), cs.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_point_clouds_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c0 = torch.arange(1, 3); c1 = torch.arange(1, 4) + 1
    x0 = torch.arange(int(c0.sum()) * 11, dtype=torch.float32).reshape(-1, 11)
    x1 = torch.arange(int(c1.sum()) * 11, dtype=torch.float32).reshape(-1, 11) + 1000
    for _, xs, cs in got:                                # identical on every rank, rank order
        assert torch.equal(cs, torch.cat([c0, c1]))
        assert torch.equal(xs, torch.cat([x0, x1]))
