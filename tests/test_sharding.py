"""CPU: the N>1 host logic (work partition + the single gather) under gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmd_gen_b200.sharding import assign_pockets, gather_point_clouds, sample_pockets, shard_pockets


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
    # a heavy pocket at the end must not leave three ranks idle (round-1 advisor finding)
    assert shard_pockets([1, 1, 1, 10], 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    tail = shard_pockets([1, 1, 1, 10], 2)
    assert max(sum([1, 1, 1, 10][a:b]) for a, b in tail) == 10


def test_assign_pockets_longest_first():
    costs = [3, 9, 2, 7, 7, 1, 4, 4, 8, 5]
    for world in (1, 2, 3, 4, 8):
        parts = assign_pockets(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        load = [sum(costs[i] for i in p) for p in parts]
        assert max(load) - min(load) <= max(costs)
        assert max(load) <= sum(costs) / world + max(costs)          # the LPT guarantee
    assert assign_pockets([1, 1, 1, 10], 2) == [[3], [0, 1, 2]]


class _StubDDPM:
    """CPU stand-in with the mirror's interface: a 'sample' is a deterministic function of (seed, global sample id,
    pocket) — like the device noise generator makes the real one — so sharding must not change any result."""
    n_dims, phar_nf = 3, 8

    def __init__(self):
        self.noise_seed, self.sample_ids = None, None
        emb = type("E", (), {"weight": torch.zeros(1)})
        self.dynamics = type("D", (), {"egnn": type("G", (), {"embedding": emb})})

    def sample_given_pocket(self, pocket, num_nodes_phar, timesteps=None):
        n = len(pocket["size"])
        phar_mask = torch.repeat_interleave(torch.arange(n), num_nodes_phar)
        com = torch.zeros(n, 3).index_add_(0, pocket["mask"], pocket["x"]) / pocket["size"][:, None]
        pocket_x = pocket["x"] - com[pocket["mask"]]                      # the sampler re-centres the pocket
        g = torch.Generator()
        rows = []
        for b in range(n):
            g.manual_seed(int(self.noise_seed) * 1000003 + int(self.sample_ids[b]))
            rows.append(torch.randn(int(num_nodes_phar[b]), 11, generator=g))
        xh = torch.cat(rows)
        xh[:, 3:] = torch.nn.functional.one_hot(xh[:, 3:].argmax(1), 8).float()
        return xh, torch.cat([pocket_x, pocket["one_hot"].float()], 1), phar_mask, pocket["mask"]


def _pocket_list():
    g = torch.Generator().manual_seed(5)
    sizes = [30, 12, 44, 25, 19, 37, 8]
    return [{"x": torch.randn(n, 3, generator=g) * 5 + 20, "one_hot": torch.nn.functional.one_hot(torch.randint(0, 20, (n,), generator=g), 20)}
            for n in sizes], [4, 6, 5, 8, 3, 7, 4]


def _pockets_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pockets, n_ph = _pocket_list()
    timing = {}
    out = sample_pockets(_StubDDPM(), pockets, 3, n_ph, seed=11, timing=timing)
    q.put((rank, [o.clone() for o in out], timing["pockets"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_pockets_is_independent_of_world_size():
    pockets, n_ph = _pocket_list()
    single = sample_pockets(_StubDDPM(), pockets, 3, n_ph, seed=11)
    assert [o.shape[0] for o in single] == [3 * k for k in n_ph]
    # back in the pocket's original frame: the stub's points are N(0, 1) around the re-centred pocket
    for o, p in zip(single, pockets):
        assert (o[:, :3].mean(0) - p["x"].mean(0)).abs().max() < 2.5
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pockets_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][2] + got[1][2] == len(pockets) and min(got[0][2], got[1][2]) >= 1
    for _, out, _ in got:                                   # every rank holds every pocket, identical to one rank alone
        for a, b in zip(out, single):
            assert torch.equal(a, b)


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
