"""Multi-GPU: (pocket, sample) work items are independent — every graph is
masked per sample (dynamics.py:143), COM removal is per sample (conditional_model.py:471).
So the work list is partitioned contiguously by pocket across ranks with NO collective in
the sampling loop; one all_gather of the padded point clouds afterwards.  The reference has
no multi-GPU sampling at all (lightning_modules.py:291-294 pins it to rank 0)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_pockets(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous partition [lo, hi) of pockets per rank with the smallest possible heaviest range (every rank gets
    a possibly empty range): bisection on the bottleneck load, greedy packing from the left."""
    n = len(costs)
    c = [float(v) for v in costs]

    def pack(limit):
        bounds, acc = [0], 0.0
        for i, v in enumerate(c):
            if acc > 0.0 and acc + v > limit:
                bounds.append(i)
                acc = 0.0
            acc += v
        return bounds

    lo, hi = (max(c) if c else 0.0), sum(c)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if len(pack(mid)) <= world_size:
            hi = mid
        else:
            lo = mid
    bounds = pack(hi * (1.0 + 1e-12) + 1e-12)
    # spare ranks: split the heaviest multi-pocket ranges further (cannot raise the bottleneck)
    while len(bounds) < world_size:
        spans = [(sum(c[a:b]), a, b) for a, b in zip(bounds, bounds[1:] + [n]) if b - a > 1]
        if not spans:
            break
        _, a, b = max(spans)
        half, acc, cut = 0.5 * sum(c[a:b]), 0.0, a + 1
        for i in range(a, b - 1):
            acc += c[i]
            cut = i + 1
            if acc >= half:
                break
        bounds.append(cut)
        bounds.sort()
    bounds = bounds + [n] * (world_size + 1 - len(bounds))
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def assign_pockets(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of pockets to ranks (pockets are independent work items, so a rank's
    set need not be contiguous): heaviest pocket first, always to the least loaded rank; ties by rank.  Each rank's
    list is returned in ascending pocket order.  Deterministic, identical on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return [sorted(v) for v in out]


def _all_gather_flat(local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """ONE fixed-size collective: [n] -> [world, n]."""
    out = torch.empty((world, local.numel()), device=local.device, dtype=local.dtype)
    try:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    except (RuntimeError, NotImplementedError):            # a backend without the flat variant
        parts = [torch.empty_like(local.view(-1)) for _ in range(world)]
        dist.all_gather(parts, local.contiguous().view(-1), group=group)
        out = torch.stack(parts)
    return out


def gather_point_clouds(xh_phar: torch.Tensor, counts: torch.Tensor, group=None,
                        max_points: Optional[int] = None, max_samples: Optional[int] = None, uniform: bool = False):
    """all_gather of per-rank results: xh_phar [n_points, W] with per-sample point counts [n_samples].  Returns
    (xh_all [sum points, W], counts_all) in rank order on every rank.

    Points and counts travel in ONE collective (counts ride as floats behind the padded points; exact below 2^24).
    With ``max_points`` / ``max_samples`` given (the caller knows the layout of every rank — bench.py, sample_pockets)
    there is no size exchange and no host synchronisation before the collective; without them one small all_gather
    of the two sizes comes first.  ``uniform=True`` asserts every rank holds exactly max_points / max_samples (weak
    scaling with identical batches): the result is a pure view of the gathered buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xh_phar, counts
    world = dist.get_world_size(group)
    dev = xh_phar.device
    W = xh_phar.shape[1]
    if max_points is None or max_samples is None:
        sizes = torch.tensor([xh_phar.shape[0], counts.shape[0]], device=dev, dtype=torch.int64)
        all_sizes = _all_gather_flat(sizes, world, group).cpu()
        max_points, max_samples = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    buf = torch.zeros(max_points * W + max_samples + 1, device=dev, dtype=torch.float32)
    buf[: xh_phar.numel()] = xh_phar.reshape(-1)
    buf[max_points * W: max_points * W + counts.shape[0]] = counts.to(dev, torch.float32)
    buf[-1] = float(counts.shape[0])
    allb = _all_gather_flat(buf, world, group)                          # [world, max_points * W + max_samples + 1]
    pts = allb[:, : max_points * W].reshape(world, max_points, W)
    cnt = allb[:, max_points * W: -1].round().to(torch.int64)           # [world, max_samples]
    if uniform:
        return pts.reshape(world * max_points, W), cnt.reshape(-1)      # same layout on every rank: pure views, no host sync
    n_smp = allb[:, -1].round().to(torch.int64).tolist()
    n_pts = [int(cnt[r, : n_smp[r]].sum()) for r in range(world)]
    xh_all = torch.cat([pts[r, : n_pts[r]] for r in range(world)])
    counts_all = torch.cat([cnt[r, : n_smp[r]] for r in range(world)])
    return xh_all, counts_all


def pocket_cost(n_res: int, n_phar: int, n_samples: int, cutoff_degree: float = 7.0) -> float:
    """Work estimate of sampling one pocket: edges per denoiser call ~ degree x nodes at a fixed cutoff."""
    return float(n_samples) * float(n_res + n_phar) * cutoff_degree


@torch.no_grad()
def sample_pockets(ddpm, pockets: Sequence[dict], n_samples: int, num_nodes_phar, seed: int = 0,
                   timesteps: Optional[int] = None, group=None, timing: Optional[dict] = None):
    """BASELINE config 4: a LIST of pockets, ``n_samples`` point clouds each, sharded over the ranks of ``group``.

    The reference has no such driver (generate_phars.py handles one pocket per process, lightning_modules.py:291-294
    pins sampling to rank 0; test.py:83-100 loops over pockets serially); per pocket this does exactly what
    ``PharPocketDDPM.generate_phars`` does (lightning_modules.py:443-504): replicate the pocket, sample_given_pocket,
    translate back into the pocket's original frame.

      * pockets are whole work items, assigned longest-first to the least loaded rank (``assign_pockets``);
      * no collective inside the sampling loop; each rank re-plans its ONE handle per pocket (grow-only workspace,
        the step graph is re-captured only when the layout changes);
      * noise comes from the device generator keyed by (seed, pocket index * n_samples + sample index): the result of a
        pocket does not depend on the number of ranks;
      * ONE all_gather at the end, sized from the (globally known) layout: no size exchange, no host sync before it.

    pockets[i] = {'x': [n_r, 3] float, 'one_hot': [n_r, residue_nf]} on any device.  num_nodes_phar: int or one int per
    pocket.  Returns, on every rank, a list with one [n_samples * n_p_i, 3 + phar_nf] tensor per pocket (coordinates
    in the pocket's original frame | one-hot type).  ``timing`` (dict) receives per-rank loop seconds."""
    import time
    from .utils import scatter_mean
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = ddpm.dynamics.egnn.embedding.weight.device
    n_pockets = len(pockets)
    n_ph = [int(num_nodes_phar)] * n_pockets if isinstance(num_nodes_phar, int) else [int(v) for v in num_nodes_phar]
    n_res = [int(p["x"].shape[0]) for p in pockets]
    W = ddpm.n_dims + ddpm.phar_nf
    costs = [pocket_cost(n_res[i], n_ph[i], n_samples) for i in range(n_pockets)]
    mine = assign_pockets(costs, world)
    pts_of = [n_samples * n_ph[i] for i in range(n_pockets)]
    rank_pts = [sum(pts_of[i] for i in mine[r]) for r in range(world)]
    max_pts = max(rank_pts) if rank_pts else 0
    local = torch.zeros((max_pts, W), device=dev, dtype=torch.float32)
    saved = (ddpm.noise_seed, ddpm.sample_ids)
    t0 = time.perf_counter()
    off = 0
    try:
        for i in mine[rank]:
            x = pockets[i]["x"].to(dev, torch.float32)
            one_hot = pockets[i]["one_hot"].to(dev)
            pocket = {
                "x": x.repeat(n_samples, 1), "one_hot": one_hot.repeat(n_samples, 1),
                "size": torch.full((n_samples,), n_res[i], device=dev, dtype=torch.int64),
                "mask": torch.repeat_interleave(torch.arange(n_samples, device=dev), n_res[i]),
            }
            com_before = scatter_mean(pocket["x"], pocket["mask"])
            ddpm.noise_seed = int(seed)
            ddpm.sample_ids = torch.arange(i * n_samples, (i + 1) * n_samples, dtype=torch.int64)
            xh_phar, xh_pocket, phar_mask, pocket_mask = ddpm.sample_given_pocket(
                pocket, torch.full((n_samples,), n_ph[i], dtype=torch.int64), timesteps=timesteps)
            shift = com_before - scatter_mean(xh_pocket[:, :ddpm.n_dims], pocket_mask)     # lightning_modules.py:495-504
            xh_phar[:, :ddpm.n_dims] += shift[phar_mask]
            local[off: off + pts_of[i]] = xh_phar
            off += pts_of[i]
    finally:
        ddpm.noise_seed, ddpm.sample_ids = saved
    if timing is not None:
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        timing["loop_s"] = time.perf_counter() - t0
        timing["pockets"] = len(mine[rank])
        timing["cost_share"] = sum(costs[i] for i in mine[rank]) / max(sum(costs), 1e-30)
    allp = _all_gather_flat(local.view(-1), world, group).view(world, max_pts, W) if world > 1 else local.view(1, max_pts, W)
    out = [None] * n_pockets
    for r in range(world):
        off = 0
        for i in mine[r]:
            out[i] = allp[r, off: off + pts_of[i]]
            off += pts_of[i]
    return out
