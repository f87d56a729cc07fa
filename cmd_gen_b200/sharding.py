"""Multi-GPU: (pocket, sample) work items are independent — every graph is
masked per sample (dynamics.py:143), COM removal is per sample (conditional_model.py:471).
So the work list is partitioned contiguously by pocket across ranks with NO collective in
the sampling loop; one all_gather of the padded point clouds afterwards.  The reference has
no multi-GPU sampling at all (lightning_modules.py:291-294 pins it to rank 0)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_pockets(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous partition [lo, hi) of pockets per rank, balanced by cumulative cost
    (e.g. expected edge count x samples).  Every rank gets a (possibly empty) range."""
    n = len(costs)
    total = float(sum(costs))
    bounds, acc, k = [0], 0.0, 1
    for i, c in enumerate(costs):
        acc += float(c)
        while k < world_size and acc >= total * k / world_size - 1e-12 and len(bounds) < world_size:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world_size:
        bounds.append(n)
    bounds.append(n)
    return [(bounds[r], max(bounds[r], bounds[r + 1])) for r in range(world_size)]


def gather_point_clouds(xh_phar: torch.Tensor, counts: torch.Tensor, group=None):
    """all_gather of ragged per-rank results: xh_phar [n_points, W] with per-sample point
    counts [n_samples].  Returns (xh_all [sum points, W], counts_all) in rank order on every
    rank.  Padded to the max over ranks so one fixed-size collective suffices."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xh_phar, counts
    world = dist.get_world_size(group)
    dev = xh_phar.device
    sizes = torch.tensor([xh_phar.shape[0], counts.shape[0]], device=dev, dtype=torch.int64)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_pts = int(max(int(s[0]) for s in all_sizes))
    max_smp = int(max(int(s[1]) for s in all_sizes))
    pad_x = torch.zeros((max_pts, xh_phar.shape[1]), device=dev, dtype=xh_phar.dtype)
    pad_x[: xh_phar.shape[0]] = xh_phar
    pad_c = torch.zeros((max_smp,), device=dev, dtype=torch.int64)
    pad_c[: counts.shape[0]] = counts.to(dev, torch.int64)
    xs = [torch.empty_like(pad_x) for _ in range(world)]
    cs = [torch.empty_like(pad_c) for _ in range(world)]
    dist.all_gather(xs, pad_x, group=group)
    dist.all_gather(cs, pad_c, group=group)
    xh_all = torch.cat([x[: int(s[0])] for x, s in zip(xs, all_sizes)])
    counts_all = torch.cat([c[: int(s[1])] for c, s in zip(cs, all_sizes)])
    return xh_all, counts_all
