"""Seeded synthetic CrossDocked-shaped pockets.

The reference ships no data, so every workload is synthetic: pocket nodes are
uniform in a ball of a given number density (0.0074 A^-3 ~ C-alpha spacing,
0.05 A^-3 ~ heavy atoms), shifted by a fixed offset to mimic a PDB frame, with
uniform random residue/atom types.  The dict layout is what
``PharPocketDDPM.generate_phars`` builds (lightning_modules.py:443-455).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

CA_DENSITY = 0.0074
FULL_ATOM_DENSITY = 0.05


def ball_points(n: int, density: float, gen: torch.Generator) -> torch.Tensor:
    radius = (3.0 * n / (4.0 * math.pi * density)) ** (1.0 / 3.0)
    d = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    r = radius * torch.rand(n, 1, generator=gen, dtype=torch.float64) ** (1.0 / 3.0)
    return (d * r).to(torch.float32)


def make_pocket_batch(sizes: Sequence[int], residue_nf: int, density: float = CA_DENSITY,
                      seed: int = 1, offset=(12.0, -7.5, 30.25), replicate: int = 1) -> Dict[str, torch.Tensor]:
    """One pocket per entry of ``sizes``; each repeated ``replicate`` times
    consecutively (the generate_phars layout when len(sizes) == 1)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    off = torch.tensor(offset, dtype=torch.float32)
    xs, hs, size, mask = [], [], [], []
    b = 0
    for n in sizes:
        x = ball_points(int(n), density, gen) + off
        types = torch.randint(0, residue_nf, (int(n),), generator=gen)
        one_hot = torch.nn.functional.one_hot(types, residue_nf)
        for _ in range(replicate):
            xs.append(x.clone())
            hs.append(one_hot.clone())
            size.append(int(n))
            mask.append(torch.full((int(n),), b, dtype=torch.int64))
            b += 1
    return {"x": torch.cat(xs), "one_hot": torch.cat(hs),
            "size": torch.tensor(size, dtype=torch.int64), "mask": torch.cat(mask)}


def draw_noise(n_draws: int, n_phar_nodes: int, width: int, seed: int = 123) -> torch.Tensor:
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    return torch.randn(n_draws, n_phar_nodes, width, generator=gen, dtype=torch.float32)
