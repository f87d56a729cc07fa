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


def write_synthetic_pdb(path, n_res: int = 60, seed: int = 7, ligand_resseq: int = 901, chain: str = "A") -> int:
    """A small fake protein (backbone + CB per residue, C-alpha-like density) with one HETATM ligand at its
    centre — enough for the generate_phars CLI / pocket-selection path.  Returns the number of residues."""
    from .constants import THREE_TO_ONE
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    ca = ball_points(n_res, CA_DENSITY, gen) + torch.tensor([12.0, -7.5, 30.25])
    names = sorted(THREE_TO_ONE)
    types = torch.randint(0, len(names), (n_res,), generator=gen).tolist()
    offs = {"N": (-1.2, 0.4, 0.1), "CA": (0.0, 0.0, 0.0), "C": (1.1, 0.7, -0.2), "O": (1.6, 1.7, 0.3), "CB": (0.1, -1.2, 0.9)}
    elem = {"N": "N", "CA": "C", "C": "C", "O": "O", "CB": "C"}
    lines, serial = [], 1
    for i in range(n_res):
        for an, d in offs.items():
            x, y, z = (ca[i] + torch.tensor(d)).tolist()
            lines.append("ATOM  %5d %-4s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s" %
                         (serial, " " + an if len(an) < 4 else an, names[types[i]], chain, i + 1, x, y, z, 1.0, 20.0, elem[an]))
            serial += 1
    com = ca.mean(0)
    for k, d in enumerate([(0.0, 0.0, 0.0), (1.4, 0.2, -0.3), (-0.8, 1.1, 0.6), (0.3, -1.3, 1.0)]):
        x, y, z = (com + torch.tensor(d)).tolist()
        lines.append("HETATM%5d %-4s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s" %
                     (serial, " C%d" % (k + 1), "LIG", chain, ligand_resseq, x, y, z, 1.0, 30.0, "C"))
        serial += 1
    lines.append("END")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return n_res
