"""TEST INFRASTRUCTURE — seeded inputs of the BASELINE-config-sized parity cases.

Shared by ``oracle/make_golden_large.py`` (which runs the UNMODIFIED reference on them in the
build container) and by the tests that replay the committed fixtures on the GPU box, so that
the multi-megabyte inputs need not be stored: a fixture holds the reference OUTPUTS plus
float64 checksums of the regenerated inputs (torch's CPU generators are deterministic for a
given torch build; the checksum catches a drift).

Cases (BASELINE.json ``configs``):
  config1 : one Calpha pocket (N_r = 150) x 10 samples, N_p = 10, all 500 steps   (configs[0])
  config2 : one Calpha pocket (N_r = 150) x 64 samples, N_p = 8 — bench.py's workload (configs[1])
  config3 : 8 distinct full-atom pockets of ~2 000 nodes, ragged N_p (configs[2]; N = 16 k: cell-list
            graph builder, lane-range segmented sum, more than one wave of node tiles)
  config5 : 2 pockets of 4 000 nodes, N_p = 12, n_layers = 9 (configs[4])
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from cmd_gen_b200.config import DynamicsConfig
from cmd_gen_b200.synthetic import CA_DENSITY, FULL_ATOM_DENSITY, draw_noise, make_pocket_batch
from oracle import diffphar_oracle as orc

DYNAMICS_CASES = {
    # name: cfg kwargs, pocket sizes, replicate, phar counts, density, pocket seed, weight seed, t values
    "config2": dict(cfg=dict(), sizes=[150], replicate=64, counts=[8] * 64, density=CA_DENSITY, pseed=1, wseed=0,
                    t_values=[1.0, 0.5, 0.002, 0.0]),
    "config3": dict(cfg=dict(residue_nf=11), sizes=[2000, 1873, 2100, 1950, 2048, 1999, 2200, 1777], replicate=1,
                    counts=[8, 4, 12, 7, 9, 5, 10, 6], density=FULL_ATOM_DENSITY, pseed=31, wseed=0,
                    t_values=[0.5, 0.002]),
    "config5": dict(cfg=dict(residue_nf=11, n_layers=9), sizes=[4000, 3900], replicate=1, counts=[12, 12],
                    density=FULL_ATOM_DENSITY, pseed=51, wseed=0, t_values=[0.3]),
}

SAMPLER_CASES = {
    # the reference's own CPU-runnable case: generate_phars on one pocket, 10 samples (BASELINE configs[0])
    "config1": dict(cfg=dict(), sizes=[150], replicate=10, counts=[10] * 10, density=CA_DENSITY, pseed=1, wseed=0,
                    T=500, noise_seed=123, trace_every=50),
}


def checksum(t: torch.Tensor) -> np.ndarray:
    """(sum, sum of |v|, sum of v * position weight) in float64: order-sensitive, cheap, exact for equal inputs."""
    v = t.detach().cpu().to(torch.float64).reshape(-1)
    w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 251.0
    return np.array([float(v.sum()), float(v.abs().sum()), float((v * w).sum())])


def edges_digest(edges: np.ndarray) -> str:
    """sha256 of the (row, col) int64 edge list in the reference's order."""
    return hashlib.sha256(np.ascontiguousarray(edges.astype(np.int64)).tobytes()).hexdigest()


def spread_state(cfg, pocket, counts, seed, spread=4.0):
    """A z_t-like state: phar points spread around each pocket's centre (so phar-residue edges exist), COM-free,
    pocket translated with it — built with the oracle's own noise_and_center."""
    B = len(counts)
    counts_t = torch.tensor(counts, dtype=torch.int64)
    mask_p = torch.repeat_interleave(torch.arange(B), counts_t)
    px = pocket["x"].clone()
    ph = pocket["one_hot"].float() / 4.0
    xh0 = torch.cat([px, ph], 1)
    mu_x = orc._scatter_mean(px, pocket["mask"], B)
    mu = torch.cat([mu_x, torch.zeros(B, cfg.phar_nf)], 1)[mask_p]
    eps = draw_noise(1, int(counts_t.sum()), cfg.n_dims + cfg.phar_nf, seed=seed)[0]
    eps[:, :3] *= spread
    z, xh_pocket = orc.noise_and_center(mu, xh0, torch.ones(()), eps, mask_p, pocket["mask"], B)
    return z, xh_pocket, mask_p, counts_t


def dynamics_inputs(name):
    c = DYNAMICS_CASES[name]
    cfg = DynamicsConfig(**c["cfg"])
    pocket = make_pocket_batch(c["sizes"], cfg.residue_nf, density=c["density"], seed=c["pseed"], replicate=c["replicate"])
    sizes = [int(s) for s in pocket["size"]]
    z, xh_pocket, mask_p, counts_t = spread_state(cfg, pocket, c["counts"], seed=c["pseed"] + 100,
                                                  spread=4.0 if c["density"] == CA_DENSITY else 6.0)
    return dict(cfg=cfg, z=z, xh_pocket=xh_pocket, mask_phar=mask_p, mask_res=pocket["mask"], sizes=sizes,
                counts=list(c["counts"]), wseed=c["wseed"], t_values=list(c["t_values"]))


def sampler_inputs(name):
    c = SAMPLER_CASES[name]
    cfg = DynamicsConfig(**c["cfg"])
    pocket = make_pocket_batch(c["sizes"], cfg.residue_nf, density=c["density"], seed=c["pseed"], replicate=c["replicate"])
    counts = list(c["counts"])
    noise = draw_noise(c["T"] + 2, sum(counts), cfg.n_dims + cfg.phar_nf, seed=c["noise_seed"])
    return dict(cfg=cfg, pocket=pocket, counts=counts, noise=noise, wseed=c["wseed"], T=c["T"],
                trace_every=c["trace_every"])
