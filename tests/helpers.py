"""Shared test helpers: golden loading and case configs (mirrors oracle/make_golden.py CASES)."""
import os

import numpy as np
import torch

from cmd_gen_b200.config import DynamicsConfig
from cmd_gen_b200.synthetic import CA_DENSITY, FULL_ATOM_DENSITY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASE_CFG = {
    "ca_small": dict(),
    "fa_small": dict(residue_nf=11, n_layers=3, inv_sublayers=2),
    "nocut": dict(n_layers=2, edge_cutoff=None, attention=False, tanh=False, norm_constant=0.0),
    "mean_agg": dict(n_layers=2, aggregation_method="mean", condition_time=False),
}


def case_config(name):
    return DynamicsConfig(**CASE_CFG[name])


def load(name):
    with np.load(os.path.join(GOLDEN, name)) as f:
        return {k: f[k] for k in f.files}


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)
