"""TEST INFRASTRUCTURE — joint-mode fixtures: the UNMODIFIED reference EGNNDynamics with
``update_pocket_coords=True`` (dynamics.py:104-107, 133-136: no coordinate mask, velocities of the pocket nodes returned,
per-sample mean removed over all nodes), run in this container in fp32 and fp64 on the inputs of make_golden.py's cases.

    python -m oracle.make_golden_joint      ->  tests/golden/dynamics_joint_<case>.npz
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmd_gen_b200.config import DynamicsConfig          # noqa: E402
from cmd_gen_b200.weights import init_weights           # noqa: E402
from cmd_gen_b200.synthetic import make_pocket_batch    # noqa: E402
from oracle import ref_shims                            # noqa: E402
from oracle.make_golden import CASES, OUT, initial_state  # noqa: E402

JOINT_CASES = ("ca_small", "fa_small", "mean_agg")


def reference_dynamics(cfg, state, dtype):
    EGNNDynamics, _ = ref_shims.load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        dyn = EGNNDynamics(
            phar_nf=cfg.phar_nf, residue_nf=cfg.residue_nf, n_dims=cfg.n_dims, joint_nf=cfg.joint_nf,
            hidden_nf=cfg.hidden_nf, device="cpu", act_fn=torch.nn.SiLU(), n_layers=cfg.n_layers,
            attention=cfg.attention, condition_time=cfg.condition_time, tanh=cfg.tanh, mode="egnn_dynamics",
            norm_constant=cfg.norm_constant, inv_sublayers=cfg.inv_sublayers, sin_embedding=False,
            normalization_factor=cfg.normalization_factor, aggregation_method=cfg.aggregation_method,
            update_pocket_coords=True, edge_cutoff=cfg.edge_cutoff)
        dyn.load_state_dict({k: v.clone() for k, v in state.items()}, strict=True)
    return (dyn.double() if dtype == torch.float64 else dyn).eval()


def gen(name):
    kw, sizes, counts, density, wseed = CASES[name]
    cfg = DynamicsConfig(**kw)
    state = init_weights(cfg, seed=wseed)
    pocket = make_pocket_batch(sizes, cfg.residue_nf, density=density, seed=11)
    z, xh_pocket, mask_p, _ = initial_state(cfg, pocket, counts, seed=7)
    mask_r = pocket["mask"]
    B = len(sizes)
    out = dict(z=z.numpy(), xh_pocket=xh_pocket.numpy(), mask_phar=mask_p.numpy(), mask_res=mask_r.numpy(),
               sizes=np.array(sizes), counts=np.array(counts), wseed=wseed)
    d32, d64 = reference_dynamics(cfg, state, torch.float32), reference_dynamics(cfg, state, torch.float64)
    ts = [1.0, 0.5, 0.002, 0.0]
    out["t_values"] = np.array(ts, dtype=np.float32)
    with torch.no_grad():
        for i, tv in enumerate(ts):
            t = torch.full((B, 1), tv, dtype=torch.float32)
            a, b = d32(z, xh_pocket, t, mask_p, mask_r)
            out[f"eps_phar_f32_{i}"], out[f"eps_res_f32_{i}"] = a.numpy(), b.numpy()
            a, b = d64(z.double(), xh_pocket.double(), t, mask_p, mask_r)
            out[f"eps_phar_f64_{i}"], out[f"eps_res_f64_{i}"] = a.numpy(), b.numpy()
    np.savez_compressed(os.path.join(OUT, f"dynamics_joint_{name}.npz"), **out)
    print(name, "pocket |vel| max", float(np.abs(out["eps_res_f64_1"][:, :3]).max()),
          "phar |vel| max", float(np.abs(out["eps_phar_f64_1"][:, :3]).max()))


if __name__ == "__main__":
    for n in JOINT_CASES:
        gen(n)
