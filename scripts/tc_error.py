"""GPU tool: (1) actual errors of the 16-bit tensor-core modes against the fp64 goldens and against the fp32 mode
at config-2 size, next to the test tolerances; (2) K1 timing, scan vs cell list, at CA and full-atom sizes.
    python scripts/tc_error.py"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmd_gen_b200 import _lib
from cmd_gen_b200.config import DynamicsConfig
from cmd_gen_b200.synthetic import make_pocket_batch
from cmd_gen_b200.weights import init_weights, pack_blob
from tests.helpers import case_config, load, T

DEV = "cuda:0"


def handle(cfg, seed, prec, graph=None):
    os.environ.pop("DIFFPHAR_GRAPH", None)
    if graph:
        os.environ["DIFFPHAR_GRAPH"] = graph
    h = _lib.Handle(cfg, DEV, prec)
    os.environ.pop("DIFFPHAR_GRAPH", None)
    h.set_weights(pack_blob(cfg, init_weights(cfg, seed)))
    return h


print("== denoiser error / max|ref| per mode (test bound: f16 1e-4, bf16 1e-3)")
for name in ["ca_small", "fa_small", "nocut", "mean_agg"]:
    g = load(f"dynamics_{name}.npz")
    cfg = case_config(name)
    B = len(g["sizes"])
    for prec in ("f16", "f16fast", "f16fast32", "bf16"):
        h = handle(cfg, int(g["wseed"]), prec)
        h.plan(g["counts"], g["sizes"])
        worst_h, worst_x = 0.0, 0.0
        for i, tv in enumerate(g["t_values"]):
            out_p, out_r = h.dynamics_forward(T(g["z"]), T(g["xh_pocket"]), torch.full((B,), float(tv)))
            out_p, out_r = out_p.cpu().numpy(), out_r.cpu().numpy()
            rp, rr = g[f"eps_phar_f64_{i}"], g[f"eps_res_f64_{i}"]
            worst_h = max(worst_h, np.abs(out_p[:, 3:] - rp[:, 3:]).max() / max(1.0, np.abs(rp[:, 3:]).max()),
                          np.abs(out_r[:, 3:] - rr[:, 3:]).max() / max(1.0, np.abs(rr[:, 3:]).max()))
            worst_x = max(worst_x, np.abs(out_p[:, :3] - rp[:, :3]).max() / max(1e-12, np.abs(rp[:, :3]).max()))
        print(f"  {name:9s} {prec:9s} features {worst_h:.2e}   velocity (rel. to max|vel|) {worst_x:.2e}")

cfg = DynamicsConfig()
B, n_res, n_ph = 64, 150, 8
pocket = make_pocket_batch([n_res], 20, seed=3, replicate=B)
gen = torch.Generator().manual_seed(4)
com = pocket["x"][:n_res].mean(0)
z = torch.cat([com + 5.0 * torch.randn(B * n_ph, 3, generator=gen), torch.randn(B * n_ph, 8, generator=gen)], 1)
xr = torch.cat([pocket["x"], pocket["one_hot"].float() / 4], 1)
t = torch.full((B,), 0.4)
outs = {}
for prec in ("fp32", "f16", "f16fast", "f16fast32", "bf16"):
    h = handle(cfg, 0, prec)
    h.plan([n_ph] * B, [n_res] * B)
    a, r = h.dynamics_forward(z, xr, t)
    outs[prec] = (a.cpu(), r.cpu())
for prec in ("f16", "f16fast", "f16fast32", "bf16"):
    a, r = outs[prec]
    rp, rr = outs["fp32"]
    print(f"  config2   {prec:9s} features {float(max((a[:, 3:] - rp[:, 3:]).abs().max() / max(1.0, float(rp[:, 3:].abs().max())), (r[:, 3:] - rr[:, 3:]).abs().max() / max(1.0, float(rr[:, 3:].abs().max())))):.2e}"
          f"   velocity {float((a[:, :3] - rp[:, :3]).abs().max() / rp[:, :3].abs().max()):.2e}")

print("== K1 radius graph: scan vs cell list (CUDA events, 20 builds)")
for label, n_res, n_ph, B, dens in [("config2 CA 150 x 64", 150, 8, 64, 0.0074), ("config3 full-atom 2000 x 16", 2000, 8, 16, 0.05),
                                    ("config3 full-atom 2000 x 64", 2000, 8, 64, 0.05), ("config5 4000 x 8", 4000, 12, 8, 0.05)]:
    pocket = make_pocket_batch([n_res], 20, density=dens, seed=3, replicate=B)
    gen = torch.Generator().manual_seed(4)
    com = pocket["x"][:n_res].mean(0)
    x = torch.cat([com + 5.0 * torch.randn(B * n_ph, 3, generator=gen), pocket["x"]]).to(DEV)
    line = f"  {label:28s}"
    for graph in ("scan", "cells"):
        h = handle(DynamicsConfig(n_layers=1), 0, "fp32", graph)
        h.plan([n_ph] * B, [n_res] * B)
        for _ in range(3):
            h.build_edges(x)
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            _lib._check(h.lib.dp_build_edges(h.h, _lib._ptr(x), st))
        e1.record()
        torch.cuda.synchronize()
        E = h.flags().last_n_edges
        us = e0.elapsed_time(e1) * 1e3 / 20
        bytes_alg = 16 * x.shape[0] + 16 * E + 4 * (x.shape[0] + 1)
        line += f"  {graph}: {us:8.1f} us ({bytes_alg / us / 1e3:6.1f} GB/s alg.)"
    print(line + f"   N={x.shape[0]} E={E}")
