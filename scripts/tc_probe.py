"""GPU probe: errors of each precision mode / TC mask against the fp64 reference goldens."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmd_gen_b200 import _lib
from cmd_gen_b200.weights import init_weights, pack_blob
from tests.helpers import case_config, load, T

case = sys.argv[1] if len(sys.argv) > 1 else "ca_small"
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
g = load(f"dynamics_{case}.npz")
cfg = case_config(case)
h = _lib.Handle(cfg, "cuda:0", prec)
h.set_weights(pack_blob(cfg, init_weights(cfg, int(g["wseed"]))))
h.plan(g["counts"], g["sizes"])
B = len(g["sizes"])
for i, tv in enumerate(g["t_values"][:2]):
    t = torch.full((B,), float(tv))
    op, orr = h.dynamics_forward(T(g["z"]), T(g["xh_pocket"]), t)
    torch.cuda.synchronize()
    op, orr = op.cpu().numpy(), orr.cpu().numpy()
    rp, rr = g[f"eps_phar_f64_{i}"], g[f"eps_res_f64_{i}"]
    print(f"case={case} prec={prec} mask={os.environ.get('DIFFPHAR_TC_MASK','3')} t={tv}: "
          f"h err {np.abs(op[:,3:]-rp[:,3:]).max():.3e} (|ref| {np.abs(rp[:,3:]).max():.3f})  "
          f"x err {np.abs(op[:,:3]-rp[:,:3]).max():.3e} (|vel| {np.abs(rp[:,:3]).max():.3e})  "
          f"res h err {np.abs(orr[:,3:]-rr[:,3:]).max():.3e}  finite={np.isfinite(op).all()}", flush=True)
