"""GPU debug tool: clock64 timeline of CTA 0 of the fused tcgen05 node kernel (last launch of one denoiser
evaluation at config-2 size).  python scripts/node_trace.py [precision] [cta] [h version]
(cta: which CTA writes the timeline, default 0 = a tile with phar rows; h version: which of the 6 node launches, default 3)"""
import ctypes as C
import os
import sys

os.environ["DIFFPHAR_TRACE"] = "1"
os.environ["DIFFPHAR_TRACE_CTA"] = sys.argv[2] if len(sys.argv) > 2 else "0"
os.environ["DIFFPHAR_TRACE_V"] = sys.argv[3] if len(sys.argv) > 3 else "3"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmd_gen_b200 import _lib
from cmd_gen_b200.config import DynamicsConfig
from cmd_gen_b200.synthetic import make_pocket_batch
from cmd_gen_b200.weights import init_weights, pack_blob

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
cfg = DynamicsConfig()
B, n_res, n_ph = 64, 150, 8
h = _lib.Handle(cfg, "cuda:0", prec)
h.set_weights(pack_blob(cfg, init_weights(cfg, 0)))
pocket = make_pocket_batch([n_res], 20, seed=3, replicate=B)
gen = torch.Generator().manual_seed(4)
com = pocket["x"][:n_res].mean(0)
z = torch.cat([com + 5.0 * torch.randn(B * n_ph, 3, generator=gen), torch.randn(B * n_ph, 8, generator=gen)], 1)
xr = torch.cat([pocket["x"], pocket["one_hot"].float() / 4], 1)
h.plan([n_ph] * B, [n_res] * B)
t = torch.full((B,), 0.4)
for _ in range(3):
    h.dynamics_forward(z, xr, t)
torch.cuda.synchronize()
n = 3 * 64 * 16
buf = (C.c_longlong * n)()
h.lib.dp_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
assert h.lib.dp_debug_trace(h.h, buf, n) == 0, h.lib.dp_last_error()
tr = [[[buf[(r * 64 + i) * 16 + k] for k in range(16)] for i in range(64)] for r in range(3)]
t0 = min(v for r in tr for it in r for v in it if v > 0)
cn = ["start", "h staged", "agg staged", "acc0 full", "t stored", "acc1 full", "h stored"] + \
     [f"{w}{b}" for b in range(4) for w in ("P full ", "P stored ")]
print("compute warp 0: " + "  ".join(f"{nm}={tr[0][0][k] - t0}" for k, nm in enumerate(cn) if k < 16 and tr[0][0][k]))
for p in range(32):
    m, w = tr[1][p], tr[2][p]
    if not any(m) and not any(w):
        continue
    print(f"panel {p:2d}  tma: wait {w[0] - t0} issue {w[1] - t0}   mma: wait {m[0] - t0} full {m[1] - t0} issued {m[2] - t0}")
