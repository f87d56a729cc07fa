"""GPU debug tool: per-role clock64 timeline of CTA 0 of the tcgen05 edge kernel (last message-kernel launch
of one denoiser evaluation at config-2 size).  DIFFPHAR_TRACE=1 python scripts/edge_trace.py [precision]"""
import ctypes as C
import os
import sys

os.environ["DIFFPHAR_TRACE"] = "2"
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
n = 6 * 64 * 16
buf = (C.c_longlong * n)()
h.lib.dp_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
rc = h.lib.dp_debug_trace(h.h, buf, n)
assert rc == 0, h.lib.dp_last_error()
tr = [[[buf[(r * 64 + i) * 16 + k] for k in range(16)] for i in range(64)] for r in range(6)]
# per-CTA global-timer stamps (ns) of the traced launch: MMA role, rows 16.., two words per CTA
ctas = [(c, tr[1][16 + (c >> 3)][2 * (c & 7)], tr[1][16 + (c >> 3)][2 * (c & 7) + 1]) for c in range(148)]
ctas = [(c, a, b) for c, a, b in ctas if a and b]
for r in range(16, 64):
    tr[1][r] = [0] * 16
if ctas:
    g0 = min(a for _, a, _ in ctas)
    print("per-CTA global timer: launch spans %.2f us (first entry -> last exit); entries within %.2f us; exits %.2f .. %.2f us" % (
        (max(b for _, _, b in ctas) - g0) / 1e3, (max(a for _, a, _ in ctas) - g0) / 1e3,
        (min(b for _, _, b in ctas) - g0) / 1e3, (max(b for _, _, b in ctas) - g0) / 1e3))
    print("  slowest CTAs (cta, entry us, exit us):", [(c, round((a - g0) / 1e3, 2), round((b - g0) / 1e3, 2)) for c, a, b in sorted(ctas, key=lambda t: -t[2])[:8]])
    import statistics
    for lo, hi in ((0, 37), (37, 55), (55, 83), (83, 148)):
        grp = [(b - g0) / 1e3 for c, _, b in ctas if lo <= c < hi]
        if grp:
            print("  CTAs %3d..%3d: exit median %.2f us (min %.2f, max %.2f)" % (lo, hi - 1, statistics.median(grp), min(grp), max(grp)))
    print("  fastest CTAs:", [(c, round((a - g0) / 1e3, 2), round((b - g0) / 1e3, 2)) for c, a, b in sorted(ctas, key=lambda t: t[2])[:4]])
t0 = min(v for r in tr for it in r for v in it if v > 0)
names = {0: ["start", "xempty", "edges done", "meta issued", "x issued", "block start", "arrive", "e0 row", "e0 Pa", "e0 Pb", "e0 done", "e1 row", "e1 Pa", "e1 Pb", "e1 done"],
         1: ["start", "full", "tempty", "issued"],
         2: ["start", "tfull", "ld", "silu", "red", "bar", "gate", "seg"],
         3: ["start", "tfull", "ld", "silu", "red", "bar", "gate", "seg"]}
names[4] = names[5] = names[3]
print("E =", h.flags().last_n_edges, " (cycles relative to the first mark, CTA 0)")
w = tr[1][63]
print(f"weights: issue {w[0] - t0}  landed {w[1] - t0}")
k = tr[1][62]
print(f"kernel entry {k[0] - t0}  exit {k[1] - t0}  (traced CTA)")
for it in range(10):
    for r, rn in ((0, "producer"), (1, "mma"), (2, "epilogue0"), (3, "epilogue1"), (4, "epilogue2"), (5, "epilogue3")):
        row = tr[r][it]
        if not any(row):
            continue
        print(f"it {it} {rn:9s} " + "  ".join(f"{nm}={row[k] - t0}" for k, nm in enumerate(names[r]) if row[k]))
