"""Small sampler workload for compute-sanitizer (scripts/gpu_sanitize.sh): every kernel family, every precision,
both segmented-sum schemes, both radius-graph builders, the in-graph frame output and the device noise generator."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmd_gen_b200 import _lib                                        # noqa: E402
from cmd_gen_b200.config import DynamicsConfig                       # noqa: E402
from cmd_gen_b200.schedule import gamma_table, step_table           # noqa: E402
from cmd_gen_b200.synthetic import make_pocket_batch                 # noqa: E402
from cmd_gen_b200.weights import init_weights, pack_blob             # noqa: E402


def main():
    precisions = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "f16", "f16fast", "bf16"]
    cfg = DynamicsConfig(n_layers=2)
    blob = pack_blob(cfg, init_weights(cfg, 0))
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, 4)
    for prec in precisions:
        for seg, graph, sizes, counts, density in (("units", "scan", [40, 31], [5, 4], 0.0074),
                                                   ("units", "fused", [40, 31], [5, 4], 0.0074),
                                                   ("lanes", "cells", [600, 530], [6, 3], 0.05)):
            os.environ["DIFFPHAR_SEG"], os.environ["DIFFPHAR_GRAPH"] = seg, graph
            h = _lib.Handle(cfg, "cuda:0", prec)
            h.set_weights(blob)
            h.plan(counts, sizes)
            h.set_step_table(tab.rows, tab.final)
            pocket = make_pocket_batch(sizes, cfg.residue_nf, density=density, seed=2)
            xh = torch.cat([pocket["x"], pocket["one_hot"].float() / 4.0], 1).cuda().contiguous()
            out, fp, fk = h.sample(xh.clone(), noise=None, seed=5, return_frames=2, norm=(1.0, 4.0, 0.0))
            torch.cuda.synchronize()
            fl = h.flags()
            # joint mode (update_pocket_coords=True): the coordinate kernel covers every edge and finishes every row
            h.set_update_pocket_coords(True)
            n_p = sum(counts)
            zj = torch.cat([xh[:n_p, :3] + 1.0, torch.zeros(n_p, cfg.phar_nf, device="cuda")], 1).contiguous()
            jp, jr = h.dynamics_forward(zj, xh, torch.full((len(counts),), 0.3))
            torch.cuda.synchronize()
            assert bool(torch.isfinite(jp).all()) and bool(torch.isfinite(jr).all())
            print(prec, seg, graph, "E", fl.last_n_edges, "finite", bool(torch.isfinite(out).all()), "launches", h.launch_count(), flush=True)
            del h


if __name__ == "__main__":
    main()
