#!/usr/bin/env python
"""Benchmark of the DiffPhar pocket-conditioned sampler hot path (BASELINE.json metric:
pharmacophore samples/s of full 500-step reverse diffusion; edges/s per step).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one complete `sample_given_pocket`-equivalent run (T=500 denoising steps +
the final p(x|z0) draw = 501 denoiser evaluations) over one batch: the config-2 workload
(one synthetic C-alpha pocket of 150 residues, 64 samples, 8 pharmacophore points each;
N = 10 112 nodes, E ~ 68 k directed edges, random-init weights of crossdocked_ca_cond.yml).
With N ranks every rank runs its own batch (weak scaling, no collective in the loop) and
the sampled point clouds are all-gathered once per run.

Kernel times behind the rooflines are CUDA events around every launch, recorded as NODES of the captured step graph and
read after each of 20 replays (`roofline.timing`); the same spans around eager launches (`*_eager`) and the message
kernel repeated 8x inside one event pair (`roofline.back_to_back`, informational) stand beside them.

Beside the headline line's `value` / `e2e` / `roofline` / `cpu_baseline` the same JSON line carries
  roofline_node : the fused node kernel against the tensor roofline (FLOPs of the node MLP + the factored first layers)
  rooflines     : K1 graph build, K3 coordinate path, K4 DDPM update against the HBM roofline (all latency-bound here)
  api_e2e       : the same run through the reference-facing Python API (ConditionalDDPM.sample_given_pocket)
  also          : BASELINE configs 3 and 5 at full size (one run each), and config 4 — a 100-pocket list sharded over the
                  ranks with `cmd_gen_b200.sharding.sample_pockets` (strong scaling: compare the N = 1, 2, 4, 8 lines)
  gpu_eager_baseline : the oracle port run eagerly in fp32 on the same GPU (ATen kernels), informational
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cmd_gen_b200.config import DynamicsConfig                      # noqa: E402
from cmd_gen_b200.schedule import gamma_table, step_table          # noqa: E402
from cmd_gen_b200.synthetic import make_pocket_batch, draw_noise   # noqa: E402
from cmd_gen_b200.weights import init_weights, pack_blob           # noqa: E402

H = 256
METRIC = "pocket-conditioned phar samples/sec (500-step EGNN sampling)"
# config2 is the configuration BASELINE.json's metric is quoted on (the default and the only bench line the driver
# reads); config3 / config5 are the larger configurations (reported under `also`), config4 the sharded pocket list.
WORKLOADS = {
    "config2": dict(n_samples=64, n_res=150, n_phar=8, T=500, residue_nf=20, n_layers=5, density=None,
                    label="config2: 1 CA pocket x 64 samples per GPU, N_r=150, N_p=8"),
    "config3": dict(n_samples=16, n_res=2000, n_phar=8, T=500, residue_nf=11, n_layers=5, density=0.05,
                    label="config3: 1 full-atom pocket x 16 samples per GPU, N_r=2000, N_p=8"),
    "config5": dict(n_samples=4, n_res=4000, n_phar=12, T=500, residue_nf=11, n_layers=9, density=0.05,
                    label="config5: 1 full-atom pocket x 4 samples per GPU, N_r=4000, N_p=12, 9 blocks"),
}
CONFIG4 = dict(n_pockets=100, n_samples=64, res_range=(100, 400), phar_range=(4, 12), seed=44,
               label="config4: 100 CA pockets (N_r ~ U[100,400], N_p ~ U[4,12]) x 64 samples, sharded over the ranks")
WORKLOAD = dict(WORKLOADS["config2"])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DIFFPHAR_PRECISION", "f16fast"),
                    choices=["fp32", "tf32", "bf16", "f16", "f16fast", "f16fast32"])
    ap.add_argument("--timesteps", type=int, default=WORKLOAD["T"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the config 3 / 4 / 5, API and eager-GPU side measurements")
    ap.add_argument("--sweep", default="", help="comma-separated batch sizes of the config-3 sweep, e.g. 16,64,256,512")
    ap.add_argument("--n-samples", type=int, default=0, help="override the workload's samples per GPU")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    if args.n_samples:
        WORKLOAD["n_samples"] = args.n_samples
        WORKLOAD["label"] = WORKLOAD["label"].replace("x %d samples" % WORKLOADS[args.workload]["n_samples"], "x %d samples" % args.n_samples)
    if args.workload != "config2" or args.steps < 2 or args.timesteps != 500:
        args.no_also = True
    return args


def config_dict(args):
    """Identical in both arms (the driver compares it)."""
    return {"workload": "%s, T=%d, hidden 256, %d blocks, cutoff 6A (crossdocked_ca_cond.yml), random-init weights"
                        % (WORKLOAD["label"], args.timesteps, WORKLOAD["n_layers"]),
            "l2": "flushed (256 MB write) between timed iterations",
            "step": "one full reverse diffusion = %d denoiser evaluations" % (args.timesteps + 1)}


def workload(rank: int, timesteps: int, w=None):
    w = w or WORKLOAD
    cfg = DynamicsConfig(residue_nf=w["residue_nf"], n_layers=w["n_layers"])
    kw = {} if w["density"] is None else {"density": w["density"]}
    pocket = make_pocket_batch([w["n_res"]], w["residue_nf"], seed=1 + rank, replicate=w["n_samples"], **kw)
    counts = [w["n_phar"]] * w["n_samples"]
    noise = draw_noise(timesteps + 2, w["n_samples"] * w["n_phar"], 3 + cfg.phar_nf, seed=123 + rank)
    xh = torch.cat([pocket["x"], pocket["one_hot"].float() / 4.0], 1).contiguous()
    return cfg, pocket, counts, noise, xh


# ---- algorithmic work per launch (stated in DESIGN.md §4) -----------------------------------------
def algorithmic_bytes_msg(N, E):
    # h[col] per edge, h[row] once per row, h_out once per node, int32 col, two fp32 edge scalars, rowptr
    return 4 * H * (E + 2 * N) + 12 * E + 4 * (N + 1)


def algorithmic_bytes_coord(Np, Ep):
    return 4 * H * (Ep + Np) + 24 * Ep + 24 * Np


def algorithmic_bytes_graph(N, E):
    return 16 * N + 16 * E + 4 * (N + 1)          # read x + sample id, write col / erow / d0 / edst, rowptr


def algorithmic_bytes_ddpm(Np, Nr):
    return 176 * Np + 24 * Nr


def node_flops_per_call(N, n_layers):
    # launch v = 0: projection of the embedded h (2 blocks of 256 outputs); v = 1 .. L: node MLP (K = 512 then 256) +
    # projection blocks (4: coordinate + next edge MLP; the last launch 2)
    blocks = [2] + [4] * (n_layers - 1) + [2]
    return sum(2 * H * ((768 if v > 0 else 0) + H * b) for v, b in enumerate(blocks)) * N


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1654.5)), "measured (MEASURED_PEAKS.json, burst)"
        except Exception:
            pass
    return 6650.0, 1650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def oracle_step_seconds(cfg, pocket, counts, noise, tab, n_timed, n_warm, device="cpu"):
    """Times `n_timed` denoising steps of the oracle port (same batch) on the host cores, or — device='cuda:0' —
    eagerly on the GPU (ATen kernels, the reference's dense N x N get_edges)."""
    from oracle import diffphar_oracle as orc
    dev = torch.device(device)
    W = {k: v.to(dev) for k, v in init_weights(cfg, 0).items()}
    B = len(counts)
    counts_t = torch.tensor(counts)
    mask_p = torch.repeat_interleave(torch.arange(B), counts_t).to(dev)
    mask_r = pocket["mask"].to(dev)
    px = pocket["x"].clone().to(dev)
    xh0 = torch.cat([px, pocket["one_hot"].float().to(dev) / 4.0], 1)
    mu = torch.cat([orc._scatter_mean(px, mask_r, B), torch.zeros(B, cfg.phar_nf, device=dev)], 1)[mask_p]
    z, xh_pocket = orc.noise_and_center(mu, xh0, torch.ones((), device=dev), noise[0].to(dev), mask_p, mask_r, B)
    times, n_edges = [], 0
    with torch.no_grad():
        for k in range(n_warm + n_timed):
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            z, xh_pocket, _, edges = orc.ddpm_step(W, cfg, tab.rows[k], z, xh_pocket, noise[k + 1].to(dev), mask_p, mask_r, B)
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if k >= n_warm:
                times.append(dt)
                n_edges = int(edges.shape[1])
    return sum(times) / len(times), n_edges


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, pocket, counts, noise, xh = workload(0, args.timesteps)
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, args.timesteps)
    B = len(counts)
    sub = 2 if args.workload == "config2" else 1   # denoising steps per bench step (bounded sample)
    per_step = []
    e = 0
    for k in range(args.warmup + args.steps):
        sec, e = oracle_step_seconds(cfg, pocket, counts, noise, tab, sub, 0)
        if k >= args.warmup:
            per_step.append(sec)
    sec = sum(per_step) / len(per_step)
    run_s = sec * (args.timesteps + 1)
    value = B / run_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": run_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),
        "extrapolated": f"{sub} denoising steps timed per bench step, x{args.timesteps + 1}",
        "edges_per_s_per_step": e / sec,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sub} of {args.timesteps + 1} denoiser steps of the same batch per bench step, "
                                   "oracle/diffphar_oracle.py (reference tree absent on the GPU box)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_mirror(cfg, dev, precision, timesteps=500):
    """The reference-facing module stack (EGNNDynamics + ConditionalDDPM mirrors) with the bench weights."""
    from cmd_gen_b200.equivariant_diffusion.dynamics import EGNNDynamics
    from cmd_gen_b200.equivariant_diffusion.conditional_model import ConditionalDDPM
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        dyn = EGNNDynamics(cfg.phar_nf, cfg.residue_nf, 3, joint_nf=cfg.joint_nf, hidden_nf=256, device=dev,
                           n_layers=cfg.n_layers, attention=cfg.attention, tanh=cfg.tanh, norm_constant=cfg.norm_constant,
                           inv_sublayers=cfg.inv_sublayers, update_pocket_coords=False, edge_cutoff=cfg.edge_cutoff,
                           precision=precision)
        dyn.load_state_dict(init_weights(cfg, 0))
        ddpm = ConditionalDDPM(dyn, cfg.phar_nf, cfg.residue_nf, 3, [[1.0, 1.0], [1.0, 1.0]], timesteps=timesteps,
                               noise_schedule="polynomial_2", noise_precision=1e-5, loss_type="l2",
                               norm_values=(1.0, 4.0)).to(dev)
    return ddpm


def timed_runs(fn, n_warm, n_timed, dev, flush=None):
    """CUDA-event time of n_timed calls of fn (each bracketed by a synchronize), after n_warm untimed ones."""
    for _ in range(n_warm):
        fn()
    torch.cuda.synchronize(dev)
    total = 0.0
    for _ in range(n_timed):
        if flush is not None:
            flush.fill_(1.0)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(dev)
        total += e0.elapsed_time(e1)
    return total / n_timed


def profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=True):
    """Per-family kernel times of n_prof denoising steps, CUDA events around every launch.
    in_graph=True : the events are event-record NODES of the captured step graph (the production graph, replayed
                    n_prof times, read after every replay) — launch durations where they count, between the dependent
                    kernels of the timed loop; the stand-alone final evaluation is not instrumented (calls = n_prof);
    in_graph=False: eager launches, one event pair per launch (calls = n_prof + 1): every span also holds the idle front
                    end of a launch into an empty stream (~4-6 us), which no graph replay pays.
    Returns (prof, flags, calls)."""
    tab_p = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, n_prof)
    h.set_step_table(tab_p.rows, tab_p.final)
    noise_p = noise_dev[: n_prof + 2].contiguous()
    h.profile_enable(in_graph if isinstance(in_graph, int) and not isinstance(in_graph, bool) else (2 if in_graph else 1))
    h.sample(xh_dev0.clone(), noise_p)
    names = ["edge_msg", "node_linear", "edge_coord", "graph_build", "ddpm", "other"]
    prof = {n: h.profile_read(i) for i, n in enumerate(names)}
    h.profile_enable(False)
    return prof, h.flags(), (n_prof if in_graph else n_prof + 1)


def side_workload(name, dev, precision, flush, n_samples=None, timesteps=500, full_run=True):
    """One full-size run of BASELINE config 3 / 5 (+ its message-kernel roofline).  full_run=False times 20 denoising
    steps only and extrapolates (the batch sweep: a full run at B = 512 takes more than half a minute)."""
    from cmd_gen_b200 import _lib
    w = dict(WORKLOADS[name])
    if n_samples:
        w["n_samples"] = n_samples
    steps = timesteps if full_run else 20
    cfg, pocket, counts, noise, xh = workload(0, steps, w)
    B, n_p = len(counts), sum(counts)
    h = _lib.Handle(cfg, dev, precision)
    h.set_weights(pack_blob(cfg, init_weights(cfg, 0)))
    h.plan(counts, [w["n_res"]] * B)
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, steps)
    h.set_step_table(tab.rows, tab.final)
    xh_dev0, noise_dev = xh.to(dev), noise.to(dev)
    ms = timed_runs(lambda: h.sample(xh_dev0.clone(), noise_dev), 1, 1 if full_run else 2, dev, flush)
    fl = h.flags()
    assert fl.edge_overflow == 0
    n_prof = min(10, steps)
    prof_e, _, _ = profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=False)
    prof, flp, calls = profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=True)
    N, E = n_p + B * w["n_res"], int(flp.last_n_edges)
    msg_ms, msg_n = prof["edge_msg"]
    node_ms, node_n = prof["node_linear"]
    hbm, tens, src = measured_peaks()
    ach = algorithmic_bytes_msg(N, E) / (msg_ms / max(msg_n, 1) * 1e-3) / 1e9
    node_tf = node_flops_per_call(N, w["n_layers"]) * calls / (node_ms * 1e-3) / 1e12
    step_us = ms * 1e3 / (steps + 1)
    out = {"workload": w["label"].replace("x %d samples" % WORKLOADS[name]["n_samples"], "x %d samples" % B),
           "value": B / (step_us * 1e-6 * (timesteps + 1)), "unit": "samples/s", "denoise_step_us": step_us,
           "nodes": N, "edges_last_step": E, "precision": precision,
           "timed": "one full %d-step run" % timesteps if full_run else "20 denoising steps, extrapolated x%d" % (timesteps + 1),
           "roofline": {"bound": "hbm", "kernel": "edge message kernel", "achieved": ach, "peak": hbm, "unit": "GB/s",
                        "frac": ach / hbm, "avg_launch_us": msg_ms / max(msg_n, 1) * 1e3,
                        "avg_launch_us_eager": prof_e["edge_msg"][0] / max(prof_e["edge_msg"][1], 1) * 1e3,
                        "bytes_per_launch": algorithmic_bytes_msg(N, E)},
           "roofline_node": {"bound": "tensor", "kernel": "fused node kernel", "achieved": node_tf, "peak": tens,
                             "unit": "TFLOP/s", "frac": node_tf / tens, "avg_launch_us": node_ms / max(node_n, 1) * 1e3},
           "kernel_ms_by_kind": {k: v[0] for k, v in prof.items()}}
    del h
    torch.cuda.empty_cache()
    return out


def config4_pockets():
    g = torch.Generator().manual_seed(CONFIG4["seed"])
    lo, hi = CONFIG4["res_range"]
    sizes = torch.randint(lo, hi + 1, (CONFIG4["n_pockets"],), generator=g).tolist()
    plo, phi = CONFIG4["phar_range"]
    n_ph = torch.randint(plo, phi + 1, (CONFIG4["n_pockets"],), generator=g).tolist()
    batch = make_pocket_batch(sizes, 20, seed=CONFIG4["seed"])
    pockets, off = [], 0
    for n in sizes:
        pockets.append({"x": batch["x"][off: off + n].clone(), "one_hot": batch["one_hot"][off: off + n].clone()})
        off += n
    return pockets, n_ph


def run_config4(ddpm, dev, world, rank):
    """BASELINE config 4 through sharding.sample_pockets: strong scaling (the work list is fixed, the ranks share it)."""
    import torch.distributed as dist
    from cmd_gen_b200.sharding import sample_pockets
    pockets, n_ph = config4_pockets()
    timing = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = sample_pockets(ddpm, pockets, CONFIG4["n_samples"], n_ph, seed=7, timing=timing)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1), timing["loop_s"] * 1e3, -timing["loop_s"] * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, loop_max, loop_min = float(t[0]), float(t[1]), -float(t[2])
    assert all(o is not None and bool(torch.isfinite(o).all()) for o in out)
    n = CONFIG4["n_pockets"] * CONFIG4["n_samples"]
    return {"workload": CONFIG4["label"], "value": n / (total_ms * 1e-3), "unit": "samples/s", "scaling": "strong",
            "n_gpus": world, "run_ms": total_ms, "slowest_rank_loop_ms": loop_max, "fastest_rank_loop_ms": loop_min,
            "load_imbalance": loop_max / max(loop_min, 1e-9) - 1.0,
            "graph_captures_rank0": ddpm.dynamics.handle(dev).graph_captures(),
            "noise": "device counter-based generator keyed by (seed, pocket, sample): independent of the rank count",
            "gather": "one all_gather sized from the globally known layout"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from cmd_gen_b200 import _lib
    from cmd_gen_b200.sharding import gather_point_clouds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg, pocket, counts, noise, xh = workload(rank, args.timesteps)
    B, n_p = len(counts), sum(counts)
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, args.timesteps)
    h = _lib.Handle(cfg, dev, args.precision)
    h.set_weights(pack_blob(cfg, init_weights(cfg, 0)))
    h.plan(counts, [WORKLOAD["n_res"]] * B)
    h.set_step_table(tab.rows, tab.final)

    xh_dev0 = xh.to(dev)
    noise_dev = noise.to(dev)
    counts_dev = torch.tensor(counts, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_run():
        xh_dev = xh_dev0.clone()
        out = h.sample(xh_dev, noise_dev)
        # every rank holds the same layout: one fixed-size collective, no size exchange, no host sync
        allx, allc = gather_point_clouds(out, counts_dev, max_points=n_p, max_samples=B, uniform=True)
        return out, allx

    # ---- device-resident timing ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        one_run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.launch_count()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1.0)                       # L2 flush between timed iterations (outside the events)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, allx = one_run()
        e1.record()
        barrier()
        total_ms += e0.elapsed_time(e1)
    launches = h.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    fl = h.flags()
    timing_experiment = bool(os.environ.get("DIFFPHAR_SKIP"))       # kernels skipped on purpose: numbers only
    assert timing_experiment or (fl.edge_overflow == 0 and torch.isfinite(out).all())
    assert allx.shape[0] == world * n_p

    # ---- end to end through the C-ABI with HOST buffers -----------------------------------
    xh_pin, noise_pin = xh.pin_memory(), noise.pin_memory()
    out_pin = torch.empty((n_p, 3 + cfg.phar_nf)).pin_memory()
    pocket_pin = torch.empty_like(xh).pin_memory()
    h.sample_host(xh_pin, noise_pin, out_pin, pocket_pin)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.sample_host(xh_pin, noise_pin, out_pin, pocket_pin)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert timing_experiment or torch.equal(out_pin, out.cpu()), "host path and device path disagree"

    # ---- config 4 (sharded pocket list) runs on every rank ------------------------------------
    also = []
    ddpm = None
    if not args.no_also:
        ddpm = build_mirror(cfg, dev, args.precision)
        try:
            also.append(run_config4(ddpm, dev, world, rank))
        except Exception as ex:                                  # a side measurement must not take the headline down
            also.append({"workload": CONFIG4["label"], "error": repr(ex)[:300]})

    if rank != 0:
        return
    # ---- rooflines, timed live with CUDA events around every launch: inside the replayed step graph (primary) and around
    # eager launches (kept beside it: what round 1 reported) -------------------------------------------------------------
    n_prof = min(20, args.timesteps)
    prof_e, _, calls_e = profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=False)
    prof, flp, calls = profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=True)
    prof_b, _, _ = profile_families(h, xh_dev0, noise_dev, n_prof, in_graph=3)     # message spans hold 8 back-to-back launches
    h.set_step_table(tab.rows, tab.final)
    N = n_p + B * WORKLOAD["n_res"]
    E, Ep = int(flp.last_n_edges), int(flp.last_n_edges_phar)
    msg_ms, msg_n = prof["edge_msg"]
    avg_ms = msg_ms / max(msg_n, 1)
    bytes_launch = algorithmic_bytes_msg(N, E)
    peak, peak_tensor, peak_kind = measured_peaks()
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9
    tot_prof = sum(v[0] for v in prof.values())
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "edge_msg_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(args.workload, {}).get(args.precision)
        except Exception:
            traffic = None
    node_ms, node_n = prof["node_linear"]
    node_tf = node_flops_per_call(N, WORKLOAD["n_layers"]) * calls / (node_ms * 1e-3) / 1e12 if node_ms else 0.0
    step_us_now = total_ms / args.steps * 1e3 / (args.timesteps + 1)

    def hbm_entry(kernel, per_call_bytes, fam):
        ms, n = prof[fam]
        a = per_call_bytes * calls / (ms * 1e-3) / 1e9 if ms else 0.0
        return {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                "us_per_denoiser_call": ms * 1e3 / calls, "launches_per_call": n / calls,
                "bytes_per_denoiser_call": per_call_bytes}

    runs = args.steps
    samples_total = B * world * runs
    value = samples_total / (total_ms * 1e-3)
    ms_per_step = total_ms / runs
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "f16fast": "f16", "f16fast32": "f16"}.get(args.precision, args.precision), "data": "synthetic",
        "config": config_dict(args),
        "nodes": N, "edges_last_step": E, "precision": args.precision,
        "edges_per_s_per_step": E * world / (ms_per_step * 1e-3 / (args.timesteps + 1)),
        "denoise_step_us": ms_per_step * 1e3 / (args.timesteps + 1),
        "e2e": {"value": B * world * runs / e2e_s, "unit": "samples/s",
                "h2d_bytes_per_step": int(xh.numel() * 4 + noise.numel() * 4),
                "d2h_bytes_per_step": int(out_pin.numel() * 4 + pocket_pin.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "edge message kernel (GCL edge MLP + gate + segmented sum)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture "
                                       "summarised in profiles/ (profiles/edge_msg_traffic.json names the file)",
                     "bytes_per_launch": bytes_launch, "avg_launch_us": avg_ms * 1e3, "launches_timed": msg_n,
                     "timing": "CUDA events recorded as nodes of the captured step graph, around every launch of %d replays "
                               "(the loop the headline times); avg_launch_us_eager = the same kernel between events around an "
                               "eager launch, which also holds the idle front end of a launch into an empty stream" % calls,
                     "avg_launch_us_eager": prof_e["edge_msg"][0] / max(prof_e["edge_msg"][1], 1) * 1e3,
                     "frac_eager": bytes_launch / (prof_e["edge_msg"][0] / max(prof_e["edge_msg"][1], 1) * 1e-3) / 1e9 / peak,
                     "back_to_back": {"launches_per_event_pair": 8,
                                      "avg_launch_us": prof_b["edge_msg"][0] / max(prof_b["edge_msg"][1], 1) / 8 * 1e3,
                                      "frac": bytes_launch / (prof_b["edge_msg"][0] / max(prof_b["edge_msg"][1], 1) / 8 * 1e-3) / 1e9 / peak,
                                      "note": "informational: the same message launch repeated 8 times between ONE event pair (eager, "
                                              "same inputs, L2-resident as in the loop) amortises the event pair's own ~5 us; the sum "
                                              "of the single-launch spans of a step exceeds the step itself by about a third"},
                     "share_of_step": msg_ms * 1e3 / calls / step_us_now,
                     "kernel_ms_by_kind": {k: v[0] for k, v in prof.items()},
                     "kernel_ms_by_kind_eager": {k: v[0] for k, v in prof_e.items()},
                     "denoiser_calls_timed": calls,
                     "note": ("working set (~30 MB) is L2-resident at this size: latency-bound, not HBM-bound"
                              if args.workload == "config2" else "steady state: >100 tiles per CTA")},
        "roofline_node": {"bound": "tensor", "kernel": "fused node kernel (node MLP + factored first layers of the next edge / coordinate MLP)",
                          "achieved": node_tf, "peak": peak_tensor, "unit": "TFLOP/s", "frac": node_tf / peak_tensor,
                          "flops_per_denoiser_call": node_flops_per_call(N, WORKLOAD["n_layers"]),
                          "us_per_denoiser_call": node_ms * 1e3 / calls, "avg_launch_us": node_ms * 1e3 / max(node_n, 1),
                          "avg_launch_us_eager": prof_e["node_linear"][0] * 1e3 / max(prof_e["node_linear"][1], 1),
                          "share_of_step": node_ms * 1e3 / calls / step_us_now},
        "rooflines": [hbm_entry("K1 radius graph -> CSR", algorithmic_bytes_graph(N, E), "graph_build"),
                      hbm_entry("K3 coordinate update (phar rows)", WORKLOAD["n_layers"] * algorithmic_bytes_coord(n_p, Ep), "edge_coord"),
                      hbm_entry("K4 DDPM update", algorithmic_bytes_ddpm(n_p, B * WORKLOAD["n_res"]), "ddpm")],
    }
    if not args.no_also:
        # ---- the same run through the reference-facing Python API (host pocket in, host point cloud out) ----------
        try:
            host_pocket = {k: (v.pin_memory() if v.dtype.is_floating_point else v) for k, v in pocket.items()}
            cnt = torch.tensor(counts)
            api = {}
            for tag, seed in (("torch_randn_noise", None), ("device_generator_noise", 11)):
                ddpm.noise_seed = seed

                def api_run():
                    pk = {k: v.to(dev, non_blocking=True) for k, v in host_pocket.items()}
                    xp, _, _, _ = ddpm.sample_given_pocket(pk, cnt)
                    return xp.cpu()
                api_run()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(max(2, min(args.steps, 5))):
                    api_run()
                api[tag] = B * max(2, min(args.steps, 5)) / (time.perf_counter() - t0)
            ddpm.noise_seed = None
            line["api_e2e"] = {"unit": "samples/s", "call": "ConditionalDDPM.sample_given_pocket (host pocket -> .to(device) -> "
                               "sample -> .cpu()), wall clock", **api,
                               "graph_captures": ddpm.dynamics.handle(dev).graph_captures()}
        except Exception as ex:
            line["api_e2e"] = {"error": repr(ex)[:300]}
        for name in ("config3", "config5"):
            try:
                also.append(side_workload(name, dev, args.precision, flush))
            except Exception as ex:
                also.append({"workload": WORKLOADS[name]["label"], "error": repr(ex)[:300]})
        # ---- the other arithmetic modes on the same batch (two timed runs each; f16fast is the headline) --------------
        try:
            modes = {}
            for prec in ("bf16", "f16", "tf32"):
                if prec == args.precision:
                    continue
                hp = _lib.Handle(cfg, dev, prec)
                hp.set_weights(pack_blob(cfg, init_weights(cfg, 0)))
                hp.plan(counts, [WORKLOAD["n_res"]] * B)
                hp.set_step_table(tab.rows, tab.final)
                ms = timed_runs(lambda: hp.sample(xh_dev0.clone(), noise_dev), 1, 2, dev, flush)
                modes[prec] = B / (ms * 1e-3)
                del hp
            line["precision_lines"] = {"unit": "samples/s per GPU", "workload": "the headline batch, device-resident, 2 timed runs per mode (rank 0)",
                                       args.precision: value / world, **modes,
                                       "note": "bf16 / f16 / tf32 keep north_star's letter (16-bit or TF32 operands only in the MLP "
                                               "contractions); f16fast also runs the edge kernels' first layer in packed f16x2"}
        except Exception as ex:
            line["precision_lines"] = {"error": repr(ex)[:300]}
        for bs in [int(v) for v in args.sweep.split(",") if v]:
            try:
                also.append(side_workload("config3", dev, args.precision, flush, n_samples=bs, full_run=False))
            except Exception as ex:
                also.append({"workload": "config3 sweep B=%d" % bs, "error": repr(ex)[:300]})
        # ---- the oracle port run eagerly on the same GPU: what "the reference on a B200" costs (informational) ----
        try:
            sec, _ = oracle_step_seconds(cfg, pocket, counts, noise, tab, 5, 2, device=str(dev))
            line["gpu_eager_baseline"] = {"value": B / (sec * (args.timesteps + 1)), "unit": "samples/s", "kind": "port",
                                          "sample": "5 timed (2 warm-up) denoising steps of the same batch, fp32 ATen kernels on "
                                                    "the same GPU, dense N x N get_edges, extrapolated x%d" % (args.timesteps + 1),
                                          "s_per_denoise_step": sec}
        except Exception as ex:
            line["gpu_eager_baseline"] = {"error": repr(ex)[:300]}
    if also:
        line["also"] = also
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        n_t, n_w = (5, 2) if args.workload == "config2" else (1, 0)
        sec, e_cpu = oracle_step_seconds(cfg, pocket, counts, noise, tab, n_t, n_w)
        line["cpu_baseline"] = {"value": B / (sec * (args.timesteps + 1)), "unit": "samples/s",
                                "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d timed (%d warm-up) denoising steps of the same %d-sample batch, "
                                          "extrapolated x%d; oracle/diffphar_oracle.py" % (n_t, n_w, B, args.timesteps + 1),
                                "s_per_denoise_step": sec}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
