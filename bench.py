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
# reads); config3 / config5 are the larger parity-test configurations, runnable here for roofline context.
WORKLOADS = {
    "config2": dict(n_samples=64, n_res=150, n_phar=8, T=500, residue_nf=20, n_layers=5, density=None,
                    label="config2: 1 CA pocket x 64 samples per GPU, N_r=150, N_p=8"),
    "config3": dict(n_samples=16, n_res=2000, n_phar=8, T=500, residue_nf=11, n_layers=5, density=0.05,
                    label="config3: 1 full-atom pocket x 16 samples per GPU, N_r=2000, N_p=8"),
    "config5": dict(n_samples=4, n_res=4000, n_phar=12, T=500, residue_nf=11, n_layers=9, density=0.05,
                    label="config5: 1 full-atom pocket x 4 samples per GPU, N_r=4000, N_p=12, 9 blocks"),
}
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
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    if args.workload != "config2":
        args.no_cpu_baseline = True          # the CPU port needs minutes per denoising step at these sizes
    return args


def workload(rank: int, timesteps: int):
    w = WORKLOAD
    cfg = DynamicsConfig(residue_nf=w["residue_nf"], n_layers=w["n_layers"])
    kw = {} if w["density"] is None else {"density": w["density"]}
    pocket = make_pocket_batch([w["n_res"]], w["residue_nf"], seed=1 + rank, replicate=w["n_samples"], **kw)
    counts = [w["n_phar"]] * w["n_samples"]
    noise = draw_noise(timesteps + 2, w["n_samples"] * w["n_phar"], 3 + cfg.phar_nf, seed=123 + rank)
    xh = torch.cat([pocket["x"], pocket["one_hot"].float() / 4.0], 1).contiguous()
    return cfg, pocket, counts, noise, xh


def algorithmic_bytes_msg(N, E):
    # h[col] per edge, h[row] once per row, h_out once per node, int32 col,
    # two fp32 edge scalars, rowptr
    return 4 * H * (E + 2 * N) + 12 * E + 4 * (N + 1)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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


def cpu_port_step_seconds(cfg, pocket, counts, noise, tab, n_timed, n_warm):
    """Times `n_timed` denoising steps of the oracle port on the host cores (same batch)."""
    from oracle import diffphar_oracle as orc
    W = init_weights(cfg, 0)
    B = len(counts)
    counts_t = torch.tensor(counts)
    mask_p = torch.repeat_interleave(torch.arange(B), counts_t)
    px = pocket["x"].clone()
    xh0 = torch.cat([px, pocket["one_hot"].float() / 4.0], 1)
    mu = torch.cat([orc._scatter_mean(px, pocket["mask"], B), torch.zeros(B, cfg.phar_nf)], 1)[mask_p]
    z, xh_pocket = orc.noise_and_center(mu, xh0, torch.ones(()), noise[0], mask_p, pocket["mask"], B)
    times, n_edges = [], 0
    with torch.no_grad():
        for k in range(n_warm + n_timed):
            t0 = time.perf_counter()
            z, xh_pocket, _, edges = orc.ddpm_step(W, cfg, tab.rows[k], z, xh_pocket, noise[k + 1], mask_p,
                                                   pocket["mask"], B)
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
    sub = 2                                   # denoising steps per bench step (bounded sample)
    per_step = []
    e = 0
    for k in range(args.warmup + args.steps):
        sec, e = cpu_port_step_seconds(cfg, pocket, counts, noise, tab, sub, 0)
        if k >= args.warmup:
            per_step.append(sec)
    sec = sum(per_step) / len(per_step)
    run_s = sec * (args.timesteps + 1)
    value = B / run_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": run_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 1 CA pocket x 64 samples, N_r=150, N_p=8, T=%d" % args.timesteps,
                   "extrapolated": f"{sub} denoising steps timed per bench step, x{args.timesteps + 1}"},
        "edges_per_s_per_step": e / sec,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sub} of {args.timesteps + 1} denoiser steps of the same batch per bench step, "
                                   "oracle/diffphar_oracle.py (reference tree absent on the GPU box)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
        allx, allc = gather_point_clouds(out, counts_dev)
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

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (edge-message kernel), timed live with CUDA events ---
    n_prof = min(20, args.timesteps)
    tab_p = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, n_prof)
    h.set_step_table(tab_p.rows, tab_p.final)
    noise_p = noise_dev[: n_prof + 2].contiguous()
    h.profile_enable(True)
    h.sample(xh_dev0.clone(), noise_p)
    names = ["edge_msg", "node_linear", "edge_coord", "graph_build", "ddpm", "other"]
    prof = {n: h.profile_read(i) for i, n in enumerate(names)}
    h.profile_enable(False)
    flp = h.flags()
    N = n_p + B * WORKLOAD["n_res"]
    E = int(flp.last_n_edges)
    msg_ms, msg_n = prof["edge_msg"]
    avg_ms = msg_ms / max(msg_n, 1)
    bytes_launch = algorithmic_bytes_msg(N, E)
    peak, peak_kind = measured_peak()
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9
    tot_prof = sum(v[0] for v in prof.values())
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "edge_msg_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(args.workload, {}).get(args.precision)
        except Exception:
            traffic = None

    runs = args.steps
    samples_total = B * world * runs
    value = samples_total / (total_ms * 1e-3)
    ms_per_step = total_ms / runs
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "f16fast": "f16", "f16fast32": "f16"}.get(args.precision, args.precision), "data": "synthetic",
        "config": {"workload": "%s, T=%d, hidden 256, %d blocks, cutoff 6A (crossdocked_ca_cond.yml), "
                               "random-init weights" % (WORKLOAD["label"], args.timesteps, WORKLOAD["n_layers"]),
                   "nodes": N, "edges_last_step": E, "precision": args.precision,
                   "l2": "flushed (256 MB write) between timed iterations",
                   "step": "one full reverse diffusion = %d denoiser evaluations" % (args.timesteps + 1)},
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
                     "bytes_per_launch": bytes_launch, "avg_launch_us": avg_ms * 1e3, "launches_timed": msg_n,
                     "share_of_step": msg_ms / tot_prof if tot_prof else None,
                     "kernel_ms_by_kind": {k: v[0] for k, v in prof.items()},
                     "note": ("working set (~30 MB) is L2-resident at this size: latency-bound, not HBM-bound"
                              if args.workload == "config2" else "steady state: >100 tiles per CTA")},
    }
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sec, e_cpu = cpu_port_step_seconds(cfg, pocket, counts, noise, tab, 5, 2)
        line["cpu_baseline"] = {"value": B / (sec * (args.timesteps + 1)), "unit": "samples/s",
                                "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "5 timed (2 warm-up) denoising steps of the same 64-sample batch, "
                                          "extrapolated x%d; oracle/diffphar_oracle.py" % (args.timesteps + 1),
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
