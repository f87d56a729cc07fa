"""TEST INFRASTRUCTURE — BASELINE-config-sized fixtures from the UNMODIFIED reference.

    python -m oracle.make_golden_large [config1 config2 config3 config5]

Runs the reference's own ``EGNNDynamics`` / ``ConditionalDDPM`` (oracle/ref_shims.py, fp32 and
``.double()``) in the build container on the seeded inputs of ``oracle/large_cases.py`` and writes
``tests/golden/large_*.npz``.  The fixtures hold reference OUTPUTS and checksums of the inputs; the
inputs themselves are regenerated from their seeds by the tests (they would be tens of MB).

  large_sampler_config1.npz : ConditionalDDPM.sample_given_pocket, ALL 500 steps, B = 10 (BASELINE configs[0]):
                              final point cloud, types, translated pocket, z every 50 steps, fp32 and fp64
  large_dynamics_config2/3/5: one EGNNDynamics.forward per t value at the full batch size: eps_hat of every
                              phar node (fp64 and fp32), of every pocket node for the first t (fp64 rounded
                              to fp32), the edge list's degrees + sha256, the count of pairs where
                              torch.cdist's mm-mode disagrees with the exact predicate
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmd_gen_b200.weights import init_weights           # noqa: E402
from oracle import large_cases as lc                    # noqa: E402
from oracle import ref_shims                            # noqa: E402
from oracle import diffphar_oracle as orc               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_dynamics(name):
    t0 = time.time()
    d = lc.dynamics_inputs(name)
    cfg = d["cfg"]
    state = init_weights(cfg, seed=d["wseed"])
    ref32 = ref_shims.build_reference_model(cfg, state, T=500)
    ref64 = ref_shims.build_reference_model(cfg, state, T=500, dtype=torch.float64)
    z, xp, mp, mr = d["z"], d["xh_pocket"], d["mask_phar"], d["mask_res"]
    B = len(d["sizes"])
    out = dict(sizes=np.array(d["sizes"]), counts=np.array(d["counts"]), wseed=d["wseed"],
               t_values=np.array(d["t_values"], dtype=np.float32), z=z.numpy(),
               z_checksum=lc.checksum(z), pocket_checksum=lc.checksum(xp))
    x_all = torch.cat([z[:, :3], xp[:, :3]], 0)
    m_all = torch.cat([mp, mr])
    with torch.no_grad():
        e_ref = ref32.dynamics.get_edges(m_all, x_all).numpy()      # the reference's own list (cdist, mm-mode above 25 rows)
        e_exact = orc.exact_edges(m_all, x_all, cfg.edge_cutoff).numpy()
        out["n_edges"] = e_ref.shape[1]
        out["n_edges_phar"] = int((e_ref[0] < z.shape[0]).sum())
        out["edges_ref_sha256"] = lc.edges_digest(e_ref)
        out["edges_exact_sha256"] = lc.edges_digest(e_exact)
        out["degrees_ref"] = np.bincount(e_ref[0], minlength=x_all.shape[0]).astype(np.int32)
        sa = set(map(tuple, e_ref.T.tolist())) if e_ref.shape != e_exact.shape or not np.array_equal(e_ref, e_exact) else None
        out["mm_mode_disagreements"] = 0 if sa is None else len(sa ^ set(map(tuple, e_exact.T.tolist())))
        print(name, "N", x_all.shape[0], "E", e_ref.shape[1], "E_p", out["n_edges_phar"],
              "mm-mode disagreements", out["mm_mode_disagreements"], flush=True)
        for i, tv in enumerate(d["t_values"]):
            t = torch.full((B, 1), tv, dtype=torch.float32)
            a32, b32 = ref32.dynamics(z, xp, t, mp, mr)
            a64, b64 = ref64.dynamics(z.double(), xp.double(), t, mp, mr)
            out[f"eps_phar_f32_{i}"] = a32.numpy()
            out[f"eps_phar_f64_{i}"] = a64.numpy()
            out[f"ref_err_x_{i}"] = float((a32[:, :3].double() - a64[:, :3]).abs().max())
            out[f"ref_err_h_{i}"] = float(max((a32[:, 3:].double() - a64[:, 3:]).abs().max(),
                                               (b32[:, 3:].double() - b64[:, 3:]).abs().max()))
            out[f"eps_res_absmax_{i}"] = float(b64[:, 3:].abs().max())
            if i == 0:
                out["eps_res_f64as32_0"] = b64[:, 3:].to(torch.float32).numpy()
            assert bool((b64[:, :3] == 0).all())
            print(name, f"t={tv}: |eps_h| max {float(a64[:, 3:].abs().max()):.4f}, |eps_x| max {float(a64[:, :3].abs().max()):.3e}, "
                  f"reference fp32-vs-fp64 err x {out[f'ref_err_x_{i}']:.2e} h {out[f'ref_err_h_{i}']:.2e}  [{time.time() - t0:.0f} s]", flush=True)
    np.savez_compressed(os.path.join(OUT, f"large_dynamics_{name}.npz"), **out)


def gen_sampler(name):
    t0 = time.time()
    d = lc.sampler_inputs(name)
    cfg = d["cfg"]
    state = init_weights(cfg, seed=d["wseed"])
    counts_t = torch.tensor(d["counts"], dtype=torch.int64)
    noise = d["noise"]
    pocket0 = d["pocket"]
    out = dict(counts=np.array(d["counts"]), sizes=pocket0["size"].numpy(), T=d["T"], wseed=d["wseed"],
               trace_every=d["trace_every"], noise_checksum=lc.checksum(noise), pocket_checksum=lc.checksum(pocket0["x"]))
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        ddpm = ref_shims.build_reference_model(cfg, state, T=d["T"], dtype=dt)
        pocket = {k: v.clone() for k, v in pocket0.items()}
        if dt == torch.float64:
            pocket["x"] = pocket["x"].double()
        trace = []
        inner = ddpm.sample_p_zs_given_zt
        calls = [0]

        def rec(*a, **k):
            zz, pp = inner(*a, **k)
            calls[0] += 1
            if calls[0] % d["trace_every"] == 0:
                trace.append(zz.clone())
            return zz, pp
        ddpm.sample_p_zs_given_zt = rec
        with torch.no_grad(), ref_shims.InjectedNoise(ddpm, noise):
            xh_phar, xh_pock, mp, mr = ddpm.sample_given_pocket(pocket, counts_t)
        assert calls[0] == d["T"]
        out[f"xh_phar_{tag}"] = xh_phar.numpy()
        out[f"pocket_x_{tag}"] = xh_pock[:, :3].numpy()
        out[f"trace_z_{tag}"] = torch.stack(trace).to(torch.float64).numpy()
        out["mask_phar"] = mp.numpy()
        print(name, tag, "done: |x| max", float(xh_phar[:, :3].abs().max()), f"[{time.time() - t0:.0f} s]", flush=True)
    err = np.abs(out["xh_phar_f32"][:, :3] - out["xh_phar_f64"][:, :3]).max()
    same = (out["xh_phar_f32"][:, 3:] == out["xh_phar_f64"][:, 3:]).all(1).mean()
    print(name, "reference fp32 vs fp64 after 500 steps: max |dx|", err, "type agreement", same, flush=True)
    np.savez_compressed(os.path.join(OUT, f"large_sampler_{name}.npz"), **out)


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    names = sys.argv[1:] or ["config2", "config1", "config5", "config3"]
    for n in names:
        if n in lc.SAMPLER_CASES:
            gen_sampler(n)
        else:
            gen_dynamics(n)


if __name__ == "__main__":
    main()
