"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (oracle/ref_shims.py) in the build container.  Run once:

    python -m oracle.make_golden

The fixtures hold inputs + reference outputs only; weights are regenerated from
``cmd_gen_b200.weights.init_weights(cfg, seed)`` on both sides.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmd_gen_b200.config import DynamicsConfig          # noqa: E402
from cmd_gen_b200.weights import init_weights           # noqa: E402
from cmd_gen_b200.schedule import gamma_table, step_table  # noqa: E402
from cmd_gen_b200.synthetic import make_pocket_batch, draw_noise, CA_DENSITY, FULL_ATOM_DENSITY  # noqa: E402
from oracle import ref_shims                            # noqa: E402
from oracle import diffphar_oracle as orc               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (cfg kwargs, pocket sizes, phar counts, density, weight seed)
    "ca_small": (dict(), [20, 35, 28], [4, 6, 5], CA_DENSITY, 0),
    "fa_small": (dict(residue_nf=11, n_layers=3, inv_sublayers=2), [60, 45], [3, 7], FULL_ATOM_DENSITY, 1),
    "nocut": (dict(n_layers=2, edge_cutoff=None, attention=False, tanh=False, norm_constant=0.0),
              [10, 12], [2, 3], CA_DENSITY, 2),
    "mean_agg": (dict(n_layers=2, aggregation_method="mean", condition_time=False), [25], [12], CA_DENSITY, 3),
}


def initial_state(cfg, pocket, counts, seed):
    """z_T-like COM-free phar state at the pocket centre + translated pocket."""
    B = len(counts)
    counts_t = torch.tensor(counts, dtype=torch.int64)
    mask_p = torch.repeat_interleave(torch.arange(B), counts_t)
    px = pocket["x"].clone()
    ph = pocket["one_hot"].float() / 4.0
    xh0 = torch.cat([px, ph], 1)
    mu_x = orc._scatter_mean(px, pocket["mask"], B)
    mu = torch.cat([mu_x, torch.zeros(B, cfg.phar_nf)], 1)[mask_p]
    eps = draw_noise(1, int(counts_t.sum()), cfg.n_dims + cfg.phar_nf, seed=seed)[0]
    # spread the points over the pocket so that phar-residue edges exist
    eps[:, :3] *= 4.0
    z, xh_pocket = orc.noise_and_center(mu, xh0, torch.ones(()), eps, mask_p, pocket["mask"], B)
    return z, xh_pocket, mask_p, counts_t


def gen_dynamics(name):
    kw, sizes, counts, density, wseed = CASES[name]
    cfg = DynamicsConfig(**kw)
    state = init_weights(cfg, seed=wseed)
    pocket = make_pocket_batch(sizes, cfg.residue_nf, density=density, seed=11)
    z, xh_pocket, mask_p, counts_t = initial_state(cfg, pocket, counts, seed=7)
    mask_r = pocket["mask"]
    ref32 = ref_shims.build_reference_model(cfg, state, T=500)
    ref64 = ref_shims.build_reference_model(cfg, state, T=500, dtype=torch.float64)
    B = len(sizes)
    out = dict(z=z.numpy(), xh_pocket=xh_pocket.numpy(), mask_phar=mask_p.numpy(),
               mask_res=mask_r.numpy(), sizes=np.array(sizes), counts=np.array(counts), wseed=wseed)
    x_all = torch.cat([z[:, :3], xh_pocket[:, :3]], 0)
    m_all = torch.cat([mask_p, mask_r])
    with torch.no_grad():
        e_ref = ref32.dynamics.get_edges(m_all, x_all)              # reference's own (cdist mm-mode)
        out["edges_ref"] = e_ref.numpy()
        if cfg.edge_cutoff is not None:
            adj = (m_all[:, None] == m_all[None, :]) & (
                torch.cdist(x_all, x_all, compute_mode="donot_use_mm_for_euclid_dist") <= cfg.edge_cutoff)
            out["edges_ref_nomm"] = torch.stack(torch.where(adj), 0).numpy()
        ts = [1.0, 0.5, 0.002, 0.0]
        out["t_values"] = np.array(ts, dtype=np.float32)
        for i, tv in enumerate(ts):
            t = torch.full((B, 1), tv, dtype=torch.float32)
            a, b = ref32.dynamics(z, xh_pocket, t, mask_p, mask_r)
            out[f"eps_phar_f32_{i}"], out[f"eps_res_f32_{i}"] = a.numpy(), b.numpy()
            a, b = ref64.dynamics(z.double(), xh_pocket.double(), t, mask_p, mask_r)
            out[f"eps_phar_f64_{i}"], out[f"eps_res_f64_{i}"] = a.numpy(), b.numpy()
        # scalar-t branch (dynamics.py:93-95)
        a, _ = ref32.dynamics(z, xh_pocket, torch.tensor([0.25]), mask_p, mask_r)
        out["eps_phar_f32_scalar_t"] = a.numpy()
    np.savez_compressed(os.path.join(OUT, f"dynamics_{name}.npz"), **out)
    print(name, "edges", out["edges_ref"].shape[1])


def gen_sampler(name, T, timesteps, tag):
    kw, sizes, counts, density, wseed = CASES[name]
    cfg = DynamicsConfig(**kw)
    state = init_weights(cfg, seed=wseed)
    pocket0 = make_pocket_batch(sizes, cfg.residue_nf, density=density, seed=21)
    counts_t = torch.tensor(counts, dtype=torch.int64)
    n_steps = T if timesteps is None else timesteps
    noise = draw_noise(n_steps + 2, int(counts_t.sum()), cfg.n_dims + cfg.phar_nf, seed=123)
    out = dict(pocket_x=pocket0["x"].numpy(), pocket_one_hot=pocket0["one_hot"].numpy(),
               pocket_mask=pocket0["mask"].numpy(), pocket_size=pocket0["size"].numpy(),
               counts=np.array(counts), noise=noise.numpy(), T=T,
               timesteps=-1 if timesteps is None else timesteps, wseed=wseed)
    for dt, tag_dt in ((torch.float32, "f32"), (torch.float64, "f64")):
        ddpm = ref_shims.build_reference_model(cfg, state, T=T, dtype=dt)
        pocket = {k: v.clone() for k, v in pocket0.items()}
        if dt == torch.float64:
            pocket["x"] = pocket["x"].double()
        trace = []
        inner = ddpm.sample_p_zs_given_zt

        def rec(*a, **k):
            z, p = inner(*a, **k)
            trace.append(z.clone())
            return z, p
        ddpm.sample_p_zs_given_zt = rec
        with torch.no_grad(), ref_shims.InjectedNoise(ddpm, noise):
            xh_phar, xh_pock, mp, mr = ddpm.sample_given_pocket(pocket, counts_t, timesteps=timesteps)
        out[f"xh_phar_{tag_dt}"] = xh_phar.numpy()
        out[f"xh_pocket_{tag_dt}"] = xh_pock.numpy()
        out[f"trace_z_{tag_dt}"] = torch.stack(trace).to(torch.float64).numpy()
        out["mask_phar"] = mp.numpy()
    np.savez_compressed(os.path.join(OUT, f"sampler_{name}_{tag}.npz"), **out)
    print("sampler", name, tag, "final |x| max", np.abs(out["xh_phar_f32"][:, :3]).max())


def gen_schedule():
    """gamma table + per-step constants straight from the reference's methods."""
    out = {}
    for sched, T, prec in (("polynomial_2", 500, 1e-5), ("polynomial_2", 100, 1e-5), ("cosine", 50, 1e-4)):
        cfg = DynamicsConfig(n_layers=1)
        ddpm = ref_shims.build_reference_model(cfg, init_weights(cfg, 0), T=T, noise_schedule=sched, precision=prec,
                                               norm_values=(1.0, 1.0) if sched == "cosine" else (1.0, 4.0))
        key = f"{sched}_{T}"
        out[f"gamma_{key}"] = ddpm.gamma.gamma.detach().numpy()
        for n_steps in (T, 10):
            rows = []
            for B in (1, 16):
                zt = torch.zeros(B, 11)
                rb = []
                for s in reversed(range(n_steps)):
                    s_arr = torch.full((B, 1), fill_value=s)
                    t_arr = s_arr + 1
                    s_arr = s_arr / n_steps
                    t_arr = t_arr / n_steps
                    g_s, g_t = ddpm.gamma(s_arr), ddpm.gamma(t_arr)
                    s2, s_ts, a_ts = ddpm.sigma_and_alpha_t_given_s(g_t, g_s, zt)
                    sig_s, sig_t = ddpm.sigma(g_s, zt), ddpm.sigma(g_t, zt)
                    c = s2 / a_ts / sig_t
                    sig = s_ts * sig_s / sig_t
                    rb.append(torch.cat([t_arr, a_ts, c, sig], 1))     # [B,4]
                rows.append(torch.stack(rb))                            # [n,B,4]
            assert all(bool((r == r[:, :1]).all()) for r in rows), "constants differ across the batch"
            assert torch.equal(rows[0][:, 0], rows[1][:, 0]), "B=1 vs B=16 constants differ"
            out[f"rows_{key}_{n_steps}"] = rows[1][:, 0].numpy()
        t0 = torch.zeros(4, 1)
        g0 = ddpm.gamma(t0)
        sig_x = ddpm.SNR(-0.5 * g0)
        net = torch.zeros(4, 11)
        sig0 = ddpm.sigma(g0, net)
        a0 = ddpm.alpha(g0, net)
        out[f"final_{key}"] = torch.cat([t0, 1.0 / a0, sig0, sig_x], 1)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "schedule.npz"), **out)
    print("schedule ok")


def gen_state_keys():
    """The weight ABI: state-dict keys + shapes of the reference modules."""
    import json
    out = {}
    for name in CASES:
        cfg = DynamicsConfig(**CASES[name][0])
        ddpm = ref_shims.build_reference_model(cfg, init_weights(cfg, CASES[name][4]), T=500)
        out[name] = {"dynamics": {k: list(v.shape) for k, v in ddpm.dynamics.state_dict().items()},
                     "ddpm_extra": {k: list(v.shape) for k, v in ddpm.state_dict().items()
                                    if not k.startswith("dynamics.")}}
    with open(os.path.join(OUT, "state_keys.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("state keys ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_state_keys()
    torch.set_num_threads(8)
    gen_schedule()
    for name in CASES:
        gen_dynamics(name)
    gen_sampler("ca_small", 500, 12, "T500_n12")
    gen_sampler("ca_small", 20, None, "T20")
    gen_sampler("fa_small", 500, 6, "T500_n6")


if __name__ == "__main__":
    main()
