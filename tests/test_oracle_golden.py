"""CPU: pins the oracle (oracle/diffphar_oracle.py) and the host schedule code
against fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from cmd_gen_b200.schedule import gamma_table, step_table
from cmd_gen_b200.weights import init_weights
from oracle import diffphar_oracle as orc
from tests.helpers import case_config, load, T

CASES = ["ca_small", "fa_small", "nocut", "mean_agg"]


@pytest.mark.parametrize("sched,Tn,prec", [("polynomial_2", 500, 1e-5), ("polynomial_2", 100, 1e-5), ("cosine", 50, 1e-4)])
def test_gamma_and_step_table_bit_exact(sched, Tn, prec):
    g = load("schedule.npz")
    key = f"{sched}_{Tn}"
    gamma = gamma_table(sched, Tn, prec)
    assert np.array_equal(gamma.numpy(), g[f"gamma_{key}"])          # bit-exact
    for n_steps in (Tn, 10):
        tab = step_table(gamma, Tn, None if n_steps == Tn else n_steps)
        assert tab.n_steps == n_steps
        assert np.array_equal(tab.rows.numpy(), g[f"rows_{key}_{n_steps}"])
    assert np.array_equal(step_table(gamma, Tn).final.numpy(), g[f"final_{key}"])


def test_schedule_known_values():
    # SURVEY.md §8c probe values of the reference (s=499, 250, 0 at T=500)
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500)
    rows = tab.rows
    for k, (ia, c, s) in {0: (1.610165, 0.989116, 0.783761), 249: (1.002679, 0.008062, 0.072804),
                          499: (1.000004, 0.001886, 0.002108)}.items():
        assert abs(1.0 / rows[k, 1].item() - ia) < 2e-6
        assert abs(rows[k, 2].item() - c) < 2e-6
        assert abs(rows[k, 3].item() - s) < 2e-6
    prod = torch.prod(1.0 / rows[:, 1].double()).item()
    assert abs(prod - 315.97) < 0.05


@pytest.mark.parametrize("name", CASES)
def test_exact_edges_match_reference(name):
    g = load(f"dynamics_{name}.npz")
    cfg = case_config(name)
    x = torch.cat([T(g["z"])[:, :3], T(g["xh_pocket"])[:, :3]])
    m = torch.cat([T(g["mask_phar"]), T(g["mask_res"])])
    e = orc.exact_edges(m, x, cfg.edge_cutoff)
    # the fixtures are centred, so cdist's mm-mode agrees with the exact predicate here
    assert np.array_equal(e.numpy(), g["edges_ref"])
    if "edges_ref_nomm" in g:
        assert np.array_equal(e.numpy(), g["edges_ref_nomm"])
    # properties the reference edge list has
    row, col = e
    assert bool(((row[1:] > row[:-1]) | ((row[1:] == row[:-1]) & (col[1:] > col[:-1]))).all())
    assert int((row == col).sum()) == x.shape[0]
    rowptr, c2 = orc.edges_to_csr(e, x.shape[0])
    assert int(rowptr[-1]) == e.shape[1]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_dynamics_matches_reference(name, dt):
    g = load(f"dynamics_{name}.npz")
    cfg = case_config(name)
    tdt = torch.float32 if dt == "f32" else torch.float64
    W = init_weights(cfg, int(g["wseed"]), dtype=torch.float64)
    W = {k: v.to(torch.float32).to(tdt) for k, v in W.items()}      # fp32 values, widened
    B = len(g["sizes"])
    for i, tv in enumerate(g["t_values"]):
        t = torch.full((B, 1), float(tv), dtype=torch.float32)
        a, b, _ = orc.dynamics_forward(W, cfg, T(g["z"]), T(g["xh_pocket"]), t, T(g["mask_phar"]), T(g["mask_res"]))
        ra, rb = g[f"eps_phar_{dt}_{i}"], g[f"eps_res_{dt}_{i}"]
        tol = 2e-6 if dt == "f32" else 1e-12
        # x channels: absolute tolerance scaled by coordinate magnitude (vel = x_out - x cancels)
        xs = float(np.abs(g["z"][:, :3]).max())
        assert np.abs(a.numpy()[:, :3] - ra[:, :3]).max() <= tol * max(xs, 1.0) * 4
        assert np.abs(a.numpy()[:, 3:] - ra[:, 3:]).max() <= tol * max(1.0, np.abs(ra[:, 3:]).max()) * 4
        assert np.abs(b.numpy() - rb).max() <= tol * max(1.0, np.abs(rb).max()) * 4
    if dt == "f32":
        a, _, _ = orc.dynamics_forward(W, cfg, T(g["z"]), T(g["xh_pocket"]), torch.tensor([0.25]),
                                       T(g["mask_phar"]), T(g["mask_res"]))
        assert np.abs(a.numpy() - g["eps_phar_f32_scalar_t"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["ca_small", "fa_small", "mean_agg"])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_joint_mode_dynamics_matches_reference(name, dt):
    """update_pocket_coords=True (dynamics.py:104-107, 133-136): every node moves, pocket velocities are returned and the
    per-sample mean over all nodes is removed; fixtures from the unmodified reference (oracle/make_golden_joint.py)."""
    g = load(f"dynamics_joint_{name}.npz")
    cfg = case_config(name)
    tdt = torch.float32 if dt == "f32" else torch.float64
    W = {k: v.to(torch.float32).to(tdt) for k, v in init_weights(cfg, int(g["wseed"]), dtype=torch.float64).items()}
    B = len(g["sizes"])
    xs = max(1.0, float(np.abs(g["z"][:, :3]).max()), float(np.abs(g["xh_pocket"][:, :3]).max()))
    tol = 2e-6 if dt == "f32" else 1e-12
    for i, tv in enumerate(g["t_values"]):
        t = torch.full((B, 1), float(tv), dtype=torch.float32)
        a, b, _ = orc.dynamics_forward(W, cfg, T(g["z"]), T(g["xh_pocket"]), t, T(g["mask_phar"]), T(g["mask_res"]),
                                       update_pocket_coords=True)
        ra, rb = g[f"eps_phar_{dt}_{i}"], g[f"eps_res_{dt}_{i}"]
        assert np.abs(rb[:, :3]).max() > 0                              # the pocket really moves in this mode
        for o, r in ((a.numpy(), ra), (b.numpy(), rb)):
            assert np.abs(o[:, :3] - r[:, :3]).max() <= tol * xs * 4
            assert np.abs(o[:, 3:] - r[:, 3:]).max() <= tol * max(1.0, np.abs(r[:, 3:]).max()) * 4
        # the velocity is mean-free per sample over ALL its nodes
        vel = torch.cat([a[:, :3], b[:, :3]]).double()
        m = torch.cat([T(g["mask_phar"]), T(g["mask_res"])])
        assert float(orc._scatter_mean(vel, m, B).abs().max()) <= 1e-6 * xs


@pytest.mark.parametrize("fixture,name", [("sampler_ca_small_T500_n12.npz", "ca_small"),
                                          ("sampler_ca_small_T20.npz", "ca_small"),
                                          ("sampler_fa_small_T500_n6.npz", "fa_small")])
def test_sampler_matches_reference_f64(fixture, name):
    """Free-running trajectories amplify rounding by up to 1/alpha_T ~ 316x, so the
    tight comparison is float64 oracle vs float64 reference."""
    g = load(fixture)
    cfg = case_config(name)
    Tn = int(g["T"])
    ts = None if int(g["timesteps"]) < 0 else int(g["timesteps"])
    W = {k: v.to(torch.float32).double() for k, v in init_weights(cfg, int(g["wseed"])).items()}
    tab = step_table(gamma_table("polynomial_2", Tn, 1e-5), Tn, ts, dtype=torch.float64)
    trace = []
    xh_phar, xh_pocket, mp, mr = orc.sample_given_pocket(
        W, cfg, tab, T(g["pocket_x"]).double(), T(g["pocket_one_hot"]), T(g["pocket_mask"]),
        T(g["counts"]), T(g["noise"]), trace=trace)
    assert np.array_equal(mp.numpy(), g["mask_phar"])
    ref = g["trace_z_f64"]
    assert len(trace) == ref.shape[0]
    for k, tr in enumerate(trace):
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(tr["z"].numpy() - ref[k]).max() <= 1e-9 * scale, k
    scale = np.abs(g["xh_phar_f64"][:, :3]).max()
    assert np.abs(xh_phar.numpy()[:, :3] - g["xh_phar_f64"][:, :3]).max() <= 1e-6 * scale   # result buffer is fp32
    assert np.array_equal(xh_phar.numpy()[:, 3:], g["xh_phar_f64"][:, 3:])                  # one-hot types
    assert np.abs(xh_pocket.numpy() - g["xh_pocket_f64"]).max() <= 1e-6 * scale


def test_sampler_f32_close_to_reference_f32():
    g = load("sampler_ca_small_T500_n12.npz")
    cfg = case_config("ca_small")
    W = init_weights(cfg, int(g["wseed"]))
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, 12)
    xh_phar, _, _, _ = orc.sample_given_pocket(W, cfg, tab, T(g["pocket_x"]), T(g["pocket_one_hot"]),
                                               T(g["pocket_mask"]), T(g["counts"]), T(g["noise"]))
    scale = np.abs(g["xh_phar_f64"][:, :3]).max()
    ref_err = np.abs(g["xh_phar_f32"][:, :3] - g["xh_phar_f64"][:, :3]).max()
    our_err = np.abs(xh_phar.numpy()[:, :3] - g["xh_phar_f64"][:, :3]).max()
    assert our_err <= max(10 * ref_err, 1e-4 * scale)
    assert np.array_equal(xh_phar.numpy()[:, 3:], g["xh_phar_f32"][:, 3:])
