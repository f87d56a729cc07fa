"""Parity at BASELINE's own configurations against fixtures minted from the UNMODIFIED reference
(oracle/make_golden_large.py -> tests/golden/large_*.npz; inputs regenerated from seeds by oracle/large_cases.py).

  config1 : sample_given_pocket, ALL 500 steps, B = 10 — final cloud, types, pocket and z every 50 steps
  config2 : one denoiser call at bench size (B = 64, N = 10 112, E = 67 584) at t in {1, 0.5, 0.002, 0}
  config3 : 8 ragged full-atom pockets (N = 16 008, E = 625 524): cell-list builder, lane-range segmented sum,
            more than one wave of node tiles
  config5 : 4 000-node pockets, 9 blocks (N = 7 924, E = 321 232)

CPU half (not gpu): the oracle restatement against the same fixtures.  GPU half: the CUDA path in ALL FOUR
precisions against the reference's fp64 outputs — not against the library's own fp32 mode.

Stated tolerances (per denoiser call; `hs` = max |reference feature output|, `xs` = max |input coordinate|,
`vs` = max |reference velocity|; measured on a B200 in profiles/r05a_parity_errors.txt):
  fp32    : features 2e-6 hs, velocity 1e-6 xs           (measured 1.0e-7 / 1.8e-6 = the reference's own fp32-vs-fp64 error)
  tf32    : features 5e-5 hs, velocity 1e-6 xs + 0.02 vs  (tcgen05 kind::tf32 tiles, fp32 storage; measured 0.6 - 1.1e-5)
  f16     : features 5e-5 hs, velocity 1e-6 xs + 0.02 vs  (f16 operands, the same 10-bit mantissa; measured 0.5 - 1.0e-5)
  f16fast : features 5e-5 hs, velocity 1e-6 xs + 0.03 vs  (measured 0.6 - 1.2e-5)
  bf16    : features 5e-4 hs, velocity 1e-6 xs + 0.05 vs  (measured 0.5 - 1.0e-4)
  the same bounds at every configuration (config 3 / 5: ~40 messages per node, up to 9 blocks — no larger error measured).
  The velocity cancels at coordinate magnitude: 1e-6 xs is ~ 8 ulp of the coordinates.
End to end after 500 steps (config 1, `scale` = max |final coordinate| = 1 097; the reference's own fp32-vs-fp64
difference is 2.7e-3 = 2.4e-6 scale): fp32 max(3 x that, 1e-5 scale); tf32 / f16 3e-5, f16fast 5e-5, bf16 2e-4 of scale
(measured 1.6 - 2.7e-3 in every mode); types identical in every mode (measured), required: fp32 all, others >= 98 %.
BASELINE config 5's "bf16 MLP tiles vs TF32 accuracy check": test_config5_bf16_tiles_vs_tf32_accuracy.
"""
import os

import numpy as np
import pytest
import torch

from cmd_gen_b200.weights import init_weights, pack_blob
from oracle import diffphar_oracle as orc
from oracle import large_cases as lc
from tests.helpers import GOLDEN, load, T

DEV = "cuda:0"
TC_TOL = {"tf32": (5e-5, 0.02), "f16": (5e-5, 0.02), "f16fast": (5e-5, 0.03), "bf16": (5e-4, 0.05)}
DEEP = {"config2": 1.0, "config3": 1.0, "config5": 1.0}
PRECISIONS = ["fp32", "tf32", "f16", "f16fast", "bf16"]


def _fixture(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} missing (python -m oracle.make_golden_large)")
    return load(name)


def _dynamics_case(name):
    g = _fixture(f"large_dynamics_{name}.npz")
    d = lc.dynamics_inputs(name)
    # the regenerated inputs are the ones the reference saw
    assert np.allclose(lc.checksum(d["z"]), g["z_checksum"], rtol=0, atol=1e-6 * abs(g["z_checksum"][1]))
    assert np.allclose(lc.checksum(d["xh_pocket"]), g["pocket_checksum"], rtol=0, atol=1e-9 * abs(g["pocket_checksum"][1]))
    assert np.array_equal(d["z"].numpy(), g["z"])
    assert list(g["sizes"]) == d["sizes"] and list(g["counts"]) == d["counts"]
    return g, d


# ----------------------------------------------------------------------------- CPU: the oracle at these sizes
@pytest.mark.parametrize("name", ["config2", "config3", "config5"])
def test_oracle_edges_at_config_size(name):
    g, d = _dynamics_case(name)
    x = torch.cat([d["z"][:, :3], d["xh_pocket"][:, :3]])
    m = torch.cat([d["mask_phar"], d["mask_res"]])
    e = orc.exact_edges(m, x, d["cfg"].edge_cutoff).numpy()
    assert e.shape[1] == int(g["n_edges"])
    assert lc.edges_digest(e) == str(g["edges_exact_sha256"])
    assert int(g["mm_mode_disagreements"]) == 0 and str(g["edges_ref_sha256"]) == str(g["edges_exact_sha256"])
    assert np.array_equal(np.bincount(e[0], minlength=x.shape[0]), g["degrees_ref"])
    assert int((e[0] < d["z"].shape[0]).sum()) == int(g["n_edges_phar"])


@pytest.mark.parametrize("name,ti", [("config2", 1), ("config3", 0)])
def test_oracle_dynamics_at_config_size(name, ti):
    g, d = _dynamics_case(name)
    cfg = d["cfg"]
    W = init_weights(cfg, d["wseed"])
    B = len(d["sizes"])
    t = torch.full((B, 1), float(g["t_values"][ti]))
    with torch.no_grad():
        a, r, _ = orc.dynamics_forward(W, cfg, d["z"], d["xh_pocket"], t, d["mask_phar"], d["mask_res"])
    ref = g[f"eps_phar_f64_{ti}"]
    xs = float(max(d["z"][:, :3].abs().max(), d["xh_pocket"][:, :3].abs().max()))
    assert np.abs(a.numpy()[:, 3:] - ref[:, 3:]).max() <= 2e-5 * max(1.0, np.abs(ref[:, 3:]).max())
    assert np.abs(a.numpy()[:, :3] - ref[:, :3]).max() <= 1e-5 * xs
    # the oracle's fp32 result against the reference's fp32 result: same op sequence, summation order aside
    ref32 = g[f"eps_phar_f32_{ti}"]
    assert np.abs(a.numpy()[:, 3:] - ref32[:, 3:]).max() <= 5e-6
    if ti == 0:
        assert np.abs(r.numpy()[:, 3:] - g["eps_res_f64as32_0"]).max() <= 2e-5 * max(1.0, float(g["eps_res_absmax_0"]))


def test_oracle_sampler_config1_first_100_steps():
    """The oracle walks the first 100 of config 1's 500 steps (the GPU test walks all of them) and must sit on the
    reference's fp64 trajectory at steps 50 and 100 to the reference's own fp32 accuracy."""
    g = _fixture("large_sampler_config1.npz")
    d = lc.sampler_inputs("config1")
    assert np.allclose(lc.checksum(d["noise"]), g["noise_checksum"], rtol=0, atol=1e-6)
    assert np.allclose(lc.checksum(d["pocket"]["x"]), g["pocket_checksum"], rtol=0, atol=1e-6)
    from cmd_gen_b200.schedule import gamma_table, step_table
    cfg = d["cfg"]
    W = init_weights(cfg, d["wseed"])
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500)
    pocket, counts, noise = d["pocket"], torch.tensor(d["counts"]), d["noise"]
    B = len(d["counts"])
    mask_p = torch.repeat_interleave(torch.arange(B), counts)
    px = pocket["x"].clone()
    xh0 = torch.cat([px, pocket["one_hot"].float() / 4.0], 1)
    mu = torch.cat([orc._scatter_mean(px, pocket["mask"], B), torch.zeros(B, cfg.phar_nf)], 1)[mask_p]
    z, xh_pocket = orc.noise_and_center(mu, xh0, torch.ones(()), noise[0], mask_p, pocket["mask"], B)
    with torch.no_grad():
        for k in range(100):
            z, xh_pocket, _, _ = orc.ddpm_step(W, cfg, tab.rows[k], z, xh_pocket, noise[k + 1], mask_p, pocket["mask"], B)
            if (k + 1) % 50 == 0:
                ref = g["trace_z_f64"][(k + 1) // 50 - 1]
                scale = np.abs(ref[:, :3]).max()
                ref_err = np.abs(g["trace_z_f32"][(k + 1) // 50 - 1] - ref).max()
                assert np.abs(z.numpy() - ref).max() <= max(10 * ref_err, 1e-5 * scale), (k, ref_err)


# ----------------------------------------------------------------------------- GPU: all four precisions vs the reference
def _handle(cfg, wseed, prec):
    from cmd_gen_b200 import _lib
    h = _lib.Handle(cfg, DEV, prec)
    h.set_weights(pack_blob(cfg, init_weights(cfg, wseed)))
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("name", ["config2", "config3", "config5"])
def test_dynamics_at_config_size_vs_reference(name, prec):
    g, d = _dynamics_case(name)
    cfg = d["cfg"]
    h = _handle(cfg, d["wseed"], prec)
    h.plan(d["counts"], d["sizes"])
    B = len(d["sizes"])
    n_p = d["z"].shape[0]
    xs = float(max(d["z"][:, :3].abs().max(), d["xh_pocket"][:, :3].abs().max()))
    # K1 at this size against the reference's own list
    x = torch.cat([d["z"][:, :3], d["xh_pocket"][:, :3]]).to(DEV)
    rowptr, col = h.build_edges(x)
    deg = (rowptr[1:] - rowptr[:-1]).cpu().numpy()
    assert np.array_equal(deg, g["degrees_ref"])
    row = np.repeat(np.arange(deg.size), deg)
    assert lc.edges_digest(np.stack([row, col.cpu().numpy().astype(np.int64)])) == str(g["edges_ref_sha256"])
    worst = {}
    for i, tv in enumerate(g["t_values"]):
        out_p, out_r = h.dynamics_forward(d["z"], d["xh_pocket"], torch.full((B,), float(tv)), want_residues=(i == 0))
        fl = h.flags()
        assert fl.edge_overflow == 0 and fl.nan_resets == 0
        assert fl.last_n_edges == int(g["n_edges"]) and fl.last_n_edges_phar == int(g["n_edges_phar"])
        out_p = out_p.cpu().numpy()
        rp = g[f"eps_phar_f64_{i}"]
        hs, vs = max(1.0, float(np.abs(rp[:, 3:]).max())), float(np.abs(rp[:, :3]).max())
        eh = float(np.abs(out_p[:, 3:] - rp[:, 3:]).max())
        ex = float(np.abs(out_p[:, :3] - rp[:, :3]).max())
        if prec == "fp32":
            tol_h, tol_x = 2e-6 * hs, 1e-6 * xs
        else:
            tol_h, tol_x = DEEP[name] * TC_TOL[prec][0] * hs, 1e-6 * xs + TC_TOL[prec][1] * vs
        worst[float(tv)] = (eh / hs, ex, float(g[f"ref_err_h_{i}"]), float(g[f"ref_err_x_{i}"]))
        assert eh <= tol_h, (name, prec, tv, eh, tol_h)
        assert ex <= tol_x, (name, prec, tv, ex, tol_x)
        if i == 0:
            out_r = out_r.cpu().numpy()
            rs = max(1.0, float(g["eps_res_absmax_0"]))
            er = float(np.abs(out_r[:, 3:] - g["eps_res_f64as32_0"]).max())
            assert er <= (2e-6 if prec == "fp32" else DEEP[name] * TC_TOL[prec][0]) * rs, (name, prec, er)
            assert np.all(out_r[:, :3] == 0.0)
            worst["res"] = er / rs
    print(f"\n[parity] {name} {prec}: N={x.shape[0]} E={int(g['n_edges'])} worst (feature rel, velocity abs, ref fp32 err h, x) per t: {worst}")


@pytest.mark.gpu
@pytest.mark.parametrize("prec", PRECISIONS)
def test_config1_all_500_steps_vs_reference(prec):
    """BASELINE configs[0] — generate_phars on one pocket, 10 samples — through the mirrored
    ConditionalDDPM.sample_given_pocket with the reference's injected noise, all 500 steps in the captured graph,
    z compared with the reference's fp64 trajectory every 50 steps (return_frames = 10 written inside the loop)."""
    from tests.test_gpu_parity import build_ddpm, inject
    g = _fixture("large_sampler_config1.npz")
    d = lc.sampler_inputs("config1")
    assert np.allclose(lc.checksum(d["noise"]), g["noise_checksum"], rtol=0, atol=1e-6)
    cfg = d["cfg"]
    bound_rel = {"fp32": None, "tf32": 3e-5, "f16": 3e-5, "f16fast": 5e-5, "bf16": 2e-4}[prec]
    ref = g["xh_phar_f64"]
    scale = float(np.abs(ref[:, :3]).max())
    ref_err = float(np.abs(g["xh_phar_f32"][:, :3] - ref[:, :3]).max())
    bound = max(3 * ref_err, 1e-5 * scale) if prec == "fp32" else bound_rel * scale

    def run(return_frames):
        ddpm = build_ddpm(cfg, d["wseed"], 500, prec)
        inject(ddpm, d["noise"])
        pocket = {k: v.clone().to(DEV) for k, v in d["pocket"].items()}
        return ddpm.sample_given_pocket(pocket, torch.tensor(d["counts"]), return_frames=return_frames), ddpm

    (xh_phar, xh_pocket, mp, mr), ddpm = run(1)
    got = xh_phar.cpu().numpy()
    err = float(np.abs(got[:, :3] - ref[:, :3]).max())
    same = float((got[:, 3:] == g["xh_phar_f32"][:, 3:]).all(1).mean())
    perr = float(np.abs(xh_pocket[:, :3].cpu().numpy() - g["pocket_x_f64"]).max())
    print(f"\n[parity] config1 {prec}: 500 steps, scale {scale:.1f}, max |dx| {err:.3e} (bound {bound:.3e}, reference fp32 "
          f"vs fp64 {ref_err:.3e}), types identical {same:.2f}, pocket err {perr:.3e}")
    assert err <= bound and perr <= bound
    assert same == 1.0 if prec == "fp32" else same >= 0.98
    dxyz = got[:, :3] - ref[:, :3]
    for b in range(len(d["counts"])):                                   # north_star: per-sample RMSD
        assert np.sqrt((dxyz[g["mask_phar"] == b] ** 2).sum(1).mean()) <= bound
    assert ddpm.dynamics.handle(DEV).graph_captures() == 1

    # the trajectory: frames written from inside the captured loop against the reference's z every 50 steps
    (fr_phar, fr_pocket, _, _), _ = run(10)
    assert fr_phar.shape[0] == 10
    assert np.abs(fr_phar[0].cpu().numpy()[:, :3] - ref[:, :3]).max() <= bound
    worst_traj, worst_feat = 0.0, 0.0
    for idx in range(1, 10):
        rz = g["trace_z_f64"][9 - idx]                                  # frame idx holds s = 50 idx, i.e. call 500 - 50 idx
        zs = float(np.abs(rz[:, :3]).max())
        gz = fr_phar[idx].cpu().numpy()
        r_err = float(np.abs(g["trace_z_f32"][9 - idx] - rz).max())
        b_x = max(3 * r_err, 1e-5 * zs) if prec == "fp32" else max(3 * r_err, 2 * bound_rel * zs)
        e_x = float(np.abs(gz[:, :3] - rz[:, :3]).max())
        e_h = float(np.abs(gz[:, 3:] - 4.0 * rz[:, 3:]).max()) / max(1.0, float(np.abs(4 * rz[:, 3:]).max()))
        worst_traj = max(worst_traj, e_x / zs)
        worst_feat = max(worst_feat, e_h)
        assert e_x <= b_x, (idx, prec, e_x, b_x)
        assert e_h <= {"fp32": 1e-4, "tf32": 3e-3, "f16": 3e-3, "f16fast": 4e-3, "bf16": 1e-2}[prec], (idx, prec, e_h)
    print(f"[parity] config1 {prec}: trajectory every 50 steps: worst |dz_x| / scale {worst_traj:.3e}, worst feature error {worst_feat:.3e}")


@pytest.mark.gpu
def test_config5_bf16_tiles_vs_tf32_accuracy():
    """BASELINE configs[4]: 4k-node pockets, 12 phar points, 9 blocks — the bf16 MLP tiles against the TF32 tiles, both
    measured against the reference's fp64 output: TF32 (and f16, the same mantissa) must be the more accurate."""
    g, d = _dynamics_case("config5")
    cfg = d["cfg"]
    B = len(d["sizes"])
    rp = g["eps_phar_f64_0"]
    hs = max(1.0, float(np.abs(rp[:, 3:]).max()))
    err = {}
    for prec in ("tf32", "f16", "bf16"):
        h = _handle(cfg, d["wseed"], prec)
        h.plan(d["counts"], d["sizes"])
        out_p, out_r = h.dynamics_forward(d["z"], d["xh_pocket"], torch.full((B,), float(g["t_values"][0])))
        e_p = float(np.abs(out_p.cpu().numpy()[:, 3:] - rp[:, 3:]).max()) / hs
        e_r = float(np.abs(out_r.cpu().numpy()[:, 3:] - g["eps_res_f64as32_0"]).max()) / max(1.0, float(g["eps_res_absmax_0"]))
        err[prec] = max(e_p, e_r)
    print(f"\n[parity] config5 accuracy vs reference fp64 (max feature error / max |ref|): {err}")
    assert err["tf32"] < err["bf16"] and err["f16"] < err["bf16"]
    assert err["tf32"] <= 5e-5 and err["bf16"] <= 5e-4
