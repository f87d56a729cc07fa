"""ctypes binding of libdiffphar_b200.so (include/diffphar_b200.h).

There is NO fallback: if the shared library is missing, or no sm_100 GPU is
visible when a compute entry point is called, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .config import DynamicsConfig, PRECISION_MODES

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffphar_b200.so")

EXPORTS = [
    "dp_last_error", "dp_abi_version", "dp_device_count", "dp_create", "dp_destroy", "dp_weight_count",
    "dp_set_weights", "dp_set_precision", "dp_plan", "dp_build_edges", "dp_get_graph", "dp_dynamics_forward",
    "dp_ddpm_update", "dp_set_step_table", "dp_sample", "dp_sample_host", "dp_get_flags", "dp_reset_flags",
    "dp_launch_count", "dp_profile_enable", "dp_profile_read", "dp_pointcloud_stats",
    "dp_sample_ex", "dp_fill_noise", "dp_sample_host_seeded", "dp_graph_captures", "dp_set_update_pocket_coords",
]


class DpConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("phar_nf", "residue_nf", "n_dims", "joint_nf", "hidden_nf", "n_layers",
                                         "inv_sublayers", "attention", "use_tanh", "condition_time",
                                         "aggregation_mean")] + \
               [(n, C.c_float) for n in ("norm_constant", "coords_range", "normalization_factor", "edge_cutoff")] + \
               [("precision", C.c_int32)]


class DpFlags(C.Structure):
    _fields_ = [("nan_resets", C.c_int32), ("edge_overflow", C.c_int32), ("max_mean_rel_err", C.c_float),
                ("last_max_cog", C.c_float), ("last_n_edges", C.c_int64), ("last_n_edges_phar", C.c_int64),
                ("f16_range", C.c_int32), ("reserved", C.c_int32)]


class DpSampleOpts(C.Structure):
    _fields_ = [("noise_dev", C.c_void_p), ("seed", C.c_uint64), ("sample_ids_host", C.c_void_p),
                ("return_frames", C.c_int32), ("norm_x", C.c_float), ("norm_h", C.c_float), ("bias_h", C.c_float),
                ("frames_phar_dev", C.c_void_p), ("frames_pocket_dev", C.c_void_p)]


class DiffPharError(RuntimeError):
    pass


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DIFFPHAR_LIB", LIB_PATH)      # A/B of two builds on one box (scripts/gpu_env_ab.sh)
    if not os.path.exists(path):
        raise DiffPharError(
            f"{path} is missing: build it with `python -m cmd_gen_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.dp_last_error.restype = C.c_char_p
    lib.dp_abi_version.restype = C.c_int
    lib.dp_device_count.restype = C.c_int
    lib.dp_create.argtypes = [C.POINTER(DpConfig), C.c_int, C.POINTER(vp)]
    lib.dp_destroy.argtypes = [vp]
    lib.dp_weight_count.argtypes = [vp]
    lib.dp_weight_count.restype = i64
    lib.dp_set_weights.argtypes = [vp, vp, i64]
    lib.dp_set_precision.argtypes = [vp, C.c_int]
    lib.dp_set_update_pocket_coords.argtypes = [vp, C.c_int32]
    lib.dp_plan.argtypes = [vp, i32, vp, vp, i64]
    lib.dp_build_edges.argtypes = [vp, vp, vp]
    lib.dp_get_graph.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), vp]
    lib.dp_dynamics_forward.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    lib.dp_ddpm_update.argtypes = [vp, i32, f32, f32, f32, vp, vp, vp, vp, vp]
    lib.dp_set_step_table.argtypes = [vp, vp, i32, vp]
    lib.dp_sample.argtypes = [vp, vp, vp, vp, vp]
    lib.dp_sample_host.argtypes = [vp, vp, vp, vp, vp]
    lib.dp_get_flags.argtypes = [vp, C.POINTER(DpFlags), vp]
    lib.dp_reset_flags.argtypes = [vp, vp]
    lib.dp_launch_count.argtypes = [vp]
    lib.dp_launch_count.restype = i64
    lib.dp_profile_enable.argtypes = [vp, i32]
    lib.dp_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.dp_pointcloud_stats.argtypes = [vp, vp, i32, C.POINTER(C.c_double), vp, vp]
    lib.dp_sample_ex.argtypes = [vp, vp, C.POINTER(DpSampleOpts), vp, vp]
    lib.dp_fill_noise.argtypes = [vp, C.c_uint64, vp, i32, vp, vp]
    lib.dp_sample_host_seeded.argtypes = [vp, vp, C.c_uint64, vp, vp, vp]
    lib.dp_graph_captures.argtypes = [vp]
    lib.dp_graph_captures.restype = i64
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise DiffPharError(f"libdiffphar_b200 error {rc}: {load_library().dp_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class Handle:
    """Owns one dp_handle (one device, one stream at a time)."""

    def __init__(self, cfg: DynamicsConfig, device, precision: str = "fp32"):
        self.lib = load_library()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise DiffPharError("the DiffPhar B200 path runs on CUDA devices only (no CPU fallback)")
        if self.lib.dp_device_count() == 0:
            raise DiffPharError("no sm_100 (B200) device visible; there is no fallback path")
        if cfg.aggregation_method not in ("sum", "mean"):
            raise ValueError(cfg.aggregation_method)
        c = DpConfig(cfg.phar_nf, cfg.residue_nf, cfg.n_dims, cfg.joint_nf, cfg.hidden_nf, cfg.n_layers,
                     cfg.inv_sublayers, int(cfg.attention), int(cfg.tanh), int(cfg.condition_time),
                     int(cfg.aggregation_method == "mean"), float(cfg.norm_constant), float(cfg.coords_range),
                     float(cfg.normalization_factor),
                     -1.0 if cfg.edge_cutoff is None else float(cfg.edge_cutoff), PRECISION_MODES[precision])
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _check(self.lib.dp_create(C.byref(c), idx, C.byref(h)))
        self.h = h
        self.layout = None
        self._grown = False
        self._keep = []

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.dp_destroy(h)
            except Exception:
                pass

    # ---- configuration -------------------------------------------------
    def weight_count(self) -> int:
        return int(self.lib.dp_weight_count(self.h))

    def set_weights(self, blob: torch.Tensor):
        blob = blob.detach().to("cpu", torch.float32).contiguous()
        _check(self.lib.dp_set_weights(self.h, _ptr(blob), blob.numel()))

    def set_precision(self, precision: str):
        _check(self.lib.dp_set_precision(self.h, PRECISION_MODES[precision]))

    def set_update_pocket_coords(self, on: bool):
        """EGNNDynamics(update_pocket_coords=...): joint mode moves (and returns velocities for) the pocket nodes too."""
        _check(self.lib.dp_set_update_pocket_coords(self.h, 1 if on else 0))

    def plan(self, phar_counts, res_counts, edge_capacity: int = 0):
        pc = torch.as_tensor(phar_counts, dtype=torch.int32, device="cpu").contiguous()
        rc = torch.as_tensor(res_counts, dtype=torch.int32, device="cpu").contiguous()
        assert pc.numel() == rc.numel()
        key = (tuple(pc.tolist()), tuple(rc.tolist()), int(edge_capacity))
        if key == self.layout:
            return
        if self.layout is not None and key[:2] == self.layout[:2] and edge_capacity == 0 and self._grown:
            return          # the automatic capacity overflowed on this layout before: keep the grown plan
        _check(self.lib.dp_plan(self.h, pc.numel(), _ptr(pc), _ptr(rc), int(edge_capacity)))
        self.layout = key
        self._grown = False
        self.n_phar, self.n_res = int(pc.sum()), int(rc.sum())

    def grow_edge_capacity(self, edge_capacity: int):
        """Re-plan the current layout with a larger edge buffer (after dp_flags.edge_overflow reported the need)."""
        pc, rc, _ = self.layout
        self.plan(pc, rc, edge_capacity=int(edge_capacity))
        self._grown = True

    def set_step_table(self, rows: torch.Tensor, final: torch.Tensor):
        rows = rows.detach().to("cpu", torch.float32).contiguous()
        final = final.detach().to("cpu", torch.float32).contiguous()
        _check(self.lib.dp_set_step_table(self.h, _ptr(rows), rows.shape[0], _ptr(final)))

    # ---- compute ---------------------------------------------------------
    def build_edges(self, x: torch.Tensor):
        """x [N,3] on the device -> (rowptr int32 [N+1], col int32 [E]) views of handle-owned memory (cloned)."""
        x = _dev_f32(x, self.device)
        st = _stream(self.device)
        _check(self.lib.dp_build_edges(self.h, _ptr(x), st))
        rp, cp, ne = C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(self.lib.dp_get_graph(self.h, C.byref(rp), C.byref(cp), C.byref(ne), st))
        n = x.shape[0]
        rowptr = _from_ptr(rp.value, n + 1, torch.int32, self.device).clone()
        col = _from_ptr(cp.value, ne.value, torch.int32, self.device).clone()
        return rowptr, col

    def dynamics_forward(self, xh_phar, xh_res, t, want_residues=True):
        xh_phar, xh_res = _dev_f32(xh_phar, self.device), _dev_f32(xh_res, self.device)
        t = _dev_f32(t.reshape(-1), self.device)
        t_stride = 0 if t.numel() == 1 else 1
        out_p = torch.empty_like(xh_phar)
        out_r = torch.empty_like(xh_res) if want_residues else None
        _check(self.lib.dp_dynamics_forward(self.h, _ptr(xh_phar), _ptr(xh_res), _ptr(t), t_stride,
                                            _ptr(out_p), _ptr(out_r), _stream(self.device)))
        return out_p, out_r

    def ddpm_update(self, kind, a, c, sigma, z, pocket, eps_hat, noise):
        """In place on z [N_p, 3+P] and pocket [N_r, 3+R] (both contiguous fp32 on the device)."""
        for t in (z, pocket):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        eps_hat = None if eps_hat is None else _dev_f32(eps_hat, self.device)
        noise = _dev_f32(noise, self.device)
        _check(self.lib.dp_ddpm_update(self.h, int(kind), float(a), float(c), float(sigma), _ptr(z), _ptr(pocket),
                                       _ptr(eps_hat), _ptr(noise), _stream(self.device)))

    def sample(self, xh_pocket: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0, sample_ids=None,
               return_frames: int = 1, norm=(1.0, 1.0, 0.0)):
        """xh_pocket (normalised, device, modified in place) -> out_phar [N_p, 3+P] = (x_final | z0 features).

        noise [n_steps+2, N_p, 3+P] injects the gaussian draws; noise=None draws them on the device from the
        counter-based generator keyed by (seed, global sample id) — sample_ids [n_samples] int64, default 0..n-1.
        return_frames > 1 also returns (frames_phar [F, N_p, 3+P], frames_pocket [F, N_r, 3+R]), un-normalised with
        norm = (norm_values[0], norm_values[1], norm_biases[1]); frame 0 is for the caller to fill."""
        assert xh_pocket.is_cuda and xh_pocket.dtype == torch.float32 and xh_pocket.is_contiguous()
        out = torch.empty((self.n_phar, 3 + self.cfg.phar_nf), device=self.device, dtype=torch.float32)
        o = DpSampleOpts()
        keep = []
        if noise is not None:
            noise = _dev_f32(noise, self.device)
            keep.append(noise)
            o.noise_dev = noise.data_ptr()
        else:
            o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            if sample_ids is not None:
                ids = torch.as_tensor(sample_ids, dtype=torch.int64, device="cpu").contiguous()
                keep.append(ids)
                o.sample_ids_host = ids.data_ptr()
        o.return_frames = int(return_frames)
        frames = None
        if return_frames > 1:
            o.norm_x, o.norm_h, o.bias_h = (float(v) for v in norm)
            frames = (torch.zeros((return_frames, self.n_phar, 3 + self.cfg.phar_nf), device=self.device),
                      torch.zeros((return_frames, self.n_res, 3 + self.cfg.residue_nf), device=self.device))
            o.frames_phar_dev, o.frames_pocket_dev = frames[0].data_ptr(), frames[1].data_ptr()
        self._keep = keep
        _check(self.lib.dp_sample_ex(self.h, _ptr(xh_pocket), C.byref(o), _ptr(out), _stream(self.device)))
        return out if frames is None else (out, frames[0], frames[1])

    def fill_noise(self, n_draws: int, seed: int, sample_ids=None) -> torch.Tensor:
        """[n_draws, N_p, 3+P] draws of the device generator (what sample(noise=None, seed=...) consumes)."""
        out = torch.empty((n_draws, self.n_phar, 3 + self.cfg.phar_nf), device=self.device, dtype=torch.float32)
        ids = None if sample_ids is None else torch.as_tensor(sample_ids, dtype=torch.int64, device="cpu").contiguous()
        _check(self.lib.dp_fill_noise(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(ids), int(n_draws), _ptr(out),
                                      _stream(self.device)))
        return out

    def sample_host_seeded(self, xh_pocket_host, seed, out_phar_host, pocket_out_host=None, sample_ids=None):
        for t in (xh_pocket_host, out_phar_host):
            assert (not t.is_cuda) and t.dtype == torch.float32 and t.is_contiguous()
        ids = None if sample_ids is None else torch.as_tensor(sample_ids, dtype=torch.int64, device="cpu").contiguous()
        _check(self.lib.dp_sample_host_seeded(self.h, _ptr(xh_pocket_host), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(ids),
                                              _ptr(out_phar_host), _ptr(pocket_out_host)))

    def graph_captures(self) -> int:
        return int(self.lib.dp_graph_captures(self.h))

    def sample_host(self, xh_pocket_host: torch.Tensor, noise_host: torch.Tensor, out_phar_host: torch.Tensor,
                    pocket_out_host: Optional[torch.Tensor] = None):
        for t in (xh_pocket_host, noise_host, out_phar_host):
            assert (not t.is_cuda) and t.dtype == torch.float32 and t.is_contiguous()
        _check(self.lib.dp_sample_host(self.h, _ptr(xh_pocket_host), _ptr(noise_host), _ptr(out_phar_host),
                                       _ptr(pocket_out_host)))

    def flags(self) -> DpFlags:
        f = DpFlags()
        _check(self.lib.dp_get_flags(self.h, C.byref(f), _stream(self.device)))
        return f

    def reset_flags(self):
        _check(self.lib.dp_reset_flags(self.h, _stream(self.device)))

    def launch_count(self) -> int:
        return int(self.lib.dp_launch_count(self.h))

    def profile_enable(self, on):
        """on: False / True = CUDA events around every EAGER launch; 2 = events recorded inside the captured step graph
        (external event-record nodes, read after every replay): launch durations in the production context."""
        _check(self.lib.dp_profile_enable(self.h, int(on)))

    def profile_read(self, which: int):
        ms, n = C.c_double(), C.c_int64()
        _check(self.lib.dp_profile_read(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def _from_ptr(ptr: int, count: int, dtype, device) -> torch.Tensor:
    """Wrap handle-owned device memory as a tensor (via __cuda_array_interface__)."""
    if count == 0:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int32: "<i4", torch.float32: "<f4"}[dtype]

    class _Wrap:
        __cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Wrap(), device=device)
