"""Drop-in for DiffPhar/equivariant_diffusion/dynamics.py::EGNNDynamics.

Same constructor signature, same state-dict keys (the weight ABI), same
``forward(xh_phars, xh_residues, t, mask_phars, mask_residues)`` and
``get_edges(batch_mask, x)`` — but the modules below only HOLD parameters; all
arithmetic runs in libdiffphar_b200.so through the C-ABI (no PyTorch fallback).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..config import DynamicsConfig, weight_spec
from ..weights import pack_blob
from .. import _lib


def _stack(sizes, final_bias=True):
    """Parameter container with nn.Linear at even positions (0, 2, 4, ...) so the
    state-dict keys match the reference's Sequential(Linear, act, Linear, ...)."""
    mods = []
    n = len(sizes) - 1
    for i in range(n):
        bias = final_bias or i < n - 1
        mods.append(nn.Linear(sizes[i], sizes[i + 1], bias=bias))
        if i < n - 1:
            mods.append(nn.Identity())
    return nn.Sequential(*mods)


class _ParamOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation lives in libdiffphar_b200.so")


def _egnn_container(cfg: DynamicsConfig) -> nn.Module:
    H, D = cfg.hidden_nf, cfg.node_nf
    egnn = _ParamOnly()
    egnn.embedding = nn.Linear(D, H)
    egnn.embedding_out = nn.Linear(H, D)
    for b in range(cfg.n_layers):
        blk = _ParamOnly()
        for g in range(cfg.inv_sublayers):
            gcl = _ParamOnly()
            gcl.edge_mlp = _stack([2 * H + 2, H, H])
            gcl.node_mlp = _stack([2 * H, H, H])
            if cfg.attention:
                gcl.att_mlp = _stack([H, 1])
            blk.add_module(f"gcl_{g}", gcl)
        eq = _ParamOnly()
        eq.coord_mlp = _stack([2 * H + 2, H, H, 1], final_bias=False)
        nn.init.xavier_uniform_(eq.coord_mlp[4].weight, gain=0.001)      # egnn_new.py:76-77
        blk.add_module("gcl_equiv", eq)
        egnn.add_module(f"e_block_{b}", blk)
    return egnn


class EGNNDynamics(nn.Module):
    def __init__(self, phar_nf, residue_nf, n_dims, joint_nf=16, hidden_nf=64, device='cpu',
                 act_fn=torch.nn.SiLU(), n_layers=4, attention=False, condition_time=True, tanh=False,
                 mode='egnn_dynamics', norm_constant=0, inv_sublayers=2, sin_embedding=False,
                 normalization_factor=100, aggregation_method='sum', update_pocket_coords=True,
                 edge_cutoff=None, precision='fp32'):
        super().__init__()
        if mode != 'egnn_dynamics':
            raise NotImplementedError(f"mode '{mode}' is outside the accelerated path (only 'egnn_dynamics')")
        if sin_embedding:
            raise NotImplementedError("sin_embedding=True is not enabled by any reference config; not built")
        if not isinstance(act_fn, torch.nn.SiLU):
            raise NotImplementedError("only SiLU activations are implemented (every reference config uses SiLU)")
        if hidden_nf != 256:
            raise NotImplementedError("hidden_nf must be 256 (compile-time tile width of the CUDA kernels)")
        self.mode = mode
        self.edge_cutoff = edge_cutoff
        self.cfg = DynamicsConfig(
            phar_nf=phar_nf, residue_nf=residue_nf, n_dims=n_dims, joint_nf=joint_nf, hidden_nf=hidden_nf,
            n_layers=n_layers, inv_sublayers=inv_sublayers, attention=bool(attention), tanh=bool(tanh),
            condition_time=bool(condition_time), norm_constant=float(norm_constant), coords_range=15.0,
            normalization_factor=float(normalization_factor), aggregation_method=aggregation_method,
            edge_cutoff=None if edge_cutoff is None else float(edge_cutoff))
        self.phar_encoder = _stack([phar_nf, 2 * phar_nf, joint_nf])
        self.phar_decoder = _stack([joint_nf, 2 * phar_nf, phar_nf])
        self.residue_encoder = _stack([residue_nf, 2 * residue_nf, joint_nf])
        self.residue_decoder = _stack([joint_nf, 2 * residue_nf, residue_nf])
        if not condition_time:
            print('Warning: dynamics model is _not_ conditioned on time.')
        self.egnn = _egnn_container(self.cfg)
        self.node_nf = self.cfg.node_nf
        self.update_pocket_coords = update_pocket_coords
        self.device = device
        self.n_dims = n_dims
        self.condition_time = condition_time
        self.precision = precision
        self._handle = None
        self._weights_tag = None
        self.to(device)

    # ------------------------------------------------------------------
    def _state_for_blob(self):
        return {k: v for k, v in self.state_dict().items()}

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def handle(self, device) -> "_lib.Handle":
        """The C-ABI handle for `device`, with the current parameters uploaded."""
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.DiffPharError("EGNNDynamics runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self._handle is None or self._handle.device != device:
            self._handle = _lib.Handle(self.cfg, device, self.precision)
            self._handle.set_update_pocket_coords(bool(self.update_pocket_coords))     # dynamics.py:104-107, 133-136
            self._weights_tag = None
        tag = self._weights_version()
        if tag != self._weights_tag:
            self._handle.set_weights(pack_blob(self.cfg, self._state_for_blob()))
            self._weights_tag = tag
        return self._handle

    def set_precision(self, precision: str):
        self.precision = precision
        if self._handle is not None:
            self._handle.set_precision(precision)

    @staticmethod
    def _counts(mask: torch.Tensor, n_samples: int):
        m = mask.detach().to('cpu', torch.int64)
        if m.numel() > 1 and bool((m[1:] < m[:-1]).any()):
            raise NotImplementedError(
                "batch masks must be sorted by sample (every reference caller builds them with "
                "repeat_interleave / num_nodes_to_batch_mask)")
        return torch.bincount(m, minlength=n_samples)

    # ------------------------------------------------------------------
    def forward(self, xh_phars=None, xh_residues=None, t=None, mask_phars=None, mask_residues=None,
                xh_atoms=None, mask_atoms=None):
        # BASELINE.json spells the phar arguments xh_atoms / mask_atoms; accept both.
        if xh_phars is None:
            xh_phars = xh_atoms
        if mask_phars is None:
            mask_phars = mask_atoms
        t = torch.as_tensor(t, device=xh_phars.device)
        n_samples = int(t.numel()) if t.numel() > 1 else \
            int(max(int(mask_phars.max()) if mask_phars.numel() else 0,
                    int(mask_residues.max()) if mask_residues.numel() else 0)) + 1
        h = self.handle(xh_phars.device)
        pc = self._counts(mask_phars, n_samples)
        rc = self._counts(mask_residues, n_samples)
        if pc.numel() != n_samples or rc.numel() != n_samples:
            raise ValueError("mask values exceed the number of samples implied by t")
        h.plan(pc, rc)
        out_p, out_r = h.dynamics_forward(xh_phars, xh_residues, t.to(torch.float32), want_residues=True)
        fl = h.flags()
        if fl.edge_overflow:
            # the device truncated the graph at the planned capacity (memory-safe, results invalid) and reported the
            # edge count it needs: grow the plan once and evaluate again
            h.reset_flags()
            h.grow_edge_capacity(int(fl.edge_overflow * 1.25) + 1024)
            out_p, out_r = h.dynamics_forward(xh_phars, xh_residues, t.to(torch.float32), want_residues=True)
            fl = h.flags()
            if fl.edge_overflow:
                raise _lib.DiffPharError("edge buffer overflow persists after re-planning")
        if fl.f16_range and self.precision not in ("fp32", "tf32"):
            print(f'Warning: f16 range exceeded (flag {fl.f16_range}); re-running with tf32 tensor-core tiles.')
            h.reset_flags()
            self.set_precision("tf32")
            out_p, out_r = h.dynamics_forward(xh_phars, xh_residues, t.to(torch.float32), want_residues=True)
            fl = h.flags()
        if fl.nan_resets:
            print('Warning: detected nan, resetting EGNN output to zero.')
            h.reset_flags()
        return out_p, out_r

    def get_edges(self, batch_mask, x):
        """[2, E] int64, sorted by (row, col); same-sample pairs within edge_cutoff incl. self loops."""
        m = batch_mask.detach().to('cpu', torch.int64)
        n = m.numel()
        drops = torch.nonzero(m[1:] < m[:-1]).reshape(-1) + 1 if n > 1 else torch.zeros(0, dtype=torch.int64)
        if drops.numel() > 1:
            raise NotImplementedError("batch_mask must be one or two sorted runs (phar nodes, then pocket nodes)")
        split = int(drops[0]) if drops.numel() == 1 else 0
        ids = torch.unique(m)
        remap = {int(v): i for i, v in enumerate(ids.tolist())}
        dense = torch.tensor([remap[int(v)] for v in m.tolist()], dtype=torch.int64)
        B = ids.numel()
        pc = torch.bincount(dense[:split], minlength=B)
        rc = torch.bincount(dense[split:], minlength=B)
        h = self.handle(x.device)
        h.plan(pc, rc)
        rowptr, col = h.build_edges(x)
        deg = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
        row = torch.repeat_interleave(torch.arange(n, device=x.device, dtype=torch.int64), deg)
        return torch.stack([row, col.to(torch.int64)], dim=0)
