"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference modules for golden-vector
generation.  Only usable where /root/reference exists (the build container); it
never travels to the GPU box and nothing in the product imports it.

The reference's hot path imports three packages this image does not have
(SURVEY.md §8c).  We inject minimal stand-ins into ``sys.modules`` *before*
importing the reference so that its own source runs unchanged:

  * ``torch_scatter.scatter_add / scatter_mean`` (torch-scatter==2.0.9,
    DiffPhar/env/environment_diffphar.yml:236; call sites conditional_model.py:412,
    452,471, en_diffusion.py:915,922,940).  Published semantics: index-sum over
    ``dim`` with ``dim_size = index.max()+1``; mean = sum / count.clamp(min=1).
  * ``rdkit``, ``Bio`` — imported by DiffPhar/utils.py:6,9 but unused on the path.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("DIFFPHAR_REFERENCE", "/root/reference/DiffPhar")


def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    total = _scatter_add(src, index, dim=dim, dim_size=dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype, device=src.device)
    count = _scatter_add(ones, index, dim=0, dim_size=total.shape[0]).clamp_(min=1)
    return total / count.view((-1,) + (1,) * (src.dim() - 1))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "equivariant_diffusion"))


def install():
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter_add = _scatter_add
        ts.scatter_mean = _scatter_mean
        sys.modules["torch_scatter"] = ts
    for name in ("rdkit", "rdkit.Chem", "Bio", "Bio.PDB", "Bio.PDB.Polypeptide"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["Bio"].PDB = sys.modules["Bio.PDB"]
    sys.modules["Bio.PDB"].Polypeptide = sys.modules["Bio.PDB.Polypeptide"]
    sys.modules["Bio.PDB.Polypeptide"].is_aa = lambda *a, **k: True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Returns (EGNNDynamics, ConditionalDDPM) classes of the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from equivariant_diffusion.dynamics import EGNNDynamics            # noqa
        from equivariant_diffusion.conditional_model import ConditionalDDPM  # noqa
    return EGNNDynamics, ConditionalDDPM


def build_reference_model(cfg, state, T=500, noise_schedule="polynomial_2", precision=1e-5,
                          norm_values=(1.0, 4.0), dtype=torch.float32, size_histogram=None):
    """Instantiate reference EGNNDynamics + ConditionalDDPM with OUR weights."""
    import contextlib
    import io
    EGNNDynamics, ConditionalDDPM = load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        dyn = EGNNDynamics(
            phar_nf=cfg.phar_nf, residue_nf=cfg.residue_nf, n_dims=cfg.n_dims,
            joint_nf=cfg.joint_nf, hidden_nf=cfg.hidden_nf, device="cpu",
            act_fn=torch.nn.SiLU(), n_layers=cfg.n_layers, attention=cfg.attention,
            condition_time=cfg.condition_time, tanh=cfg.tanh, mode="egnn_dynamics",
            norm_constant=cfg.norm_constant, inv_sublayers=cfg.inv_sublayers,
            sin_embedding=False, normalization_factor=cfg.normalization_factor,
            aggregation_method=cfg.aggregation_method, update_pocket_coords=False,
            edge_cutoff=cfg.edge_cutoff)
        dyn.load_state_dict({k: v.clone() for k, v in state.items()}, strict=True)
        ddpm = ConditionalDDPM(
            dynamics=dyn, phar_nf=cfg.phar_nf, residue_nf=cfg.residue_nf, n_dims=cfg.n_dims,
            size_histogram=[[1.0, 1.0], [1.0, 1.0]] if size_histogram is None else size_histogram, timesteps=T, parametrization="eps",
            noise_schedule=noise_schedule, noise_precision=precision, loss_type="l2",
            norm_values=norm_values, norm_biases=(None, 0.0))
    if dtype == torch.float64:
        ddpm = ddpm.double()
    return ddpm.eval()


class InjectedNoise:
    """Replaces ``sample_gaussian`` (en_diffusion.py:946-949) by a FIFO over a
    pre-drawn tensor so both implementations consume identical noise."""

    def __init__(self, ddpm, noise):
        self.ddpm, self.noise, self.k = ddpm, noise, 0

    def __enter__(self):
        def pop(size, device):
            out = self.noise[self.k].to(device)
            assert tuple(out.shape) == tuple(size), (out.shape, size)
            self.k += 1
            return out.clone()
        self._saved = self.ddpm.sample_gaussian
        self.ddpm.sample_gaussian = pop
        return self

    def __exit__(self, *exc):
        self.ddpm.sample_gaussian = self._saved
        return False
