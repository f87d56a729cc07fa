"""Architecture description of the DiffPhar denoiser + the weight ABI.

The state-dict key names and ``[out, in]`` row-major fp32 shapes below are the
weight ABI of the reference checkpoints (reference ctor code at
DiffPhar/equivariant_diffusion/dynamics.py:21-43 and egnn_new.py:15-29,69-85,
159-191).  ``weight_spec`` fixes ONE canonical order for them; the C-ABI call
``dp_set_weights`` (include/diffphar_b200.h) takes a flat fp32 blob in exactly
this order.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import List, Tuple

# "f16fast": f16 operands, packed f16x2 first layer + tanh-form SiLU in the edge kernels (bench default);
# "f16fast32": its A/B variant with the producer's tanh in fp32
PRECISION_MODES = {"fp32": 0, "tf32": 1, "bf16": 2, "f16": 3, "f16fast": 4, "f16fast32": 5}


@dataclass(frozen=True)
class DynamicsConfig:
    phar_nf: int = 8
    residue_nf: int = 20
    n_dims: int = 3
    joint_nf: int = 32
    hidden_nf: int = 256
    n_layers: int = 5
    inv_sublayers: int = 1
    attention: bool = True
    tanh: bool = True
    condition_time: bool = True
    norm_constant: float = 1.0
    coords_range: float = 15.0          # egnn_new.py:187 passes the raw range to every block
    normalization_factor: float = 100.0
    aggregation_method: str = "sum"
    edge_cutoff: float | None = 6.0

    @property
    def node_nf(self) -> int:
        return self.joint_nf + (1 if self.condition_time else 0)

    def as_dict(self):
        return asdict(self)


def weight_spec(cfg: DynamicsConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Canonical (key, shape) list.  Keys are relative to the EGNNDynamics
    module (i.e. what follows ``ddpm.dynamics.`` in a Lightning checkpoint)."""
    P, R, J, H = cfg.phar_nf, cfg.residue_nf, cfg.joint_nf, cfg.hidden_nf
    D = cfg.node_nf
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def lin(name, out_f, in_f, bias=True):
        spec.append((f"{name}.weight", (out_f, in_f)))
        if bias:
            spec.append((f"{name}.bias", (out_f,)))

    lin("phar_encoder.0", 2 * P, P)
    lin("phar_encoder.2", J, 2 * P)
    lin("phar_decoder.0", 2 * P, J)
    lin("phar_decoder.2", P, 2 * P)
    lin("residue_encoder.0", 2 * R, R)
    lin("residue_encoder.2", J, 2 * R)
    lin("residue_decoder.0", 2 * R, J)
    lin("residue_decoder.2", R, 2 * R)
    lin("egnn.embedding", H, D)
    lin("egnn.embedding_out", D, H)
    for b in range(cfg.n_layers):
        pre = f"egnn.e_block_{b}"
        for g in range(cfg.inv_sublayers):
            lin(f"{pre}.gcl_{g}.edge_mlp.0", H, 2 * H + 2)
            lin(f"{pre}.gcl_{g}.edge_mlp.2", H, H)
            lin(f"{pre}.gcl_{g}.node_mlp.0", H, 2 * H)
            lin(f"{pre}.gcl_{g}.node_mlp.2", H, H)
            if cfg.attention:
                lin(f"{pre}.gcl_{g}.att_mlp.0", 1, H)
        lin(f"{pre}.gcl_equiv.coord_mlp.0", H, 2 * H + 2)
        lin(f"{pre}.gcl_equiv.coord_mlp.2", H, H)
        lin(f"{pre}.gcl_equiv.coord_mlp.4", 1, H, bias=False)
    return spec


def weight_count(cfg: DynamicsConfig) -> int:
    n = 0
    for _, shp in weight_spec(cfg):
        k = 1
        for s in shp:
            k *= s
        n += k
    return n
