"""Deterministic random-init weights and blob packing for the weight ABI.

There is no checkpoint in the reference tree and no network, so benchmarks and
parity tests use random-init weights of the configured architecture.  The
distributions are PyTorch's defaults for ``nn.Linear`` (uniform +-1/sqrt(fan_in)
for weight and bias) and the reference's xavier(gain=1e-3) for the last
coordinate layer (egnn_new.py:76-77), drawn from OUR OWN seeded generator so
the same tensors exist on the GPU box where /root/reference does not.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import DynamicsConfig, weight_spec


def init_weights(cfg: DynamicsConfig, seed: int = 0, coord_gain: float = 1e-3,
                 dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for key, shape in weight_spec(cfg):
        if key.endswith("coord_mlp.4.weight"):
            fan_out, fan_in = shape
            bound = coord_gain * math.sqrt(6.0 / (fan_in + fan_out))
        elif key.endswith(".weight"):
            bound = 1.0 / math.sqrt(shape[1])
        else:  # bias: fan_in of the matching weight
            wshape = dict(weight_spec(cfg))[key[:-4] + "weight"]
            bound = 1.0 / math.sqrt(wshape[1])
        t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0) * bound
        out[key] = t.to(dtype)
    return out


def pack_blob(cfg: DynamicsConfig, state: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Flatten a state dict (keys relative to EGNNDynamics) into the canonical
    fp32 blob ``dp_set_weights`` expects.  Raises KeyError/ValueError on a
    missing key or shape mismatch, like ``load_state_dict(strict=True)``."""
    parts = []
    for key, shape in weight_spec(cfg):
        if key not in state:
            raise KeyError(f"missing weight '{key}'")
        t = state[key].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"weight '{key}' has shape {tuple(t.shape)}, expected {shape}")
        parts.append(t.reshape(-1))
    return torch.cat(parts).contiguous()
