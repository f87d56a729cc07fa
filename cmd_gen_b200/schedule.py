"""Noise schedule and the per-step constant table of the reverse diffusion.

Follows (restated, not copied) the reference arithmetic:
  * gamma table: en_diffusion.py:1119-1149 (clip + polynomial), :1099-1116 (cosine),
    :1152-1184 (log-ratio -> float32 Parameter), lookup :1186-1188;
  * sigma/alpha given s: en_diffusion.py:79-103; sigma/alpha/SNR: :859-872;
  * the way sample_given_pocket forms s/t: conditional_model.py:428-433;
  * how sample_p_zs_given_zt combines them: conditional_model.py:345-366;
  * the final p(x|z0) constants: conditional_model.py:111-119, en_diffusion.py:153-165.

Every sample in a batch shares s and t while sampling, so the [B,1] tensors of
the reference collapse to scalars per step.  The table is produced ON THE HOST
with the same torch fp32 op sequence the reference uses, so the constants fed to
the fused DDPM kernel are bit-identical to the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


def _clip_alphas2(alphas2: np.ndarray, clip_value: float = 0.001) -> np.ndarray:
    padded = np.concatenate([np.ones(1), alphas2], axis=0)
    ratio = np.clip(padded[1:] / padded[:-1], a_min=clip_value, a_max=1.0)
    return np.cumprod(ratio, axis=0)


def polynomial_alphas2(timesteps: int, s: float = 1e-4, power: float = 3.0) -> np.ndarray:
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    a2 = (1 - np.power(x / steps, power)) ** 2
    a2 = _clip_alphas2(a2, clip_value=0.001)
    return (1 - 2 * s) * a2 + s


def cosine_alphas2(timesteps: int, s: float = 0.008, raise_to_power: float = 1) -> np.ndarray:
    steps = timesteps + 2
    x = np.linspace(0, steps, steps)
    cum = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    cum = cum / cum[0]
    betas = np.clip(1 - (cum[1:] / cum[:-1]), a_min=0, a_max=0.999)
    cum = np.cumprod(1.0 - betas, axis=0)
    if raise_to_power != 1:
        cum = np.power(cum, raise_to_power)
    return cum


def gamma_table(noise_schedule: str, timesteps: int, precision: float) -> torch.Tensor:
    """float32 gamma[T+1] = -(log alpha^2 - log sigma^2)."""
    if noise_schedule == "cosine":
        a2 = cosine_alphas2(timesteps)
    elif "polynomial" in noise_schedule:
        parts = noise_schedule.split("_")
        if len(parts) != 2:
            raise ValueError(noise_schedule)
        a2 = polynomial_alphas2(timesteps, s=precision, power=float(parts[1]))
    else:
        raise ValueError(noise_schedule)
    log_ratio = np.log(a2) - np.log(1 - a2)
    return torch.from_numpy(-log_ratio).float()


@dataclass
class StepTable:
    """rows[k] for k = 0..n_steps-1 is the k-th EXECUTED step (s = n_steps-1-k):
    columns (t, alpha_ts, c_eps, sigma_noise).  final = (t0=0, inv_alpha0, sigma0, sigma_x)."""
    rows: torch.Tensor      # [n_steps, 4] float32
    final: torch.Tensor     # [4] float32
    n_steps: int


def step_table(gamma: torch.Tensor, T: int, timesteps: int | None = None,
               dtype: torch.dtype = torch.float32) -> StepTable:
    """dtype=float64 reproduces what the reference computes after ``.double()``
    (the fp32 gamma values widened; used only by the fp64 accuracy yardstick)."""
    gamma = gamma.detach().to("cpu", torch.float32).to(dtype)
    n = T if timesteps is None else int(timesteps)
    def lookup(t):
        return gamma[torch.round(t * T).long()]

    # One [1,1] evaluation per step, like the reference's per-step [B,1] calls:
    # torch's vectorised transcendental kernels differ from these in the last ulp
    # when the 500 steps are batched into one tensor (observed), so do not batch.
    out_rows = []
    for s in range(n - 1, -1, -1):
        s_arr = torch.full((1, 1), fill_value=s)          # int64, as conditional_model.py:429
        t_arr = s_arr + 1
        s_arr = s_arr / n                                  # int64 / python int -> float32
        t_arr = t_arr / n
        g_s, g_t = lookup(s_arr), lookup(t_arr)
        sigma2_ts = -torch.expm1(F.softplus(g_s) - F.softplus(g_t))
        alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
        sigma_ts = torch.sqrt(sigma2_ts)
        sigma_s = torch.sqrt(torch.sigmoid(g_s))
        sigma_t = torch.sqrt(torch.sigmoid(g_t))
        c_eps = sigma2_ts / alpha_ts / sigma_t
        sigma = sigma_ts * sigma_s / sigma_t
        out_rows.append(torch.cat([t_arr.to(dtype), alpha_ts, c_eps, sigma], dim=1))
    rows = torch.cat(out_rows, dim=0).contiguous()

    t0 = torch.zeros(1, 1)
    g0 = lookup(t0)
    sigma_x = torch.exp(-(-0.5 * g0))
    sigma0 = torch.sqrt(torch.sigmoid(g0))
    alpha0 = torch.sqrt(torch.sigmoid(-g0))
    inv_alpha0 = 1.0 / alpha0
    final = torch.cat([t0.to(dtype), inv_alpha0, sigma0, sigma_x], dim=1).reshape(4).contiguous()
    return StepTable(rows=rows, final=final, n_steps=n)
