"""Drop-in for DiffPhar/equivariant_diffusion/conditional_model.py::ConditionalDDPM —
the sampling half (training losses are outside the accelerated path).

``sample_given_pocket`` keeps the reference signature, return tuple, caller-dict mutation
and assertion/print behaviour (conditional_model.py:388-465) but runs the whole reverse
diffusion inside libdiffphar_b200.so: ONE call (dp_sample) replays a CUDA graph of the
denoising step; the reference's per-step host syncs become device-side flags read once.
The per-step methods (``sample_p_zs_given_zt`` ...) are kept for callers that drive the
loop themselves; they call the same kernels one step at a time.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _lib
from ..schedule import step_table
from ..utils import num_nodes_to_batch_mask, scatter_add, scatter_mean
from .en_diffusion import DistributionNodes, PredefinedNoiseSchedule, ScheduleMixin


class ConditionalDDPM(torch.nn.Module, ScheduleMixin):
    def __init__(self, dynamics, phar_nf, residue_nf, n_dims, size_histogram, timesteps=1000,
                 parametrization='eps', noise_schedule='learned', noise_precision=1e-4, loss_type='vlb',
                 norm_values=(1., 1.), norm_biases=(None, 0.)):
        super().__init__()
        assert loss_type in {'vlb', 'l2'}
        assert parametrization == 'eps'
        if noise_schedule == 'learned':
            raise NotImplementedError("learned noise schedules are outside the accelerated sampling path")
        self.loss_type = loss_type
        self.gamma = PredefinedNoiseSchedule(noise_schedule, timesteps=timesteps, precision=noise_precision)
        self.dynamics = dynamics
        self.phar_nf, self.residue_nf, self.n_dims = phar_nf, residue_nf, n_dims
        self.num_classes = phar_nf
        self.T = timesteps
        self.parametrization = parametrization
        self.norm_values, self.norm_biases = norm_values, norm_biases
        self.register_buffer('buffer', torch.zeros(1))
        self.size_distribution = DistributionNodes(size_histogram)
        self._check_norm_values()
        assert not self.dynamics.update_pocket_coords          # conditional_model.py:18
        self._tables = {}
        self.stepwise = False           # True: drive the loop from the host through the per-step API (reference control flow)
        self.noise_seed = None          # int: draw the sampler's noise with the library's counter-based device generator
        self.sample_ids = None          # [n_samples] global sample ids selecting the noise streams (default 0..n-1)

    def _check_norm_values(self, num_stdevs=8):               # en_diffusion.py:64-77
        g0 = self.gamma(torch.zeros((1, 1)))
        sigma_0 = torch.sqrt(torch.sigmoid(g0)).item()
        nv = self.norm_values[1]
        if sigma_0 * num_stdevs > 1. / nv:
            raise ValueError(f'Value for normalization value {nv} probably too large with sigma_0 '
                             f'{sigma_0:.5f} and 1 / norm_value = {1. / nv}')

    # ---------------------------------------------------------------- host glue
    def normalize(self, phar=None, pocket=None):              # en_diffusion.py:874-889 (mutates the dicts)
        for d in (phar, pocket):
            if d is not None:
                d['x'] = d['x'] / self.norm_values[0]
                d['one_hot'] = (d['one_hot'].float() - self.norm_biases[1]) / self.norm_values[1]
        return phar, pocket

    def unnormalize(self, x, h_cat):
        return x * self.norm_values[0], h_cat * self.norm_values[1] + self.norm_biases[1]

    def unnormalize_z(self, z_phar, z_pocket):
        nd = self.n_dims
        xp, hp = self.unnormalize(z_phar[:, :nd], z_phar[:, nd:])
        xk, hk = self.unnormalize(z_pocket[:, :nd], z_pocket[:, nd:])
        return torch.cat([xp, hp], dim=1), torch.cat([xk, hk], dim=1)

    @classmethod
    def remove_mean_batch(cls, x_phar, x_pocket, phar_indices, pocket_indices):
        mean = scatter_mean(x_phar, phar_indices)
        return x_phar - mean[phar_indices], x_pocket - mean[pocket_indices]

    @staticmethod
    def assert_mean_zero_with_mask(x, node_mask, eps=1e-10):
        largest = x.abs().max().item()
        error = scatter_add(x, node_mask).abs().max().item()
        rel = error / (largest + eps)
        assert rel < 1e-2, f'Mean is not zero, relative_error {rel}'

    @staticmethod
    def sample_gaussian(size, device):
        return torch.randn(size, device=device)

    def _draw(self, n_draws, size, device):
        """n_draws gaussian blocks.  If ``sample_gaussian`` was replaced on the instance
        (noise injection, as the parity harness does) it is called once per draw in order."""
        if 'sample_gaussian' in self.__dict__:
            return torch.stack([self.sample_gaussian(size, device) for _ in range(n_draws)]).to(torch.float32)
        return torch.randn((n_draws,) + tuple(size), device=device)

    def _table(self, timesteps):
        key = (timesteps, self.gamma.gamma.data_ptr(), self.gamma.gamma._version)
        if key not in self._tables:
            self._tables = {key: step_table(self.gamma.gamma, self.T, timesteps)}
        return self._tables[key]

    def _planned_handle(self, device, phar_mask, pocket_mask, n_samples):
        h = self.dynamics.handle(device)
        h.plan(self.dynamics._counts(phar_mask, n_samples), self.dynamics._counts(pocket_mask, n_samples))
        return h

    @staticmethod
    def _denoise(h, z, xh_pocket, t):
        """One denoiser evaluation of the per-step API with the reference's guards (dynamics.py:129-131): the NaN
        reset is reported, an edge-capacity overflow re-plans once and evaluates again."""
        for attempt in range(2):
            eps_hat, _ = h.dynamics_forward(z, xh_pocket, t, want_residues=False)
            fl = h.flags()
            if not fl.edge_overflow:
                break
            h.reset_flags()
            if attempt == 1:
                raise _lib.DiffPharError("edge buffer overflow persists after re-planning")
            h.grow_edge_capacity(int(fl.edge_overflow * 1.25) + 1024)
        if fl.nan_resets:
            print('Warning: detected nan, resetting EGNN output to zero.')
            h.reset_flags()
        return eps_hat

    @staticmethod
    def _scalar(v, what):
        v = v.reshape(-1)
        if v.numel() > 1 and not bool((v == v[0]).all()):
            raise NotImplementedError(f"{what} differs across the batch; the fused DDPM update takes one "
                                      "constant per step (all samples share s and t while sampling)")
        return float(v[0])

    # ---------------------------------------------------------------- per-step API
    def compute_x_pred(self, net_out, zt, gamma_t, batch_mask):      # en_diffusion.py:153-165
        sigma_t = self.sigma(gamma_t, target_tensor=net_out)
        alpha_t = self.alpha(gamma_t, target_tensor=net_out)
        return 1. / alpha_t[batch_mask] * (zt - sigma_t[batch_mask] * net_out)

    def sample_normal(self, *args):
        raise NotImplementedError("Has been replaced by sample_normal_zero_com()")

    def sample_combined_position_feature_noise(self, *args):
        raise NotImplementedError("Use sample_normal_zero_com() instead.")

    def sample(self, *args):
        raise NotImplementedError("Conditional model does not support sampling without given pocket.")

    # ---------------------------------------------------------------- loss terms (forward values only)
    # conditional_model.py:20-106, 158-320 and en_diffusion.py:167-179, 834-848, 908-944.  The denoiser evaluation —
    # the only expensive part — runs through the C-ABI (one dp_dynamics_forward with a per-sample t); everything
    # else is the reference's [B]-sized torch arithmetic.  VALUES ONLY: there is no backward through the CUDA kernels,
    # so this serves validation / test NLL (lightning_modules.py:264-287) and loss monitoring, not optimisation.
    def subspace_dimensionality(self, input_size):
        return (input_size - 1) * self.n_dims

    def delta_log_px(self, num_nodes):
        import numpy as np
        return -self.subspace_dimensionality(num_nodes) * np.log(self.norm_values[0])

    def log_pN(self, N_phar, N_pocket):
        return self.size_distribution.log_prob_n1_given_n2(N_phar, N_pocket)

    @staticmethod
    def sum_except_batch(x, indices):
        return scatter_add(x.sum(-1), indices)

    @staticmethod
    def cdf_standard_gaussian(x):
        import math
        return 0.5 * (1. + torch.erf(x / math.sqrt(2)))

    @staticmethod
    def gaussian_KL(q_mu_minus_p_mu_squared, q_sigma, p_sigma, d):
        return d * torch.log(p_sigma / q_sigma) + 0.5 * (d * q_sigma ** 2 + q_mu_minus_p_mu_squared) / (p_sigma ** 2) - 0.5 * d

    def sample_timesteps(self, lowest_t, n, device):
        """t ~ U{lowest_t..T} per sample (conditional_model.py:212-214); a hook like sample_gaussian for the parity harness."""
        return torch.randint(lowest_t, self.T + 1, size=(n, 1), device=device).float()

    def kl_prior(self, xh_phar, mask_phar, num_nodes):
        batch_size = len(num_nodes)
        ones = torch.ones((batch_size, 1), device=xh_phar.device)
        gamma_T = self.gamma(ones)
        alpha_T = self.alpha(gamma_T, xh_phar)
        mu_T = alpha_T[mask_phar] * xh_phar
        mu_T_x, mu_T_h = mu_T[:, :self.n_dims], mu_T[:, self.n_dims:]
        sigma_T_x = self.sigma(gamma_T, mu_T_x).squeeze()
        sigma_T_h = self.sigma(gamma_T, mu_T_h).squeeze()
        kl_h = self.gaussian_KL(self.sum_except_batch(mu_T_h ** 2, mask_phar), sigma_T_h, torch.ones_like(sigma_T_h), d=1)
        kl_x = self.gaussian_KL(self.sum_except_batch(mu_T_x ** 2, mask_phar), sigma_T_x, torch.ones_like(sigma_T_x),
                                self.subspace_dimensionality(num_nodes))
        return kl_x + kl_h

    def log_constants_p_x_given_z0(self, n_nodes, device):
        import numpy as np
        batch_size = len(n_nodes)
        gamma_0 = self.gamma(torch.zeros((batch_size, 1), device=device))
        log_sigma_x = 0.5 * gamma_0.view(batch_size)
        return self.subspace_dimensionality(n_nodes) * (-log_sigma_x - 0.5 * np.log(2 * np.pi))

    def log_pxh_given_z0_without_constants(self, phar, z_0_phar, eps_phar, net_out_phar, gamma_0, epsilon=1e-10):
        nd = self.n_dims
        z_h = z_0_phar[:, nd:]
        sigma_0_cat = self.sigma(gamma_0, target_tensor=z_0_phar) * self.norm_values[1]
        log_px = -0.5 * self.sum_except_batch((eps_phar[:, :nd] - net_out_phar[:, :nd]) ** 2, phar['mask'])
        onehot = phar['one_hot'] * self.norm_values[1] + self.norm_biases[1]
        centered = z_h * self.norm_values[1] + self.norm_biases[1] - 1
        log_ph_cat = torch.log(self.cdf_standard_gaussian((centered + 0.5) / sigma_0_cat[phar['mask']])
                               - self.cdf_standard_gaussian((centered - 0.5) / sigma_0_cat[phar['mask']]) + epsilon)
        log_prob = log_ph_cat - torch.logsumexp(log_ph_cat, dim=1, keepdim=True)
        return log_px, self.sum_except_batch(log_prob * onehot, phar['mask'])

    def noised_representation(self, xh_phar, xh0_pocket, phar_mask, pocket_mask, gamma_t):
        nd = self.n_dims
        alpha_t, sigma_t = self.alpha(gamma_t, xh_phar), self.sigma(gamma_t, xh_phar)
        eps = self.sample_gaussian(size=(len(phar_mask), nd + self.phar_nf), device=phar_mask.device)
        z_t = alpha_t[phar_mask] * xh_phar + sigma_t[phar_mask] * eps
        xh_pocket = xh0_pocket.detach().clone()
        z_t[:, :nd], xh_pocket[:, :nd] = self.remove_mean_batch(z_t[:, :nd], xh_pocket[:, :nd], phar_mask, pocket_mask)
        return z_t, xh_pocket, eps

    def xh_given_zt_and_epsilon(self, z_t, epsilon, gamma_t, batch_mask):
        alpha_t, sigma_t = self.alpha(gamma_t, z_t), self.sigma(gamma_t, z_t)
        return z_t / alpha_t[batch_mask] - epsilon * sigma_t[batch_mask] / alpha_t[batch_mask]

    @torch.no_grad()
    def forward(self, phar, pocket, return_info=False):
        """The loss / NLL terms of one batch (conditional_model.py:198-320), same tuple, forward values only."""
        nd = self.n_dims
        phar, pocket = self.normalize(phar, pocket)
        delta_log_px = self.delta_log_px(phar['size'])
        lowest_t = 0 if self.training else 1
        t_int = self.sample_timesteps(lowest_t, phar['size'].size(0), phar['x'].device)
        s_int = t_int - 1
        t_is_zero = (t_int == 0).float()
        t_is_not_zero = 1 - t_is_zero
        s, t = s_int / self.T, t_int / self.T
        gamma_s = self.inflate_batch_array(self.gamma(s), phar['x'])
        gamma_t = self.inflate_batch_array(self.gamma(t), phar['x'])
        xh0_phar = torch.cat([phar['x'], phar['one_hot']], dim=1)
        xh0_pocket = torch.cat([pocket['x'], pocket['one_hot']], dim=1)
        xh0_phar[:, :nd], xh0_pocket[:, :nd] = self.remove_mean_batch(xh0_phar[:, :nd], xh0_pocket[:, :nd],
                                                                     phar['mask'], pocket['mask'])
        z_t, xh_pocket, eps_t = self.noised_representation(xh0_phar, xh0_pocket, phar['mask'], pocket['mask'], gamma_t)
        net_out, _ = self.dynamics(z_t, xh_pocket, t, phar['mask'], pocket['mask'])
        xh_phar_hat = self.xh_given_zt_and_epsilon(z_t, net_out, gamma_t, phar['mask'])
        error_t = self.sum_except_batch((eps_t - net_out) ** 2, phar['mask'])
        SNR_weight = (1 - self.SNR(gamma_s - gamma_t)).squeeze(1)
        assert error_t.size() == SNR_weight.size()
        neg_log_constants = -self.log_constants_p_x_given_z0(n_nodes=phar['size'], device=error_t.device)
        kl_prior = self.kl_prior(xh0_phar, phar['mask'], phar['size'])
        if self.training:
            log_px, log_ph = self.log_pxh_given_z0_without_constants(phar, z_t, eps_t, net_out, gamma_t)
            loss_0_x = -log_px * t_is_zero.squeeze()
            loss_0_h = -log_ph * t_is_zero.squeeze()
            error_t = error_t * t_is_not_zero.squeeze()
        else:
            t_zeros = torch.zeros_like(s)
            gamma_0 = self.inflate_batch_array(self.gamma(t_zeros), phar['x'])
            z_0, xh_pocket, eps_0 = self.noised_representation(xh0_phar, xh0_pocket, phar['mask'], pocket['mask'], gamma_0)
            net_out_0, _ = self.dynamics(z_0, xh_pocket, t_zeros, phar['mask'], pocket['mask'])
            log_px, log_ph = self.log_pxh_given_z0_without_constants(phar, z_0, eps_0, net_out_0, gamma_0)
            loss_0_x, loss_0_h = -log_px, -log_ph
        log_pN = self.log_pN(phar['size'], pocket['size'])
        info = {'eps_hat_phar_x': scatter_mean(net_out[:, :nd].abs().mean(1), phar['mask']).mean(),
                'eps_hat_phar_h': scatter_mean(net_out[:, nd:].abs().mean(1), phar['mask']).mean()}
        loss_terms = (delta_log_px, error_t, torch.tensor(0.0), SNR_weight, loss_0_x, torch.tensor(0.0), loss_0_h,
                      neg_log_constants, kl_prior, log_pN, t_int.squeeze(), xh_phar_hat)
        return (*loss_terms, info) if return_info else loss_terms

    def sample_normal_zero_com(self, mu_phar, xh0_pocket, sigma, phar_mask, pocket_mask, fix_noise=False):
        if fix_noise:
            raise NotImplementedError("fix_noise option isn't implemented yet")
        dev = mu_phar.device
        n_samples = int(max(int(phar_mask.max()), int(pocket_mask.max()))) + 1
        h = self._planned_handle(dev, phar_mask, pocket_mask, n_samples)
        eps = self.sample_gaussian(size=(len(phar_mask), self.n_dims + self.phar_nf), device=phar_mask.device)
        z = mu_phar.detach().to(torch.float32).contiguous().clone()
        pocket = xh0_pocket.detach().to(torch.float32).contiguous().clone()
        sig = self._scalar(torch.as_tensor(sigma, dtype=torch.float32), "sigma")
        h.ddpm_update(2, 1.0, 0.0, sig, z, pocket, None, eps)
        return z, pocket

    def sample_p_zs_given_zt(self, s, t, zt_phar, xh0_pocket, phar_mask, pocket_mask, fix_noise=False):
        if fix_noise:
            raise NotImplementedError("fix_noise option isn't implemented yet")
        gamma_s, gamma_t = self.gamma(s), self.gamma(t)
        sigma2_ts, sigma_ts, alpha_ts = self.sigma_and_alpha_t_given_s(gamma_t, gamma_s, zt_phar)
        sigma_s = self.sigma(gamma_s, target_tensor=zt_phar)
        sigma_t = self.sigma(gamma_t, target_tensor=zt_phar)
        c_eps = sigma2_ts / alpha_ts / sigma_t
        sigma = sigma_ts * sigma_s / sigma_t
        n_samples = int(s.shape[0])
        h = self._planned_handle(zt_phar.device, phar_mask, pocket_mask, n_samples)
        eps_hat = self._denoise(h, zt_phar, xh0_pocket, t.to(torch.float32))
        eps = self.sample_gaussian(size=(len(phar_mask), self.n_dims + self.phar_nf), device=phar_mask.device)
        self.assert_mean_zero_with_mask(zt_phar[:, :self.n_dims], phar_mask)
        z = zt_phar.detach().to(torch.float32).contiguous().clone()
        pocket = xh0_pocket.detach().to(torch.float32).contiguous().clone()
        h.ddpm_update(0, self._scalar(alpha_ts, "alpha_t|s"), self._scalar(c_eps, "sigma2/alpha/sigma"),
                      self._scalar(sigma, "sigma"), z, pocket, eps_hat, eps)
        return z, pocket

    def sample_p_xh_given_z0(self, z0_phar, xh0_pocket, phar_mask, pocket_mask, batch_size, fix_noise=False):
        if fix_noise:
            raise NotImplementedError("fix_noise option isn't implemented yet")
        dev = z0_phar.device
        t_zeros = torch.zeros(size=(batch_size, 1), device=dev)
        gamma_0 = self.gamma(t_zeros)
        sigma_x = self.SNR(-0.5 * gamma_0)
        h = self._planned_handle(dev, phar_mask, pocket_mask, batch_size)
        eps_hat = self._denoise(h, z0_phar, xh0_pocket, t_zeros)
        net = eps_hat
        inv_alpha0 = 1. / self.alpha(gamma_0, target_tensor=net)
        sigma0 = self.sigma(gamma_0, target_tensor=net)
        eps = self.sample_gaussian(size=(len(phar_mask), self.n_dims + self.phar_nf), device=phar_mask.device)
        z = z0_phar.detach().to(torch.float32).contiguous().clone()
        pocket = xh0_pocket.detach().to(torch.float32).contiguous().clone()
        h.ddpm_update(1, self._scalar(inv_alpha0, "1/alpha_0"), self._scalar(sigma0, "sigma_0"),
                      self._scalar(sigma_x, "sigma_x"), z, pocket, eps_hat, eps)
        nd = self.n_dims
        x_phar, h_phar = self.unnormalize(z[:, :nd], z0_phar[:, nd:])
        x_pocket, h_pocket = self.unnormalize(pocket[:, :nd], pocket[:, nd:])
        h_phar = F.one_hot(torch.argmax(h_phar, dim=1), self.phar_nf)
        return x_phar, h_phar, x_pocket, h_pocket

    # ---------------------------------------------------------------- the sampler
    @torch.no_grad()
    def sample_given_pocket(self, pocket, num_nodes_phar, return_frames=1, timesteps=None):
        timesteps = self.T if timesteps is None else timesteps
        assert 0 < return_frames <= timesteps
        assert timesteps % return_frames == 0
        n_samples = len(pocket['size'])
        device = pocket['x'].device
        nd = self.n_dims

        _, pocket = self.normalize(pocket=pocket)
        xh0_pocket = torch.cat([pocket['x'], pocket['one_hot']], dim=1)
        phar_mask = num_nodes_to_batch_mask(n_samples, num_nodes_phar, device)
        if self.stepwise:
            return self._sample_with_frames(pocket, xh0_pocket, phar_mask, n_samples, return_frames, timesteps)

        h = self._planned_handle(device, phar_mask, pocket['mask'], n_samples)
        tab = self._table(timesteps)
        h.set_step_table(tab.rows, tab.final)
        # gaussian draws: torch's generator as in the reference (sample_gaussian, en_diffusion.py:946-949; also the hook
        # the parity harness replaces), or — with `noise_seed` set — the library's counter-based device generator,
        # whose draws depend on (seed, global sample id) only: the same sample comes out whatever the batching / sharding
        noise = None if self.noise_seed is not None and 'sample_gaussian' not in self.__dict__ else \
            self._draw(timesteps + 2, (len(phar_mask), nd + self.phar_nf), device)
        norm = (float(self.norm_values[0]), float(self.norm_values[1]), float(self.norm_biases[1]))
        grown = False
        while True:
            xh_pocket = xh0_pocket.detach().to(torch.float32).contiguous().clone()
            res = h.sample(xh_pocket, noise, seed=self.noise_seed or 0, sample_ids=self.sample_ids,
                           return_frames=return_frames, norm=norm)
            fl = h.flags()                                    # ONE host read for the whole run
            if fl.edge_overflow:
                # the radius graph outgrew the planned edge capacity (the device truncated it, memory-safe): re-plan
                # with the edge count it reported and run again on the same noise
                h.reset_flags()
                if grown:
                    raise _lib.DiffPharError(f"edge buffer overflow persists at capacity {cap}")
                cap = int(fl.edge_overflow * 1.25) + 1024
                h.grow_edge_capacity(cap)
                grown = True
                continue
            if fl.f16_range and self.dynamics.precision not in ("fp32", "tf32"):
                # features or squared distances left f16's range (no cutoff and far-apart points, or very large
                # activations): the 16-bit result would differ from the reference — repeat with fp32 storage
                print(f'Warning: f16 range exceeded (flag {fl.f16_range}); re-running with tf32 tensor-core tiles.')
                h.reset_flags()
                self.dynamics.set_precision("tf32")
                continue
            break
        out, frames_phar, frames_pocket = res if return_frames > 1 else (res, None, None)
        if fl.nan_resets:
            print('Warning: detected nan, resetting EGNN output to zero.')
        assert fl.max_mean_rel_err < 1e-2, f'Mean is not zero, relative_error {fl.max_mean_rel_err}'
        h.reset_flags()

        x_phar, h_phar = self.unnormalize(out[:, :nd], out[:, nd:])
        x_pocket, h_pocket = self.unnormalize(xh_pocket[:, :nd], xh_pocket[:, nd:])
        h_phar = F.one_hot(torch.argmax(h_phar, dim=1), self.phar_nf)
        self.assert_mean_zero_with_mask(x_phar, phar_mask)
        if return_frames == 1:                                # conditional_model.py:449-457
            max_cog = scatter_add(x_phar, phar_mask).abs().max().item()
            if max_cog > 5e-2:
                print(f'Warning CoG drift with error {max_cog:.3f}. Projecting the positions down.')
                x_phar, x_pocket = self.remove_mean_batch(x_phar, x_pocket, phar_mask, pocket['mask'])
        out_phar = torch.cat([x_phar, h_phar.to(x_phar.dtype)], dim=1)
        out_pocket = torch.cat([x_pocket, h_pocket], dim=1)
        if return_frames > 1:                                 # frames were written inside the captured loop; frame 0 = the result
            frames_phar[0], frames_pocket[0] = out_phar, out_pocket
            return frames_phar, frames_pocket, phar_mask, pocket['mask']
        return out_phar, out_pocket, phar_mask, pocket['mask']

    def _sample_with_frames(self, pocket, xh0_pocket, phar_mask, n_samples, return_frames, timesteps):
        """The reference's loop driven from the host through the per-step API, one fused step at a time (kept for
        callers that step the sampler themselves, and as the checker of the in-graph frame output)."""
        device = xh0_pocket.device
        nd = self.n_dims
        mu_x = scatter_mean(pocket['x'], pocket['mask'])
        mu_h = torch.zeros((n_samples, self.phar_nf), device=device)
        mu = torch.cat((mu_x, mu_h), dim=1)[phar_mask]
        sigma = torch.ones(1, device=device)
        z, xh_pocket = self.sample_normal_zero_com(mu, xh0_pocket, sigma, phar_mask, pocket['mask'])
        self.assert_mean_zero_with_mask(z[:, :nd], phar_mask)
        out_phar = torch.zeros((return_frames,) + z.size(), device=device)
        out_pocket = torch.zeros((return_frames,) + xh_pocket.size(), device=device)
        for s in reversed(range(0, timesteps)):
            s_array = torch.full((n_samples, 1), fill_value=s, device=device)
            t_array = (s_array + 1) / timesteps
            s_array = s_array / timesteps
            z, xh_pocket = self.sample_p_zs_given_zt(s_array, t_array, z, xh_pocket, phar_mask, pocket['mask'])
            if (s * return_frames) % timesteps == 0:
                idx = (s * return_frames) // timesteps
                out_phar[idx], out_pocket[idx] = self.unnormalize_z(z, xh_pocket)
        x_phar, h_phar, x_pocket, h_pocket = self.sample_p_xh_given_z0(z, xh_pocket, phar_mask, pocket['mask'], n_samples)
        self.assert_mean_zero_with_mask(x_phar, phar_mask)
        out_phar[0] = torch.cat([x_phar, h_phar], dim=1)
        out_pocket[0] = torch.cat([x_pocket, h_pocket], dim=1)
        return out_phar.squeeze(0), out_pocket.squeeze(0), phar_mask, pocket['mask']
