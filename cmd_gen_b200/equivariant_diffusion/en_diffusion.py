"""Host-side pieces of DiffPhar/equivariant_diffusion/en_diffusion.py that the sampler
path needs: the predefined noise schedule module (state-dict key ``gamma.gamma``), the
node-count prior, and the small [B,1] schedule helpers callers may still use.

Out of scope (not reachable from generate_phars): the learned
GammaNetwork, the joint-mode ``sample``/``inpaint``, the training losses.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..schedule import gamma_table


class PredefinedNoiseSchedule(torch.nn.Module):
    """Lookup table gamma[T+1] (en_diffusion.py:1152-1188)."""

    def __init__(self, noise_schedule, timesteps, precision):
        super().__init__()
        self.timesteps = timesteps
        self.gamma = torch.nn.Parameter(gamma_table(noise_schedule, timesteps, precision), requires_grad=False)

    def forward(self, t):
        t_int = torch.round(t * self.timesteps).long()
        return self.gamma[t_int]


class DistributionNodes:
    """Joint histogram over (n_phar, n_pocket) sizes (en_diffusion.py:952-1030)."""

    def __init__(self, histogram):
        histogram = torch.as_tensor(histogram).float() + 1e-3
        prob = histogram / histogram.sum()
        n1, n2 = prob.shape
        self.idx_to_n_nodes = torch.stack(torch.meshgrid(torch.arange(n1), torch.arange(n2), indexing="ij"),
                                          dim=-1).view(-1, 2)
        self.n_nodes_to_idx = {tuple(x.tolist()): i for i, x in enumerate(self.idx_to_n_nodes)}
        self.prob = prob
        self.m = torch.distributions.Categorical(self.prob.view(-1), validate_args=True)
        self.n1_given_n2 = [torch.distributions.Categorical(prob[:, j], validate_args=True) for j in range(n2)]
        self.n2_given_n1 = [torch.distributions.Categorical(prob[i, :], validate_args=True) for i in range(n1)]
        print("Entropy of n_nodes: H[N]", self.m.entropy().item())

    def sample(self, n_samples=1):
        idx = self.m.sample((n_samples,))
        a, b = self.idx_to_n_nodes[idx].T
        return a, b

    def sample_conditional(self, n1=None, n2=None):
        assert (n1 is None) ^ (n2 is None), "Exactly one input argument must be None"
        m = self.n1_given_n2 if n2 is not None else self.n2_given_n1
        c = n2 if n2 is not None else n1
        return torch.tensor([m[i].sample() for i in c], device=c.device)

    def log_prob(self, batch_n_nodes_1, batch_n_nodes_2):
        assert batch_n_nodes_1.dim() == 1 and batch_n_nodes_2.dim() == 1
        idx = torch.tensor([self.n_nodes_to_idx[(a, b)]
                            for a, b in zip(batch_n_nodes_1.tolist(), batch_n_nodes_2.tolist())])
        return self.m.log_prob(idx).to(batch_n_nodes_1.device)

    def log_prob_n1_given_n2(self, n1, n2):
        assert n1.dim() == 1 and n2.dim() == 1
        lp = torch.stack([self.n1_given_n2[c].log_prob(i.cpu()) for i, c in zip(n1, n2)])
        return lp.to(n1.device)

    def log_prob_n2_given_n1(self, n2, n1):
        assert n1.dim() == 1 and n2.dim() == 1
        lp = torch.stack([self.n2_given_n1[c].log_prob(i.cpu()) for i, c in zip(n2, n1)])
        return lp.to(n2.device)


class ScheduleMixin:
    """sigma/alpha/SNR helpers with the reference's op order (en_diffusion.py:79-103, 849-872)."""

    @staticmethod
    def inflate_batch_array(array, target):
        return array.view((array.size(0),) + (1,) * (len(target.size()) - 1))

    def sigma(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(gamma)), target_tensor)

    def alpha(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(-gamma)), target_tensor)

    @staticmethod
    def SNR(gamma):
        return torch.exp(-gamma)

    def sigma_and_alpha_t_given_s(self, gamma_t, gamma_s, target_tensor):
        sigma2 = self.inflate_batch_array(-torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t)), target_tensor)
        log_a2 = F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)
        alpha = self.inflate_batch_array(torch.exp(0.5 * log_a2), target_tensor)
        return sigma2, torch.sqrt(sigma2), alpha
