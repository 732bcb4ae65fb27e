"""Kernel-backed mirror of the reference's lib/stochastic.py.

``NormalStochasticBlock2d`` keeps its three conv submodules (``conv_in_p``, ``conv_in_q``,
``conv_out`` -> same state_dict keys, lib/stochastic.py:24-27) and its forward signature / output
dict (lib/stochastic.py:29-112); everything between the convs is ONE fused kernel
(lvae_stoch_fwd / lvae_stoch_bwd).
"""
from __future__ import annotations

import torch
from torch import nn

from lvae_b200 import ops
from .nn import Conv2d


class NormalStochasticBlock2d(nn.Module):
    """Project to (mu, logvar) of q (and p), sample z, return conv_out(z) and the KL terms."""

    def __init__(self, c_in, c_vars, c_out, kernel=3, transform_p_params=True):
        super().__init__()
        assert kernel % 2 == 1
        pad = kernel // 2
        self.transform_p_params = transform_p_params
        self.c_in, self.c_out, self.c_vars = c_in, c_out, c_vars
        if transform_p_params:
            self.conv_in_p = Conv2d(c_in, 2 * c_vars, kernel, padding=pad)
        self.conv_in_q = Conv2d(c_in, 2 * c_vars, kernel, padding=pad)
        self.conv_out = Conv2d(c_vars, c_out, kernel, padding=pad)

    def forward(self, p_params, q_params=None, forced_latent=None, use_mode=False, force_constant_output=False,
                analytical_kl=False):
        assert (forced_latent is None) or (not use_mode)
        if self.transform_p_params:
            p_params = self.conv_in_p(p_params)
        else:
            assert p_params.size(1) == 2 * self.c_vars
        if q_params is not None:
            q_params = self.conv_in_q(q_params)
        eps = None
        if forced_latent is None and not use_mode:
            eps = ops.pop_eps()      # only set by parity tests (ops.inject)
        lowp = self.conv_out.weight.is_cuda and getattr(self, "compute_dtype", torch.float32) == torch.bfloat16
        z, z_lp, kl_samplewise, kl_spatial, logprob_p, logprob_q = ops.stochastic_core(
            q_params, p_params, eps=eps, forced=forced_latent, use_mode=use_mode, analytical=analytical_kl,
            lowp_copy=lowp)
        if force_constant_output:
            # prior experiment (lib/stochastic.py:71-73): one sample shared by the whole batch;
            # log p(z) is still evaluated under each row's own p
            z = z[0:1].expand_as(z).contiguous()
            p_shared = p_params[0:1].expand_as(p_params).contiguous()
            z, z_lp, _, _, logprob_p, _ = ops.stochastic_core(None, p_params, forced=z, lowp_copy=lowp)
            p_params = p_shared
        out = self.conv_out(z, x_lowp=z_lp)
        data = {
            "z": z,
            "p_params": p_params,
            "q_params": q_params,
            "logprob_p": logprob_p,
            "logprob_q": logprob_q,
            # the element-wise KL map is never materialised: kl_samplewise is its fused sum
            "kl_elementwise": None,
            "kl_samplewise": kl_samplewise,
            "kl_spatial": kl_spatial,
        }
        return out, data


def logistic_rsample(mu_ls):
    """Reparameterised logistic sample (lib/stochastic.py:115-138).  Only used by the
    single-logistic likelihood, which is outside the accelerated path: plain tensor ops."""
    try:
        mu, log_scale = torch.chunk(mu_ls, 2, dim=1)
    except TypeError:
        mu, log_scale = mu_ls
    u = torch.empty_like(mu).uniform_(1e-7, 1 - 1e-7)
    # log(1 - u), not log1p(-u): bit-for-bit what the reference draws from the same generator state (:135)
    return mu + log_scale.exp() * (torch.log(u) - torch.log(1 - u))


def sample_from_discretized_mix_logistic(l):
    """Sample in [-1, 1] from the 10-component mixture (lib/stochastic.py:141-206), fused kernel."""
    return ops.dmol_sample(l) * 2 - 1
