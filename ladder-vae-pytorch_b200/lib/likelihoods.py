"""Kernel-backed mirror of the reference's lib/likelihoods.py.

``LikelihoodModule.forward(input_, x) -> (ll, {mean, mode, sample, params})`` as in
lib/likelihoods.py:33-48.  The Bernoulli and 10-component discretized-logistic-mixture heads (the
two used by every BASELINE config) run fused forward/backward kernels; the Gaussian and
single-logistic heads are outside the accelerated path ("next") and evaluate their
density with ordinary tensor ops on the GPU after our conv.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from lvae_b200 import ops
from .nn import Conv2d
from .stochastic import logistic_rsample


class LikelihoodModule(nn.Module):

    def distr_params(self, x):
        raise NotImplementedError

    @staticmethod
    def mean(params):
        return None

    @staticmethod
    def mode(params):
        return None

    @staticmethod
    def sample(params):
        return None

    def log_likelihood(self, x, params):
        raise NotImplementedError

    def forward(self, input_, x):
        params = self.distr_params(input_)
        info = {"mean": self.mean(params), "mode": self.mode(params), "sample": self.sample(params), "params": params}
        ll = None if x is None else self.log_likelihood(x, params)
        return ll, info


class BernoulliLikelihood(LikelihoodModule):
    """sigmoid(conv) probabilities; ll = -BCE summed per image, logs clamped at -100 (likelihoods.py:51-78,385-388)."""

    def __init__(self, ch_in, color_channels):
        super().__init__()
        self.parameter_net = Conv2d(ch_in, color_channels, kernel_size=3, padding=1)

    def forward(self, input_, x):
        logits = self.parameter_net(input_)
        if logits.dtype != torch.float32:
            logits = logits.float()
        prob, ll = ops.bernoulli_loglik(logits, x)      # one kernel: sigmoid + log-likelihood
        with torch.no_grad():
            info = {"mean": prob, "mode": torch.round(prob), "sample": ops.bernoulli_sample(prob), "params": prob}
        return ll, info

    def distr_params(self, x):
        return ops.bernoulli_loglik(self.parameter_net(x).float(), None)[0]

    @staticmethod
    def mean(params):
        return params

    @staticmethod
    def mode(params):
        return torch.round(params)

    @staticmethod
    def sample(params):
        return ops.bernoulli_sample(params)

    def log_likelihood(self, x, params):
        return log_bernoulli(x, params, reduce="none")


def log_bernoulli(x, mean, reduce="mean"):
    """x log p + (1-x) log(1-p) with BCE's -100 clamp, summed over (c,h,w) (likelihoods.py:385-388).
    Takes *probabilities*; used when the caller did not go through the fused module forward."""
    lp = torch.clamp(torch.log(mean), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - mean), min=-100.0)
    return _reduce((x * lp + (1.0 - x) * l1p).sum((1, 2, 3)), reduce)


class DiscretizedLogisticMixLikelihood(LikelihoodModule):
    """PixelCNN++ mixture of 10 discretized logistics on RGB images scaled to [0,1]
    (likelihoods.py:183-230).  mean / mode are None like the reference."""

    def __init__(self, ch_in, n_components=10):
        super().__init__()
        if n_components != 10:
            raise NotImplementedError("the fused kernel is specialised for 10 mixture components")
        self.parameter_net = Conv2d(ch_in, 10 * n_components, kernel_size=3, padding=1)

    def forward(self, input_, x):
        # bf16 tensor-core path: conv + log-likelihood as one node (its backward feeds the tcgen05 dgrad / wgrad directly)
        if x is not None and not self.parameter_net._forward_hooks:
            fused = ops.dmol_head(input_, self.parameter_net, x)
            if fused is not None:
                ll, l = fused
                params = {"mean": None, "all_params": l}
                return ll, {"mean": None, "mode": None, "sample": self.sample(params), "params": params}
        return super().forward(input_, x)

    def distr_params(self, x):
        l = self.parameter_net(x)
        return {"mean": None, "all_params": l if l.dtype == torch.float32 else l.float()}

    @staticmethod
    def mean(params):
        return params["mean"]

    @staticmethod
    def mode(params):
        return params["mean"]

    @staticmethod
    def sample(params):
        with torch.no_grad():
            return ops.dmol_sample(params["all_params"])     # already rescaled to [0,1] and clamped

    def log_likelihood(self, x, params):
        return ops.dmol_loglik(params["all_params"], x)


def discretized_mix_logistic_loss(x, l):
    """Negative log-likelihood per image for x in [-1,1] (likelihoods.py:291-382), fused kernel."""
    return -ops.dmol_loglik(l, (x + 1) / 2)


# ----------------------------------------------------------------------------- not on the accelerated path
class GaussianLikelihood(LikelihoodModule):
    """Per-pixel Gaussian head (likelihoods.py:81-114): our conv, tensor-op density."""

    def __init__(self, ch_in, color_channels):
        super().__init__()
        self.parameter_net = Conv2d(ch_in, 2 * color_channels, kernel_size=3, padding=1)

    def distr_params(self, x):
        mean, lv = self.parameter_net(x).float().chunk(2, dim=1)
        return {"mean": mean, "logvar": lv}

    @staticmethod
    def mean(params):
        return params["mean"]

    @staticmethod
    def mode(params):
        return params["mean"]

    @staticmethod
    def sample(params):
        return params["mean"] + (params["logvar"] / 2).exp() * torch.randn_like(params["mean"])

    def log_likelihood(self, x, params):
        return log_normal(x, params["mean"], params["logvar"], reduce="none")


class DiscretizedLogisticLikelihood(LikelihoodModule):
    """Single discretized logistic per sub-pixel (likelihoods.py:117-180): our conv, tensor-op density."""

    log_scale_bias = -1.0

    def __init__(self, ch_in, color_channels, n_bins, double=False):
        super().__init__()
        self.n_bins, self.double_precision = n_bins, double
        self.parameter_net = Conv2d(ch_in, 2 * color_channels, kernel_size=3, padding=1)

    def distr_params(self, x):
        mean, ls = self.parameter_net(x).float().chunk(2, dim=1)
        return {"mean": mean + 0.5, "logscale": (ls + self.log_scale_bias).clamp(min=-7.0)}

    @staticmethod
    def mean(params):
        return params["mean"]

    @staticmethod
    def mode(params):
        return params["mean"]

    @staticmethod
    def sample(params):
        return logistic_rsample((params["mean"], params["logscale"])).clamp(min=0.0, max=1.0)

    def log_likelihood(self, x, params):
        x = x * (255 / 256) + 1 / 512
        return log_discretized_logistic(x, params["mean"], params["logscale"], n_bins=self.n_bins, reduce="none",
                                        double=self.double_precision)


def log_discretized_logistic(x, mean, log_scale, n_bins=256, reduce="mean", double=False):
    """Log mass of the bin containing x under Logistic(mean, exp(log_scale)) (likelihoods.py:233-288)."""
    log_scale = _input_check(x, mean, log_scale, reduce)
    eps = 1e-7
    if double:
        log_scale, x, mean, eps = log_scale.double(), x.double(), mean.double(), 1e-14
    scale = log_scale.exp().expand_as(x)
    lo = torch.floor(x * n_bins) / n_bins
    upper = torch.where(lo < (n_bins - 1) / n_bins, torch.sigmoid((lo + 1 / n_bins - mean) / scale), torch.ones_like(lo))
    lower = torch.where(lo >= 1 / n_bins, torch.sigmoid((lo - mean) / scale), torch.zeros_like(lo))
    out = _reduce(torch.log(upper - lower + eps).sum((1, 2, 3)), reduce)
    return out.float() if double else out


def log_normal(x, mean, logvar, reduce="mean"):
    """Diagonal Gaussian log-density summed over (c,h,w) (likelihoods.py:391-411)."""
    logvar = _input_check(x, mean, logvar, reduce)
    lp = -0.5 * ((x - mean) ** 2 / logvar.exp() + logvar + _LOG_2PI)
    return _reduce(lp.sum((1, 2, 3)), reduce)


# The reference adds log(2 pi) as a float32 0-dim tensor (likelihoods.py:408), so float64 inputs also see the
# float32-rounded constant (1.8378770351...); a Python double here would differ from it by 3e-8 per element.
_LOG_2PI = float(torch.tensor(2 * math.pi, dtype=torch.float32).log())


def _reduce(x, reduce):
    if reduce == "mean":
        return x.mean()
    if reduce == "sum":
        return x.sum()
    return x


def _input_check(x, mean, scale_param, reduce):
    assert x.dim() == 4
    assert x.size() == mean.size()
    if scale_param.numel() == 1:
        scale_param = scale_param.view(1, 1, 1, 1)
    if reduce not in ["mean", "sum", "none"]:
        raise RuntimeError("unrecognized reduction method '{}'".format(reduce))
    return scale_param
