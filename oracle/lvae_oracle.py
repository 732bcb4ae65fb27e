"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Ladder-VAE hot path.

This is the *oracle* for the B200 kernels: a functional (state_dict in, dict
out) restatement, in plain PyTorch CPU ops, of what the reference computes on
the path BASELINE.json names (ELBO training step + importance-weighted bound).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product package never does; it
fails loudly when its CUDA library is missing.

Every function cites the reference file:line (paths under /root/reference)
it follows.  The restatement is pinned two ways (tests/test_oracle.py):
  * live, in the build container, against the unmodified reference imported
    through ``oracle/ref_loader.py`` (skipped where /root/reference is absent);
  * everywhere, against ``tests/golden/*.npz`` written by ``oracle/make_golden.py``
    from the unmodified reference.
PARITY UNPINNED for four helpers whose source (boilr==0.7.4, requirements.txt:6)
is not in the reference tree: free_bits_kl, pad/crop_img_tensor, Interpolate.
They are restated from documented behaviour and pinned only
against the stand-ins in ref_loader.py.

Tensors are logical NCHW fp32/fp64 like the reference.  Parameters are a flat
``dict`` keyed exactly like ``LadderVAE.state_dict()``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field, asdict
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class LVAEConfig:
    """Constructor arguments of LadderVAE (models/lvae.py:17-35), same names."""
    color_ch: int
    z_dims: Sequence[int]
    img_shape: Sequence[int]
    blocks_per_layer: int = 2
    downsample: Optional[Sequence[int]] = None
    nonlin: str = "elu"
    merge_type: Optional[str] = None
    batchnorm: bool = True
    stochastic_skip: bool = False
    n_filters: int = 32
    dropout: Optional[float] = None
    free_bits: float = 0.0
    learn_top_prior: bool = False
    likelihood_form: Optional[str] = None
    res_block_type: Optional[str] = None
    gated: bool = False
    no_initial_downscaling: bool = False
    analytical_kl: bool = False

    def __post_init__(self):
        self.z_dims = list(self.z_dims)
        self.img_shape = tuple(self.img_shape)
        if self.downsample is None:                       # lvae.py:52-53
            self.downsample = [0] * len(self.z_dims)
        self.downsample = list(self.downsample)
        assert len(self.downsample) == len(self.z_dims)   # lvae.py:61
        assert max(self.downsample) <= self.blocks_per_layer  # lvae.py:60

    @property
    def n_layers(self) -> int:
        return len(self.z_dims)

    @property
    def overall_downscale_factor(self) -> int:            # lvae.py:56-58
        f = 2 ** sum(self.downsample)
        return f if self.no_initial_downscaling else 2 * f

    def padded_size(self, hw) -> List[int]:               # lvae.py:327-349
        d = self.overall_downscale_factor
        return [((int(s) - 1) // d + 1) * d for s in hw]

    def top_prior_shape(self, n=1):                       # lvae.py:364-372
        d = self.overall_downscale_factor
        ph, pw = self.padded_size(self.img_shape)
        return (n, 2 * self.z_dims[-1], ph // d, pw // d)

    def kwargs(self) -> dict:
        return asdict(self)


# The five BASELINE.json configs (SURVEY.md 8, table at the top).
def baseline_config(name: str) -> LVAEConfig:
    common = dict(blocks_per_layer=4, n_filters=64, nonlin="elu", gated=True,
                  stochastic_skip=True, merge_type="residual", res_block_type="bacdbacd",
                  dropout=0.2, batchnorm=True, learn_top_prior=True, analytical_kl=False)
    if name == "mnist3":
        return LVAEConfig(color_ch=1, z_dims=[32] * 3, img_shape=(28, 28), downsample=[1, 1, 1],
                          free_bits=0.5, likelihood_form="bernoulli", **common)
    if name == "mnist12":
        return LVAEConfig(color_ch=1, z_dims=[32] * 12, img_shape=(28, 28),
                          downsample=[0, 0, 0, 1] * 3, free_bits=1.0,
                          likelihood_form="bernoulli", **common)
    if name == "cifar15":
        return LVAEConfig(color_ch=3, z_dims=[32] * 15, img_shape=(32, 32),
                          downsample=[0, 0, 0, 0, 1] * 3, free_bits=1.0,
                          likelihood_form="discr_log_mix", **common)
    if name == "celeba20":
        return LVAEConfig(color_ch=3, z_dims=[32] * 20, img_shape=(64, 64),
                          downsample=[0, 0, 0, 0, 1] * 4, free_bits=1.0,
                          likelihood_form="discr_log_mix", **common)
    raise KeyError(name)


# --------------------------------------------------------------------------- state_dict layout
def _res_block_entries(cfg: LVAEConfig, gated: bool):
    """Sequential layout of ResidualBlock.block (lib/nn.py:48-96): list of
    (index, kind) with kind in {'bn','act','conv','drop','gate'}."""
    t, out = cfg.res_block_type, []
    if t == "cabdcabd":                                   # nn.py:50-62
        for _ in range(2):
            out += ["conv", "act"]
            if cfg.batchnorm:
                out.append("bn")
            if cfg.dropout is not None:
                out.append("drop")
    elif t == "bacdbac":                                  # nn.py:64-76
        for i in range(2):
            if cfg.batchnorm:
                out.append("bn")
            out += ["act", "conv"]
            if cfg.dropout is not None and i == 0:
                out.append("drop")
    elif t == "bacdbacd":                                 # nn.py:78-89
        if cfg.dropout is None:
            raise TypeError("bacdbacd builds nn.Dropout2d(None) (lib/nn.py:89)")
        for _ in range(2):
            if cfg.batchnorm:
                out.append("bn")
            out += ["act", "conv", "drop"]
    else:
        raise ValueError("unrecognized block type '{}'".format(t))  # nn.py:92
    if gated:
        out.append("gate")                                # nn.py:94-95
    return list(enumerate(out))


def param_shapes(cfg: LVAEConfig) -> "OrderedDict[str, tuple]":
    """Names, order and shapes of ``LadderVAE(...).state_dict()``."""
    C, sd = cfg.n_filters, OrderedDict()

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = (cout, cin, k, k)
        sd[name + ".bias"] = (cout,)

    def bn(name):
        sd[name + ".weight"] = (C,)
        sd[name + ".bias"] = (C,)
        sd[name + ".running_mean"] = (C,)
        sd[name + ".running_var"] = (C,)
        sd[name + ".num_batches_tracked"] = ()

    def res_block(prefix, gated):                         # lib/nn.py:5-99
        for idx, kind in _res_block_entries(cfg, gated):
            if kind == "conv":
                conv("%s.block.%d" % (prefix, idx), C, C, 3)
            elif kind == "bn":
                bn("%s.block.%d" % (prefix, idx))
            elif kind == "gate":
                conv("%s.block.%d.conv" % (prefix, idx), 2 * C, C, 1)   # nn.py:118

    def resampling_block(prefix, resample, gated, transposed):  # lvae_layers.py:222-306
        if resample:
            conv(prefix + ".pre_conv", C, C, 3)           # ConvTranspose is (Cin,Cout,3,3): same here
        res_block(prefix + ".res", gated)

    def merge(prefix):                                    # lvae_layers.py:323-360
        if cfg.merge_type == "linear":
            conv(prefix + ".layer", C, 2 * C, 1)
        elif cfg.merge_type == "residual":
            conv(prefix + ".layer.0", C, 2 * C, 1)
            res_block(prefix + ".layer.1", True)

    L = cfg.n_layers
    conv("first_bottom_up.0", C, cfg.color_ch, 5)         # lvae.py:74-75
    resampling_block("first_bottom_up.2", False, False, False)  # lvae.py:77-84 (no gated= passed)
    for i in range(L):                                    # lvae.py:90 onwards (top_down_layers first)
        p = "top_down_layers.%d" % i
        top = i == L - 1
        if top:
            sd[p + ".top_prior_params"] = cfg.top_prior_shape()   # lvae_layers.py:55-58
        left = cfg.downsample[i]
        for b in range(cfg.blocks_per_layer):             # lvae_layers.py:64-82
            resampling_block("%s.deterministic_block.%d" % (p, b), left > 0, cfg.gated, True)
            left -= 1 if left > 0 else 0
        if not top:                                       # stochastic.py:24-27
            conv(p + ".stochastic.conv_in_p", 2 * cfg.z_dims[i], C, 3)
        conv(p + ".stochastic.conv_in_q", 2 * cfg.z_dims[i], C, 3)
        conv(p + ".stochastic.conv_out", C, cfg.z_dims[i], 3)
        if not top:
            merge(p + ".merge")                           # lvae_layers.py:94-103
            if cfg.stochastic_skip:                       # lvae_layers.py:106-113 (always 'residual')
                conv(p + ".skip_connection_merger.layer.0", C, 2 * C, 1)
                res_block(p + ".skip_connection_merger.layer.1", True)
    for i in range(L):
        left = cfg.downsample[i]
        for b in range(cfg.blocks_per_layer):             # lvae_layers.py:201-215
            resampling_block("bottom_up_layers.%d.net.%d" % (i, b), left > 0, cfg.gated, False)
            left -= 1 if left > 0 else 0
    off = 0 if cfg.no_initial_downscaling else 1          # lvae.py:141-156
    for b in range(cfg.blocks_per_layer):
        resampling_block("final_top_down.%d" % (b + off), False, cfg.gated, True)
    nout = {"bernoulli": cfg.color_ch, "discr_log_mix": 100}.get(cfg.likelihood_form)
    if nout is None:
        raise RuntimeError("Unrecognized likelihood '{}'".format(cfg.likelihood_form))  # lvae.py:169
    conv("likelihood.parameter_net", nout, C, 3)          # likelihoods.py:55,199
    # state_dict order: a module's own parameters come before its children; ModuleLists in
    # attribute-registration order (first_bottom_up, top_down_layers, bottom_up_layers, ...).
    return sd


def make_params(cfg: LVAEConfig, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, platform-independent synthetic weights (numpy RandomState):
    conv weights N(0, 1/fan_in), small biases, non-trivial BatchNorm affine and
    running statistics, non-zero top prior.  Used for every golden fixture."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            t = torch.from_numpy(0.1 * rng.standard_normal(shape)).to(dtype)
        elif leaf == "running_var":
            t = torch.from_numpy(1.0 + 0.2 * np.abs(rng.standard_normal(shape))).to(dtype)
        elif leaf == "top_prior_params":
            t = torch.from_numpy(0.2 * rng.standard_normal(shape)).to(dtype)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            # residual-branch convs are damped so that a deep eval-mode stack (BatchNorm on
            # these synthetic running statistics) keeps O(1) activations
            gain = 0.35 if ".block." in name else (0.5 if ".conv_in_" in name else 1.0)
            t = torch.from_numpy(gain * rng.standard_normal(shape) / math.sqrt(fan_in)).to(dtype)
        elif leaf == "weight":   # BatchNorm gamma
            t = torch.from_numpy(1.0 + 0.1 * rng.standard_normal(shape)).to(dtype)
        else:                    # biases, BatchNorm beta
            t = torch.from_numpy(0.05 * rng.standard_normal(shape)).to(dtype)
        out[name] = t
    return out


def latent_shapes(cfg: LVAEConfig, batch: int) -> List[tuple]:
    """Shapes of z_i bottom to top (what eps must look like)."""
    ph, pw = cfg.padded_size(cfg.img_shape)
    f = 1 if cfg.no_initial_downscaling else 2
    shapes = []
    for i in range(cfg.n_layers):
        f *= 2 ** cfg.downsample[i]
        shapes.append((batch, cfg.z_dims[i], ph // f, pw // f))
    return shapes


def n_dropout_calls(cfg: LVAEConfig) -> int:
    return len(dropout_channels(cfg))


def dropout_channels(cfg: LVAEConfig) -> List[int]:
    """Channel count of every Dropout2d call in forward order (all n_filters)."""
    if cfg.dropout is None:
        return []
    per = lambda gated: sum(1 for _, k in _res_block_entries(cfg, gated) if k == "drop")
    L, n = cfg.n_layers, 0
    n += per(False)                                        # stem block
    n += L * cfg.blocks_per_layer * per(cfg.gated)         # bottom-up
    for i in range(L):
        if i != L - 1:
            n += per(True)                                 # merge (residual only)
            if cfg.merge_type != "residual":
                n -= per(True)
            if cfg.stochastic_skip:
                n += per(True)
        n += cfg.blocks_per_layer * per(cfg.gated)
    n += cfg.blocks_per_layer * per(cfg.gated)             # final top-down
    return [cfg.n_filters] * n


# --------------------------------------------------------------------------- helpers (boilr, unpinned)
def pad_img(x, size):
    """boilr.nn.pad_img_tensor (call site lvae.py:324): centred zero pad."""
    dr, dc = size[0] - x.shape[2], size[1] - x.shape[3]
    return F.pad(x, [dc // 2, dc - dc // 2, dr // 2, dr - dr // 2])


def crop_img(x, size):
    """boilr.nn.crop_img_tensor (call sites lvae.py:185,357): inverse of pad_img."""
    dr, dc = x.shape[2] - size[0], x.shape[3] - size[1]
    return x[:, :, dr // 2: x.shape[2] - (dr - dr // 2), dc // 2: x.shape[3] - (dc - dc // 2)]


def free_bits_kl(kl, free_bits, eps=1e-6):
    """boilr.nn.free_bits_kl (call site lvae.py:197): per-sample, per-layer clamp, batch mean."""
    if free_bits < eps:
        return kl.mean(0)
    return kl.clamp(min=free_bits).mean(0)


# --------------------------------------------------------------------------- forward machinery
class _Run:
    def __init__(self, cfg, params, training, eps, masks, generator=None):
        self.cfg, self.P, self.training = cfg, params, training
        self.eps = list(eps) if eps is not None else None
        self.masks = list(masks) if masks is not None else None
        self.gen = generator
        self.act = {"relu": F.relu, "leakyrelu": F.leaky_relu, "elu": F.elu, "selu": F.selu}[cfg.nonlin]

    def conv(self, name, x, stride=1, pad=None):
        w = self.P[name + ".weight"]
        pad = w.shape[-1] // 2 if pad is None else pad
        return F.conv2d(x, w, self.P[name + ".bias"], stride=stride, padding=pad)

    def tconv(self, name, x):          # lvae_layers.py:270-276
        return F.conv_transpose2d(x, self.P[name + ".weight"], self.P[name + ".bias"],
                                  stride=2, padding=1, output_padding=1)

    def bn(self, name, x):             # nn.BatchNorm2d defaults: momentum .1, eps 1e-5
        if self.training:
            self.P[name + ".num_batches_tracked"] += 1
        return F.batch_norm(x, self.P[name + ".running_mean"], self.P[name + ".running_var"],
                            self.P[name + ".weight"], self.P[name + ".bias"],
                            self.training, 0.1, 1e-5)

    def drop(self, x):                 # nn.Dropout2d: whole channels, scale 1/(1-p)
        p = self.cfg.dropout
        if not self.training or p is None or p == 0.0:
            return x
        if self.masks is not None:
            m = self.masks.pop(0)
        else:
            keep = torch.full((x.shape[0], x.shape[1], 1, 1), 1.0 - p, dtype=x.dtype)
            m = torch.bernoulli(keep, generator=self.gen) / (1.0 - p)
        return x * m.to(x.dtype)

    def normal(self, shape, dtype):
        if self.eps is not None:
            e = self.eps.pop(0)
            assert tuple(e.shape) == tuple(shape), (tuple(e.shape), tuple(shape))
            return e.to(dtype)
        return torch.randn(shape, dtype=dtype, generator=self.gen)


def residual_block(r: _Run, prefix, x, gated):
    """ResidualBlock.forward (lib/nn.py:98-99): block(x) + x."""
    h = x
    for idx, kind in _res_block_entries(r.cfg, gated):
        name = "%s.block.%d" % (prefix, idx)
        if kind == "conv":
            h = r.conv(name, h)
        elif kind == "bn":
            h = r.bn(name, h)
        elif kind == "act":
            h = r.act(h)
        elif kind == "drop":
            h = r.drop(h)
        elif kind == "gate":           # GateLayer2d.forward (nn.py:121-126)
            a, g = r.conv(name + ".conv", h).chunk(2, dim=1)
            h = r.act(a) * torch.sigmoid(g)
    return h + x


def resampling_block(r: _Run, prefix, x, resample, gated, top_down):
    """ResBlockWithResampling.forward (lvae_layers.py:300-306)."""
    if resample:
        x = r.tconv(prefix + ".pre_conv", x) if top_down else r.conv(prefix + ".pre_conv", x, stride=2, pad=1)
    return residual_block(r, prefix + ".res", x, gated)


def merge_layer(r: _Run, prefix, x, y, merge_type):
    """MergeLayer.forward (lvae_layers.py:358-360)."""
    h = torch.cat((x, y), dim=1)
    if merge_type == "linear":
        return r.conv(prefix + ".layer", h)
    h = r.conv(prefix + ".layer.0", h)
    return residual_block(r, prefix + ".layer.1", h, True)


_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def normal_log_prob(z, mu, lv):
    """Normal(mu, exp(lv/2)).log_prob(z) as torch.distributions computes it
    (stochastic.py:46,79): -((z-mu)^2)/(2 var) - log(sigma) - log(sqrt(2 pi)),
    with sigma = exp(lv/2), var = sigma^2, log(sigma) = log(exp(lv/2))."""
    sigma = (lv / 2).exp()
    return -((z - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - _HALF_LOG_2PI


def normal_kl(mu_q, lv_q, mu_p, lv_p):
    """kl_divergence(Normal q, Normal p) (stochastic.py:87) in torch's form:
    0.5 * (var_ratio + t1 - 1 - log var_ratio), var_ratio = (sq/sp)^2, t1 = ((mq-mp)/sp)^2."""
    sq, sp = (lv_q / 2).exp(), (lv_p / 2).exp()
    var_ratio = (sq / sp) ** 2
    t1 = ((mu_q - mu_p) / sp) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def stochastic_block(r: _Run, prefix, p_params, q_params, transform_p, analytical_kl,
                     forced_latent=None, use_mode=False, force_constant_output=False):
    """NormalStochasticBlock2d.forward (lib/stochastic.py:29-112)."""
    if transform_p:
        p_params = r.conv(prefix + ".conv_in_p", p_params)          # :39-40
    p_mu, p_lv = p_params.chunk(2, dim=1)                            # :45
    if q_params is not None:
        q_params = r.conv(prefix + ".conv_in_q", q_params)           # :50
        q_mu, q_lv = q_params.chunk(2, dim=1)
        s_mu, s_lv = q_mu, q_lv
    else:
        s_mu, s_lv = p_mu, p_lv
    if forced_latent is not None:                                    # :60-67
        z = forced_latent
    elif use_mode:
        z = s_mu
    else:
        shape = torch.broadcast_shapes(s_mu.shape, s_lv.shape)
        z = s_mu + (s_lv / 2).exp() * r.normal(shape, s_mu.dtype)
    if force_constant_output:                                        # :71-73
        z = z[0:1].expand_as(z).clone()
        p_params = p_params[0:1].expand_as(p_params).clone()
    out = r.conv(prefix + ".conv_out", z)                            # :76
    logprob_p = normal_log_prob(z, p_mu, p_lv).sum((1, 2, 3))        # :79
    if q_params is not None:
        logprob_q = normal_log_prob(z, q_mu, q_lv).sum((1, 2, 3))    # :84
        kl_an = normal_kl(q_mu, q_lv, p_mu, p_lv)                    # :87
        if analytical_kl:
            kl_el = kl_an
        else:                                                        # kl_normal_mc :209-226
            pm, pl = p_params.chunk(2, dim=1)
            kl_el = normal_log_prob(z, q_mu, q_lv) - normal_log_prob(z, pm, pl)
        kl_sample = kl_el.sum((1, 2, 3))                             # :92
        kl_spatial = kl_an.sum(1)                                    # :96
    else:
        logprob_q = kl_el = kl_sample = kl_spatial = None
    return out, dict(z=z, p_params=p_params, q_params=q_params, logprob_p=logprob_p,
                     logprob_q=logprob_q, kl_elementwise=kl_el, kl_samplewise=kl_sample,
                     kl_spatial=kl_spatial)


def bottomup_pass(r: _Run, x):
    """LadderVAE.bottomup_pass (lvae.py:216-227) incl. first_bottom_up (lvae.py:73-84)."""
    cfg = r.cfg
    stride = 1 if cfg.no_initial_downscaling else 2
    h = r.act(r.conv("first_bottom_up.0", x, stride=stride, pad=2))
    h = resampling_block(r, "first_bottom_up.2", h, False, False, False)
    bu = []
    for i in range(cfg.n_layers):
        left = cfg.downsample[i]
        for b in range(cfg.blocks_per_layer):
            h = resampling_block(r, "bottom_up_layers.%d.net.%d" % (i, b), h, left > 0, cfg.gated, False)
            left -= 1 if left > 0 else 0
        bu.append(h)
    return bu


def topdown_layer(r: _Run, i, inp, skip_in, bu_value, n_img_prior=None, forced_latent=None,
                  use_mode=False, constant=False):
    """TopDownLayer.forward (lvae_layers.py:115-178)."""
    cfg, p = r.cfg, "top_down_layers.%d" % i
    top = i == cfg.n_layers - 1
    if top:
        p_params = r.P[p + ".top_prior_params"]                      # :131
        if n_img_prior is not None:
            p_params = p_params.expand(n_img_prior, -1, -1, -1)      # :135-136
    else:
        p_params = inp
    if bu_value is not None:
        q_params = bu_value if top else merge_layer(r, p + ".merge", bu_value, p_params, cfg.merge_type)
    else:
        q_params = None
    x, data = stochastic_block(r, p + ".stochastic", p_params, q_params, not top,
                               cfg.analytical_kl, forced_latent, use_mode, constant)
    if cfg.stochastic_skip and not top:                              # :166-167
        x = merge_layer(r, p + ".skip_connection_merger", x, skip_in, "residual")
    left = cfg.downsample[i]
    for b in range(cfg.blocks_per_layer):                            # :174
        x = resampling_block(r, "%s.deterministic_block.%d" % (p, b), x, left > 0, cfg.gated, True)
        left -= 1 if left > 0 else 0
    return x, data


def topdown_pass(r: _Run, bu_values=None, n_img_prior=None, mode_layers=(), constant_layers=(),
                 forced_latent=None):
    """LadderVAE.topdown_pass (lvae.py:229-315)."""
    cfg = r.cfg
    inference = bu_values is not None
    if inference != (n_img_prior is None):                           # :248-251
        raise RuntimeError("Number of images for top-down generation has to be given "
                           "if and only if we're not doing inference")
    if inference and (len(mode_layers) > 0 or len(constant_layers) > 0):   # :252-255
        raise RuntimeError("Prior experiments (e.g. sampling from mode) are not"
                           " compatible with inference mode")
    L = cfg.n_layers
    z, kl, kls = [None] * L, [None] * L, [None] * L
    forced = forced_latent if forced_latent is not None else [None] * L
    logp, out = 0.0, None
    for i in reversed(range(L)):                                     # :274-303
        out, d = topdown_layer(r, i, out, out, bu_values[i] if inference else None, n_img_prior,
                               forced[i], i in mode_layers, i in constant_layers)
        z[i], kl[i], kls[i] = d["z"], d["kl_samplewise"], d["kl_spatial"]
        logp = logp + d["logprob_p"].mean()
    if not cfg.no_initial_downscaling:                               # Interpolate(scale=2), lvae.py:144
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=False)
    off = 0 if cfg.no_initial_downscaling else 1
    for b in range(cfg.blocks_per_layer):                            # lvae.py:145-156, 306
        out = resampling_block(r, "final_top_down.%d" % (b + off), out, False, cfg.gated, True)
    return out, dict(z=z, kl=kl, kl_spatial=kls, logprob_p=logp)


# --------------------------------------------------------------------------- likelihoods
def bernoulli_log_lik(x, prob):
    """log_bernoulli (likelihoods.py:385-388): -BCE on probabilities; BCE clamps each log at -100."""
    # F.binary_cross_entropy is the very call the reference makes; its backward is the
    # (p - x) / max(p (1 - p), 1e-12) form, finite even where p saturates to 0 or 1.
    return -F.binary_cross_entropy(prob, x, reduction="none").sum((1, 2, 3))


def dmol_log_lik(x01, l, nr_mix=10):
    """-discretized_mix_logistic_loss(2x-1, l) (likelihoods.py:226-230, 291-382), NCHW throughout.
    Channel map (likelihoods.py:305-318): [0:M] mixture logits; colour c: 10+30c+[0:M] means,
    +[M:2M] log-scales (clamped at -7), +[2M:3M] coefficients (tanh)."""
    x = x01 * 2 - 1
    M = nr_mix
    logit = l[:, :M]
    blk = lambda c, j: l[:, M + 3 * M * c + j * M: M + 3 * M * c + (j + 1) * M]
    mean = [blk(c, 0) for c in range(3)]
    ls = [torch.clamp(blk(c, 1), min=-7.0) for c in range(3)]
    co = [torch.tanh(blk(c, 2)) for c in range(3)]
    xc = [x[:, c:c + 1] for c in range(3)]                            # broadcast over mixtures
    m = [mean[0],
         mean[1] + co[0] * xc[0],                                     # :324-325
         mean[2] + co[1] * xc[0] + co[2] * xc[1]]                     # :327-329
    total = 0.0
    for c in range(3):
        cen = xc[c] - m[c]
        inv = torch.exp(-ls[c])
        plus_in = inv * (cen + 1.0 / 255.0)
        min_in = inv * (cen - 1.0 / 255.0)
        cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
        log_cdf_plus = plus_in - F.softplus(plus_in)
        log_one_minus_cdf_min = -F.softplus(min_in)
        mid_in = inv * cen
        log_pdf_mid = mid_in - ls[c] - 2.0 * F.softplus(mid_in)
        c3 = (cdf_delta > 1e-5).to(x.dtype)                           # :367-375 float-mask blends
        inner2 = c3 * torch.log(torch.clamp(cdf_delta, min=1e-12)) + (1 - c3) * (log_pdf_mid - math.log(127.5))
        c2 = (xc[c] > 0.999).to(x.dtype)
        inner = c2 * log_one_minus_cdf_min + (1 - c2) * inner2
        c1 = (xc[c] < -0.999).to(x.dtype)
        total = total + c1 * log_cdf_plus + (1 - c1) * inner
    lp = total + torch.log_softmax(logit, dim=1)                      # :376-377
    return torch.logsumexp(lp, dim=1).sum((1, 2))                     # :378-381 (sign flipped back)


def dmol_sample(l, generator=None, nr_mix=10):
    """sample_from_discretized_mix_logistic (stochastic.py:141-206) then (s+1)/2 clamp
    (likelihoods.py:221-225).  RNG order: mixture-choice uniforms, then logistic uniforms."""
    B, _, H, W = l.shape
    M = nr_mix
    u1 = torch.empty(B, M, H, W, dtype=l.dtype).uniform_(1e-5, 1.0 - 1e-5, generator=generator)
    sel = (l[:, :M] - torch.log(-torch.log(u1))).argmax(dim=1, keepdim=True)
    blk = lambda c, j: l[:, M + 3 * M * c + j * M: M + 3 * M * c + (j + 1) * M].gather(1, sel)
    mean = [blk(c, 0) for c in range(3)]
    ls = [torch.clamp(blk(c, 1), min=-7.0) for c in range(3)]
    co = [torch.tanh(blk(c, 2)) for c in range(3)]
    u = torch.empty(B, 3, H, W, dtype=l.dtype).uniform_(1e-5, 1.0 - 1e-5, generator=generator)
    xs = [mean[c] + torch.exp(ls[c]) * (torch.log(u[:, c:c + 1]) - torch.log(1.0 - u[:, c:c + 1])) for c in range(3)]
    x0 = xs[0].clamp(-1.0, 1.0)
    x1 = (xs[1] + co[0] * x0).clamp(-1.0, 1.0)
    x2 = (xs[2] + co[1] * x0 + co[2] * x1).clamp(-1.0, 1.0)
    return ((torch.cat([x0, x1, x2], dim=1) + 1) / 2).clamp(0.0, 1.0)


def likelihood(r: _Run, h, x):
    """LikelihoodModule.forward (likelihoods.py:33-48) for the two in-scope heads."""
    cfg = r.cfg
    raw = r.conv("likelihood.parameter_net", h)
    if cfg.likelihood_form == "bernoulli":                             # :51-78
        prob = torch.sigmoid(raw)
        with torch.no_grad():
            sample = (torch.rand(prob.shape, dtype=prob.dtype, generator=r.gen) < prob).to(prob.dtype)
        info = dict(mean=prob, mode=torch.round(prob), sample=sample, params=prob)
        ll = None if x is None else bernoulli_log_lik(x, prob)
    else:                                                              # :183-230
        with torch.no_grad():
            sample = dmol_sample(raw, r.gen)
        info = dict(mean=None, mode=None, sample=sample, params=dict(mean=None, all_params=raw))
        ll = None if x is None else dmol_log_lik(x, raw)
    return ll, info


# --------------------------------------------------------------------------- public entry points
def forward(params: Dict[str, torch.Tensor], cfg: LVAEConfig, x, eps=None, masks=None,
            training=True, generator=None) -> dict:
    """LadderVAE.forward (lvae.py:172-214).  ``params`` is mutated in train mode
    (BatchNorm running statistics), exactly like the module's buffers."""
    r = _Run(cfg, params, training, eps, masks, generator)
    x_pad = pad_img(x, cfg.padded_size(x.shape[2:]))                   # :176
    bu = bottomup_pass(r, x_pad)                                       # :179
    out, td = topdown_pass(r, bu)                                      # :182
    out = crop_img(out, x.shape[2:])                                   # :185
    ll, info = likelihood(r, out, x)                                   # :188
    kl = torch.cat([k.unsqueeze(1) for k in td["kl"]], dim=1)          # :192-193
    kl_sep = kl.sum(1)
    return dict(ll=ll, z=td["z"], kl=kl_sep.mean(), kl_sep=kl_sep, kl_avg_layerwise=kl.mean(0),
                kl_spatial=td["kl_spatial"], kl_loss=free_bits_kl(kl, cfg.free_bits).sum(),
                logp=td["logprob_p"], out_mean=info["mean"], out_mode=info["mode"],
                out_sample=info["sample"], likelihood_params=info["params"], kl_layers=kl)


def sample_prior(params, cfg: LVAEConfig, n_imgs, mode_layers=(), constant_layers=(), eps=None,
                 generator=None):
    """LadderVAE.sample_prior (lvae.py:351-362)."""
    r = _Run(cfg, params, False, eps, None, generator)
    out, _ = topdown_pass(r, None, n_imgs, tuple(mode_layers or ()), tuple(constant_layers or ()))
    out = crop_img(out, cfg.img_shape)
    return likelihood(r, out, None)[1]["sample"]


def loss_terms(out: dict, beta: float = 1.0) -> dict:
    """LVAEExperiment.forward_pass (experiment/experiment_manager.py:329-344)."""
    recons_sep = -out["ll"]
    elbo_sep = -(recons_sep + out["kl_sep"])
    return dict(elbo_sep=elbo_sep, elbo=elbo_sep.mean(), recons=recons_sep.mean(),
                loss=recons_sep.mean() + out["kl_loss"] * beta)


def l2_norm(params: Dict[str, torch.Tensor], names: Iterable[str]):
    """experiment_manager.py:346-350: sqrt(sum_p sum(p^2)) over *parameters* (not buffers)."""
    tot = 0.0
    for n in names:
        tot = tot + torch.sum(params[n] ** 2)
    return tot.sqrt()


def trainable_names(cfg: LVAEConfig) -> List[str]:
    """Names ``model.parameters()`` yields (buffers excluded; top prior only if learned
    is still a Parameter -- requires_grad False -- so it IS in parameters(), lvae_layers.py:56-58)."""
    return [n for n in param_shapes(cfg)
            if n.rsplit(".", 1)[-1] not in ("running_mean", "running_var", "num_batches_tracked")]


def adamax_step(p, g, exp_avg, exp_inf, step, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.Adamax single-tensor update (optimizer at experiment_manager.py:78-80)."""
    if wd != 0.0:
        g = g + wd * p
    exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    torch.maximum(exp_inf * b2, g.abs() + eps, out=exp_inf)
    p.addcdiv_(exp_avg, exp_inf, value=-lr / (1 - b1 ** step))


def iw_bound(elbo_sep_samples: torch.Tensor) -> torch.Tensor:
    """boilr VAEExperimentManager.test_procedure's IW estimate (call site evaluate.py:30;
    SURVEY.md 3.3): per image logsumexp_k(elbo_sep[:, k]) - log K.  UNPINNED (boilr)."""
    K = elbo_sep_samples.shape[1]
    return torch.logsumexp(elbo_sep_samples, dim=1) - math.log(K)


class TrainState:
    """Reference training step on CPU: zero_grad, forward, loss, backward, Adamax
    (SURVEY.md 3.1).  Used as the CPU baseline by bench.py and as the checker in tests."""

    def __init__(self, cfg: LVAEConfig, params: Dict[str, torch.Tensor], lr=3e-4):
        self.cfg, self.P = cfg, params
        self.names = trainable_names(cfg)
        for n in self.names:
            frozen = n.endswith("top_prior_params") and not cfg.learn_top_prior
            self.P[n].requires_grad_(not frozen)
        self.opt = torch.optim.Adamax([self.P[n] for n in self.names if self.P[n].requires_grad], lr=lr)
        self.step_count = 0

    def step(self, x, eps=None, masks=None, generator=None, update=True):
        self.opt.zero_grad(set_to_none=True)
        out = forward(self.P, self.cfg, x, eps, masks, True, generator)
        terms = loss_terms(out)
        terms["loss"].backward()
        if update:
            self.opt.step()
            self.step_count += 1
        return out, terms
