"""Kernel-backed mirror of the reference's models/lvae_layers.py: TopDownLayer, BottomUpLayer,
ResBlockWithResampling (+ the two deterministic aliases), MergeLayer, SkipConnectionMerger.
Same constructor arguments, submodule names (state_dict keys), forward signatures and
exceptions (models/lvae_layers.py:8-376)."""
from __future__ import annotations

import torch
from torch import nn

from lvae_b200 import ops
from lvae_b200.lib.nn import Conv2d, ConvTranspose2d, LeakyReLU, ResidualBlock, ResidualGatedBlock, _hooked
from lvae_b200.lib.stochastic import NormalStochasticBlock2d


class ResBlockWithResampling(nn.Module):
    """[strided / transposed 3x3 or 1x1 pre-conv] -> ResidualBlock -> [1x1 post-conv]
    (models/lvae_layers.py:222-306).  Bottom-up blocks halve, top-down blocks double the resolution."""

    def __init__(self, mode, c_in, c_out, nonlin=LeakyReLU, resample=False, res_block_kernel=None, groups=1,
                 batchnorm=True, res_block_type=None, dropout=None, min_inner_channels=None, gated=None):
        super().__init__()
        assert mode in ["top-down", "bottom-up"]
        inner = max(c_out, min_inner_channels or 0)
        if resample and mode == "bottom-up":
            self.pre_conv = Conv2d(c_in, inner, kernel_size=3, padding=1, stride=2, groups=groups)
        elif resample:
            self.pre_conv = ConvTranspose2d(c_in, inner, kernel_size=3, padding=1, stride=2, groups=groups,
                                            output_padding=1)
        elif c_in != inner:
            self.pre_conv = Conv2d(c_in, inner, 1, groups=groups)
        else:
            self.pre_conv = None
        self.res = ResidualBlock(channels=inner, nonlin=nonlin, kernel=res_block_kernel, groups=groups,
                                 batchnorm=batchnorm, dropout=dropout, gated=gated, block_type=res_block_type)
        self.post_conv = Conv2d(inner, c_out, 1, groups=groups) if inner != c_out else None

    def forward(self, x):
        if self.pre_conv is not None:
            x = self.pre_conv(x)
        x = self.res(x)
        if self.post_conv is not None:
            x = self.post_conv(x)
        return x


class TopDownDeterministicResBlock(ResBlockWithResampling):

    def __init__(self, *args, upsample=False, **kwargs):
        kwargs["resample"] = upsample
        super().__init__("top-down", *args, **kwargs)


class BottomUpDeterministicResBlock(ResBlockWithResampling):

    def __init__(self, *args, downsample=False, **kwargs):
        kwargs["resample"] = downsample
        super().__init__("bottom-up", *args, **kwargs)


class _BlockStack(nn.Sequential):
    """nn.Sequential of residual blocks (same child indices, hence the same state_dict keys) that tells the kernel layer when
    a block's output is consumed by the next gated block of the stack and by nothing else -- no resampling / 1x1 conv in
    between, no hooks: the two blocks' adjacent elementwise backward passes then run as one launch (ops._pending_bn1)."""

    def forward(self, x):
        mods = list(self)
        for i, m in enumerate(mods):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            ops.mark_block_output_private(
                nxt is not None and isinstance(m, ResBlockWithResampling) and isinstance(nxt, ResBlockWithResampling)
                and m.post_conv is None and nxt.pre_conv is None and not _hooked(m) and not _hooked(nxt)
                and not self._forward_hooks)
            x = m(x)
            ops.mark_block_output_private(False)
        return x


def _resampling_stack(block_cls, n_blocks, n_filters, n_resample, flag, **kw):
    """n_blocks residual blocks, the first n_resample of which change resolution."""
    return [block_cls(n_filters, n_filters, **{flag: i < n_resample}, **kw) for i in range(n_blocks)]


class BottomUpLayer(nn.Module):
    """Deterministic inference layer: a stack of bottom-up residual blocks (models/lvae_layers.py:181-219)."""

    def __init__(self, n_res_blocks, n_filters, downsampling_steps=0, nonlin=None, batchnorm=True, dropout=None,
                 res_block_type=None, gated=None):
        super().__init__()
        self.net = _BlockStack(*_resampling_stack(
            BottomUpDeterministicResBlock, n_res_blocks, n_filters, downsampling_steps, "downsample",
            nonlin=nonlin, batchnorm=batchnorm, dropout=dropout, res_block_type=res_block_type, gated=gated))

    def forward(self, x):
        return self.net(x)


class MergeLayer(nn.Module):
    """Merge two maps: 1x1 conv over their channel concatenation, optionally followed by a gated
    residual block (models/lvae_layers.py:323-360).  The concatenation is never materialised: the
    conv kernel reads both inputs."""

    def __init__(self, channels, merge_type, nonlin=LeakyReLU, batchnorm=True, dropout=None, res_block_type=None):
        super().__init__()
        try:
            iter(channels)
        except TypeError:
            channels = [channels] * 3
        else:
            if len(channels) == 1:
                channels = [channels[0]] * 3
        assert len(channels) == 3
        if merge_type == "linear":
            self.layer = Conv2d(channels[0] + channels[1], channels[2], 1)
        elif merge_type == "residual":
            self.layer = nn.Sequential(
                Conv2d(channels[0] + channels[1], channels[2], 1, padding=0),
                ResidualGatedBlock(channels[2], nonlin, batchnorm=batchnorm, dropout=dropout,
                                   block_type=res_block_type))

    def forward(self, x, y):
        if _hooked(self):
            return self.layer(torch.cat((x, y), dim=1))
        if isinstance(self.layer, nn.Sequential):
            blk = self.layer[1]
            first = blk.block[0] if getattr(blk, "_whole_block", False) else None      # BatchNorm that opens the block
            return blk(self.layer[0](x, x2=y, stats_bn=first if (first is not None and self.training) else None))
        return self.layer(x, x2=y)


class SkipConnectionMerger(MergeLayer):
    """Merge layer around the stochastic node; always the 'residual' kind (models/lvae_layers.py:363-376)."""

    merge_type = "residual"

    def __init__(self, channels, nonlin, batchnorm, dropout, res_block_type):
        super().__init__(channels, self.merge_type, nonlin, batchnorm, dropout=dropout, res_block_type=res_block_type)


class TopDownLayer(nn.Module):
    """One rung of the generative ladder (models/lvae_layers.py:8-178).

    Inference: q_params = merge(bottom-up value, p_params) (top layer: the bottom-up value itself,
    p_params = learned prior); z ~ q; optional skip merge with the layer above; deterministic
    top-down residual stack with upsampling.  Generation: z ~ p."""

    def __init__(self, z_dim, n_res_blocks, n_filters, is_top_layer=False, downsampling_steps=None, nonlin=None,
                 merge_type=None, batchnorm=True, dropout=None, stochastic_skip=False, res_block_type=None,
                 gated=None, learn_top_prior=False, top_prior_param_shape=None, analytical_kl=False):
        super().__init__()
        self.is_top_layer = is_top_layer
        self.z_dim = z_dim
        self.stochastic_skip = stochastic_skip
        self.learn_top_prior = learn_top_prior
        self.analytical_kl = analytical_kl
        if is_top_layer:
            self.top_prior_params = nn.Parameter(torch.zeros(top_prior_param_shape), requires_grad=learn_top_prior)
        self.deterministic_block = _BlockStack(*_resampling_stack(
            TopDownDeterministicResBlock, n_res_blocks, n_filters, downsampling_steps or 0, "upsample",
            nonlin=nonlin, batchnorm=batchnorm, dropout=dropout, res_block_type=res_block_type, gated=gated))
        self.stochastic = NormalStochasticBlock2d(c_in=n_filters, c_vars=z_dim, c_out=n_filters,
                                                  transform_p_params=(not is_top_layer))
        if not is_top_layer:
            self.merge = MergeLayer(channels=n_filters, merge_type=merge_type, nonlin=nonlin, batchnorm=batchnorm,
                                    dropout=dropout, res_block_type=res_block_type)
            if stochastic_skip:
                self.skip_connection_merger = SkipConnectionMerger(
                    channels=n_filters, nonlin=nonlin, batchnorm=batchnorm, dropout=dropout,
                    res_block_type=res_block_type)

    def forward(self, input_=None, skip_connection_input=None, inference_mode=False, bu_value=None,
                n_img_prior=None, forced_latent=None, use_mode=False, force_constant_output=False):
        if self.is_top_layer and not (input_ is None and skip_connection_input is None):
            raise ValueError("In top layer, inputs should be None")
        if self.is_top_layer:
            p_params = self.top_prior_params
            if n_img_prior is not None:
                p_params = p_params.expand(n_img_prior, -1, -1, -1)
        else:
            p_params = input_
        q_params = None
        if inference_mode:
            q_params = bu_value if self.is_top_layer else self.merge(bu_value, p_params)
        x, data_stoch = self.stochastic(p_params=p_params, q_params=q_params, forced_latent=forced_latent,
                                        use_mode=use_mode, force_constant_output=force_constant_output,
                                        analytical_kl=self.analytical_kl)
        if self.stochastic_skip and not self.is_top_layer:
            x = self.skip_connection_merger(x, skip_connection_input)
        x_pre_residual = x
        x = self.deterministic_block(x)
        data = {k: data_stoch[k] for k in ("z", "kl_samplewise", "kl_spatial", "logprob_p", "logprob_q")}
        return x, x_pre_residual, data
