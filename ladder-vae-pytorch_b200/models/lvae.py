"""Kernel-backed mirror of the reference's models/lvae.py: ``LadderVAE`` with the same 18
constructor arguments, attributes, submodule names (state_dict layout), methods and output dict
(models/lvae.py:15-372), so the reference's experiment code can use it unchanged.

Activations are physically NHWC between our kernels and logically NCHW at every API boundary.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from lvae_b200 import ops
from lvae_b200.boilr_compat import BaseGenerativeModel
from lvae_b200.lib.likelihoods import (BernoulliLikelihood, DiscretizedLogisticLikelihood,
                                       DiscretizedLogisticMixLikelihood, GaussianLikelihood)
from lvae_b200.lib.nn import Conv2d, Dropout2d, Interpolate, NONLIN
from .lvae_layers import BottomUpDeterministicResBlock, BottomUpLayer, TopDownDeterministicResBlock, TopDownLayer, _BlockStack


class LadderVAE(BaseGenerativeModel):

    def __init__(self, color_ch, z_dims, blocks_per_layer=2, downsample=None, nonlin="elu", merge_type=None,
                 batchnorm=True, stochastic_skip=False, n_filters=32, dropout=None, free_bits=0.0,
                 learn_top_prior=False, img_shape=None, likelihood_form=None, res_block_type=None, gated=False,
                 no_initial_downscaling=False, analytical_kl=False):
        super().__init__()
        self.color_ch = color_ch
        self.z_dims = z_dims
        self.blocks_per_layer = blocks_per_layer
        self.downsample = downsample if downsample is not None else [0] * len(z_dims)
        self.n_layers = len(z_dims)
        self.stochastic_skip = stochastic_skip
        self.n_filters = n_filters
        self.dropout = dropout
        self.free_bits = free_bits
        self.learn_top_prior = learn_top_prior
        self.img_shape = tuple(img_shape)
        self.res_block_type = res_block_type
        self.gated = gated

        # every downsampling step halves the resolution; the stem halves it once more by default
        self.overall_downscale_factor = np.power(2, sum(self.downsample))
        if not no_initial_downscaling:
            self.overall_downscale_factor *= 2
        assert max(self.downsample) <= self.blocks_per_layer
        assert len(self.downsample) == self.n_layers

        act = NONLIN[nonlin]
        block_kw = dict(nonlin=act, batchnorm=batchnorm, dropout=dropout, res_block_type=res_block_type)

        self.first_bottom_up = nn.Sequential(
            Conv2d(color_ch, n_filters, 5, padding=2, stride=1 if no_initial_downscaling else 2),
            act(),
            BottomUpDeterministicResBlock(c_in=n_filters, c_out=n_filters, **block_kw))

        self.top_down_layers = nn.ModuleList([])
        self.bottom_up_layers = nn.ModuleList([])
        for i in range(self.n_layers):
            self.bottom_up_layers.append(BottomUpLayer(
                n_res_blocks=blocks_per_layer, n_filters=n_filters, downsampling_steps=self.downsample[i],
                gated=gated, **block_kw))
            self.top_down_layers.append(TopDownLayer(
                z_dim=z_dims[i], n_res_blocks=blocks_per_layer, n_filters=n_filters,
                is_top_layer=(i == self.n_layers - 1), downsampling_steps=self.downsample[i], merge_type=merge_type,
                stochastic_skip=stochastic_skip, learn_top_prior=learn_top_prior,
                top_prior_param_shape=self.get_top_prior_param_shape(), gated=gated, analytical_kl=analytical_kl,
                **block_kw))

        final = [] if no_initial_downscaling else [Interpolate(scale=2)]
        final += [TopDownDeterministicResBlock(c_in=n_filters, c_out=n_filters, gated=gated, **block_kw)
                  for _ in range(blocks_per_layer)]
        self.final_top_down = _BlockStack(*final)

        if likelihood_form == "bernoulli":
            self.likelihood = BernoulliLikelihood(n_filters, color_ch)
        elif likelihood_form == "gaussian":
            self.likelihood = GaussianLikelihood(n_filters, color_ch)
        elif likelihood_form == "discr_log":
            self.likelihood = DiscretizedLogisticLikelihood(n_filters, color_ch, 256)
        elif likelihood_form == "discr_log_mix":
            self.likelihood = DiscretizedLogisticMixLikelihood(n_filters)
        else:
            raise RuntimeError("Unrecognized likelihood '{}'".format(likelihood_form))

        self._n_dropout_sites = sum(1 for m in self.modules() if isinstance(m, Dropout2d))
        self.compute_dtype = torch.float32

    def set_compute_dtype(self, dtype):
        """torch.float32: exact-fp32 CUDA-core convolutions (parity mode).  torch.bfloat16: activations
        are stored in bf16 and the 64-channel convolutions run on the tensor cores with fp32
        accumulation; the stochastic-block parameters and the likelihood parameters stay fp32."""
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be torch.float32 or torch.bfloat16")
        self.compute_dtype = dtype
        from lvae_b200.lib.stochastic import NormalStochasticBlock2d
        for m in self.modules():
            if isinstance(m, NormalStochasticBlock2d):
                m.compute_dtype = dtype
                m.conv_in_q.spec.out_fp32 = True
                if m.transform_p_params:
                    m.conv_in_p.spec.out_fp32 = True
        self.likelihood.parameter_net.spec.out_fp32 = True
        return self

    # ------------------------------------------------------------------ forward pieces
    def _bn_arena(self, device):
        """One float64 scratch arena for the statistics accumulators of every BatchNorm2d (zeroed once per forward)."""
        arena = getattr(self, "_lvae_bn_arena", None)
        if arena is None or arena.device != device:
            from lvae_b200.lib.nn import BatchNorm2d
            bns = [m for m in self.modules() if isinstance(m, BatchNorm2d)]
            arena = torch.zeros((max(1, len(bns)), 3, ops.BN_STRIPES, 2, self.n_filters), dtype=torch.float64, device=device)
            for i, bn in enumerate(bns):
                if bn.num_features == self.n_filters:
                    bn._lvae_scratch = arena[i]
                    bn._lvae_scratch_owned = False
                    bn._lvae_epoch_fwd = bn._lvae_epoch_bwd = bn._lvae_epoch_out = -1
            self._lvae_bn_arena = arena
        return arena

    def _begin(self, batch, device):
        """Zero the BatchNorm scratch arena and draw every Dropout2d mask of this pass (one launch each)."""
        if self.training:
            self._bn_arena(device).zero_()
            ops.new_forward_epoch()
        if self.training and self.dropout:
            ops.prepare_masks(self._n_dropout_sites, batch, self.n_filters, self.dropout, device)
        else:
            ops.clear_masks()

    def forward(self, x):
        self._begin(x.shape[0], x.device)
        bu_values = self.bottomup_pass(self.pad_input(x))
        return self.forward_from_bottomup(x, bu_values)

    def forward_from_bottomup(self, x, bu_values):
        """Top-down pass, likelihood and KL bookkeeping given the bottom-up activations.  In eval()
        the bottom-up pass is deterministic, so the importance-weighted evaluator computes it once
        per image batch and calls this once per sample (the reference recomputes it K times)."""
        img_size = x.size()[2:]
        out, td = self.topdown_pass(bu_values)
        out = ops.crop(out, img_size)
        ll, lik = self.likelihood(out, x)
        ops.clear_masks()

        # free bits + KL bookkeeping (models/lvae.py:192-198 of the reference) over the (layers, batch) matrix whose rows
        # the stochastic kernels wrote: one launch (lvae_kl_bookkeeping), one more in the backward
        book = td["book"]
        return {
            "ll": ll,
            "z": td["z"],
            "kl": book["kl"],
            "kl_sep": book["kl_sep"],
            "kl_avg_layerwise": book["kl_avg_layerwise"],
            "kl_spatial": td["kl_spatial"],
            "kl_loss": book["kl_loss"],
            "logp": book["logp"],
            "out_mean": lik["mean"],
            "out_mode": lik["mode"],
            "out_sample": lik["sample"],
            "likelihood_params": lik["params"],
        }

    def bottomup_pass(self, x):
        x = self.first_bottom_up(x)
        bu_values = []
        for layer in self.bottom_up_layers:
            x = layer(x)
            bu_values.append(x)
        return bu_values

    def topdown_pass(self, bu_values=None, n_img_prior=None, mode_layers=None, constant_layers=None,
                     forced_latent=None):
        mode_layers = [] if mode_layers is None else mode_layers
        constant_layers = [] if constant_layers is None else constant_layers
        prior_experiment = len(mode_layers) > 0 or len(constant_layers) > 0
        inference_mode = bu_values is not None
        if inference_mode != (n_img_prior is None):
            raise RuntimeError("Number of images for top-down generation has to be given "
                               "if and only if we're not doing inference")
        if inference_mode and prior_experiment:
            raise RuntimeError("Prior experiments (e.g. sampling from mode) are not"
                               " compatible with inference mode")
        L = self.n_layers
        z, kl, kl_spatial = [None] * L, [None] * L, [None] * L
        if forced_latent is None:
            forced_latent = [None] * L
        logprob_p = [None] * L
        out = None
        if inference_mode:
            ops.begin_kl_rows(L, bu_values[-1].shape[0], bu_values[-1].device)
        for i in reversed(range(L)):
            bu_value = bu_values[i] if inference_mode else None
            out, _pre_residual, aux = self.top_down_layers[i](
                out, skip_connection_input=out, inference_mode=inference_mode, bu_value=bu_value,
                n_img_prior=n_img_prior, use_mode=i in mode_layers, force_constant_output=i in constant_layers,
                forced_latent=forced_latent[i])
            z[i], kl[i], kl_spatial[i] = aux["z"], aux["kl_samplewise"], aux["kl_spatial"]
            logprob_p[i] = aux["logprob_p"]
        out = self.final_top_down(out)
        data = {"z": z, "kl": kl, "kl_spatial": kl_spatial}
        if inference_mode:
            # sum over layers of the batch means of log p(z) (lvae.py:301-302) comes out of the same launch as the KL terms
            data["book"] = ops.kl_bookkeeping(kl, logprob_p, self.free_bits, rows=ops.end_kl_rows())
            data["logprob_p"] = data["book"]["logp"]
        else:
            data["logprob_p"] = torch.stack(logprob_p, dim=0).mean(dim=1).sum()
        return out, data

    def pad_input(self, x):
        """Zero-pad (centred) to the next multiple of the overall downscale factor and hand the
        image to the kernels as an NHWC activation."""
        return ops.pad_image(x, self.get_padded_size(x.size()), out_dtype=self.compute_dtype)

    def get_padded_size(self, size):
        dwnsc = self.overall_downscale_factor
        if len(size) == 4:
            size = size[2:]
        if len(size) != 2:
            raise RuntimeError("input size must be either (N, C, H, W) or (H, W), but it "
                               "has length {} (size={})".format(len(size), size))
        return list(int(((s - 1) // dwnsc + 1) * dwnsc) for s in size)

    def sample_prior(self, n_imgs, mode_layers=None, constant_layers=None):
        ops.clear_masks()
        out, _ = self.topdown_pass(n_img_prior=n_imgs, mode_layers=mode_layers, constant_layers=constant_layers)
        out = ops.crop(out, self.img_shape)
        _, lik = self.likelihood(out, None)
        return lik["sample"]

    def get_top_prior_param_shape(self, n_imgs=1):
        dwnsc = self.overall_downscale_factor
        sz = self.get_padded_size(self.img_shape)
        return (n_imgs, self.z_dims[-1] * 2, sz[0] // dwnsc, sz[1] // dwnsc)
