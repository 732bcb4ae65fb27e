"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For every case below the reference ``models.lvae.LadderVAE`` is built with the
case's constructor arguments, loaded with ``oracle.lvae_oracle.make_params``
weights (numpy RandomState -> identical on every machine), run in float64 and
float32 on inputs from ``make_inputs`` with injected eps / dropout masks, and
its outputs are stored: ll, per-layer KL, kl_loss, loss, z and kl_spatial
checksums, and (sum, L2) of every parameter gradient.  Inputs and weights are
NOT stored; tests regenerate them from the seeds in the file.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys

import numpy as np
import torch

from . import lvae_oracle as O
from . import ref_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def small_cfg(**kw):
    base = dict(color_ch=3, z_dims=[8, 8, 8], img_shape=(16, 16), blocks_per_layer=2, downsample=[0, 1, 1],
                n_filters=16, dropout=0.0, free_bits=0.5, learn_top_prior=True,
                likelihood_form="discr_log_mix", res_block_type="bacdbacd", gated=True,
                stochastic_skip=True, merge_type="residual")
    base.update(kw)
    return O.LVAEConfig(**base)


def cases():
    """name -> (cfg, batch, training, weight_seed, input_seed, n_iw_samples)"""
    c = {}
    c["mnist3_train_b4"] = (O.baseline_config("mnist3"), 4, True, 11, 101, 0)
    c["mnist3_eval_b4"] = (O.baseline_config("mnist3"), 4, False, 11, 102, 3)
    c["mnist12_eval_b2"] = (O.baseline_config("mnist12"), 2, False, 12, 103, 4)
    c["mnist12_train_b2"] = (O.baseline_config("mnist12"), 2, True, 12, 107, 0)
    c["cifar15_train_b2"] = (O.baseline_config("cifar15"), 2, True, 13, 104, 0)
    c["celeba20_train_b1"] = (O.baseline_config("celeba20"), 2, True, 14, 105, 0)
    c["small_dmol_train_b4"] = (small_cfg(), 4, True, 15, 106, 0)
    c["small_dmol_eval_b4"] = (small_cfg(), 4, False, 15, 106, 2)
    c["small_bern_bacdbac"] = (small_cfg(color_ch=1, likelihood_form="bernoulli", img_shape=(12, 10),
                                         res_block_type="bacdbac", dropout=0.3, z_dims=[4, 6], downsample=[1, 0],
                                         blocks_per_layer=1, nonlin="leakyrelu", free_bits=0.0,
                                         learn_top_prior=False), 3, True, 16, 108, 0)
    c["small_bern_cabdcabd_linear"] = (small_cfg(color_ch=1, likelihood_form="bernoulli", img_shape=(12, 10),
                                                 res_block_type="cabdcabd", dropout=0.1, z_dims=[4, 6],
                                                 downsample=[1, 0], blocks_per_layer=1, nonlin="relu",
                                                 gated=False, merge_type="linear", stochastic_skip=False,
                                                 analytical_kl=True, no_initial_downscaling=True), 3, True, 17, 109, 0)
    c["small_dmol_nobn_selu"] = (small_cfg(batchnorm=False, nonlin="selu", res_block_type="bacdbac", dropout=None,
                                           analytical_kl=True), 2, True, 18, 110, 0)
    return c


def make_inputs(cfg: O.LVAEConfig, batch: int, seed: int, training: bool, n_samples: int = 1):
    """Synthetic inputs, platform independent.  Returns float64 tensors:
    x; eps[k] = list of per-layer noise in top->bottom (execution) order; masks = list of
    Dropout2d keep masks (B,C,1,1) scaled by 1/(1-p) in execution order (train only)."""
    rng = np.random.RandomState(seed)
    shp = (batch, cfg.color_ch) + tuple(cfg.img_shape)
    if cfg.likelihood_form == "bernoulli":
        x = (rng.random_sample(shp) < 0.15).astype(np.float64)
    else:
        x = rng.randint(0, 256, size=shp).astype(np.float64) / 255.0
    eps = [[torch.from_numpy(rng.standard_normal(s)) for s in reversed(O.latent_shapes(cfg, batch))]
           for _ in range(max(1, n_samples))]
    masks = None
    if training and cfg.dropout:
        p = cfg.dropout
        masks = [torch.from_numpy((rng.random_sample((batch, c, 1, 1)) >= p).astype(np.float64) / (1.0 - p))
                 for c in O.dropout_channels(cfg)]
    return torch.from_numpy(x), eps, masks


def run_reference(cfg, batch, training, wseed, iseed, n_iw, dtype):
    ref = ref_loader.load_reference()
    model = ref["lvae"].LadderVAE(**cfg.kwargs()).to(dtype)
    model.load_state_dict(O.make_params(cfg, wseed, dtype))
    model.train(training)
    x, eps, masks = make_inputs(cfg, batch, iseed, training, n_iw)
    x = x.to(dtype)
    res = {}
    mctx = ref_loader.DropoutMaskQueue([m.to(dtype) for m in masks]) if masks else contextlib.nullcontext()
    with torch.set_grad_enabled(training), ref_loader.EpsQueue([e.to(dtype) for e in eps[0]]), mctx:
        out = model(x)
    kl_layers = torch.stack([k for k in out["kl_spatial"]], 0) if False else None
    recons_sep = -out["ll"]
    loss = recons_sep.mean() + out["kl_loss"]          # experiment_manager.py:339-344, beta = 1
    res["ll"] = out["ll"].detach().double().numpy()
    res["kl_sep"] = out["kl_sep"].detach().double().numpy()
    res["kl_avg_layerwise"] = out["kl_avg_layerwise"].detach().double().numpy()
    res["kl_loss"] = np.float64(out["kl_loss"].item())
    res["kl"] = np.float64(out["kl"].item())
    res["logp"] = np.float64(float(out["logp"]))
    res["loss"] = np.float64(loss.item())
    res["z_sum"] = np.array([z.double().sum().item() for z in out["z"]])
    res["z_abs"] = np.array([z.double().abs().sum().item() for z in out["z"]])
    res["kl_spatial_sum"] = np.array([k.double().sum().item() for k in out["kl_spatial"]])
    lp = out["likelihood_params"]
    lp = lp["all_params"] if isinstance(lp, dict) else lp
    res["lik_params_sum"] = np.float64(lp.double().sum().item())
    res["lik_params_abs"] = np.float64(lp.double().abs().sum().item())
    if training:
        loss.backward()
        names, gs, gn = [], [], []
        for n, p in model.named_parameters():
            names.append(n)
            g = p.grad.double() if p.grad is not None else torch.zeros(())
            gs.append(g.sum().item())
            gn.append(g.pow(2).sum().sqrt().item())
        res["grad_names"] = np.array(names)
        res["grad_sum"] = np.array(gs)
        res["grad_l2"] = np.array(gn)
        sd = model.state_dict()
        rn = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
        res["running_names"] = np.array(rn)
        res["running_sum"] = np.array([sd[k].double().sum().item() for k in rn])
    if n_iw:
        # IW bound the way boilr's test_procedure does it: K independent forwards.
        cols = []
        with torch.no_grad():
            for k in range(n_iw):
                with ref_loader.EpsQueue([e.to(dtype) for e in eps[k]]):
                    o = model(x)
                cols.append((o["ll"] - o["kl_sep"]).double())
        m = torch.stack(cols, 1)
        res["elbo_sep_samples"] = m.numpy()
        res["iw_bound"] = (torch.logsumexp(m, 1) - np.log(n_iw)).numpy()
    return res


def main(only=None):
    if not ref_loader.reference_available():
        sys.exit("reference tree not present; golden files can only be made in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    for name, (cfg, batch, training, wseed, iseed, n_iw) in cases().items():
        if only and name not in only:
            continue
        blob = {"meta": np.array(json.dumps(dict(cfg=cfg.kwargs(), batch=batch, training=training,
                                                 weight_seed=wseed, input_seed=iseed, n_iw=n_iw)))}
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            r = run_reference(cfg, batch, training, wseed, iseed, n_iw, dt)
            for k, v in r.items():
                blob["%s_%s" % (tag, k)] = v
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **blob)
        print("wrote %s (%.1f KB) loss=%.6f" % (path, os.path.getsize(path) / 1024, blob["f64_loss"]))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
