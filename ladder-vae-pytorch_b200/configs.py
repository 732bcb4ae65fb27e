"""The model configurations BASELINE.json names, as LadderVAE constructor arguments (reference README.md:21,175-195 and
the defaults of experiment/experiment_manager.py:153-253).  Product-side copy: bench.py and the
profiling scripts build their models from here, the oracle keeps its own table (tests check that the two agree)."""
from __future__ import annotations

from types import SimpleNamespace

_COMMON = dict(blocks_per_layer=4, n_filters=64, nonlin="elu", gated=True, stochastic_skip=True, merge_type="residual",
               res_block_type="bacdbacd", dropout=0.2, batchnorm=True, learn_top_prior=True, analytical_kl=False,
               no_initial_downscaling=False)

_TABLE = {
    "mnist3": dict(color_ch=1, z_dims=[32] * 3, img_shape=(28, 28), downsample=[1, 1, 1], free_bits=0.5,
                   likelihood_form="bernoulli"),
    "mnist12": dict(color_ch=1, z_dims=[32] * 12, img_shape=(28, 28), downsample=[0, 0, 0, 1] * 3, free_bits=1.0,
                    likelihood_form="bernoulli"),
    "cifar15": dict(color_ch=3, z_dims=[32] * 15, img_shape=(32, 32), downsample=[0, 0, 0, 0, 1] * 3, free_bits=1.0,
                    likelihood_form="discr_log_mix"),
    "celeba20": dict(color_ch=3, z_dims=[32] * 20, img_shape=(64, 64), downsample=[0, 0, 0, 0, 1] * 4, free_bits=1.0,
                     likelihood_form="discr_log_mix"),
}


def baseline_kwargs(name: str) -> dict:
    """Keyword arguments of lvae_b200.LadderVAE for one of 'mnist3', 'mnist12', 'cifar15', 'celeba20'."""
    kw = dict(_COMMON)
    kw.update(_TABLE[name])
    kw["z_dims"] = list(kw["z_dims"])
    kw["downsample"] = list(kw["downsample"])
    return kw


def baseline_config(name: str) -> SimpleNamespace:
    """The same as an attribute bag with a .kwargs() method (what bench.py's synthetic-input helper reads)."""
    kw = baseline_kwargs(name)
    ns = SimpleNamespace(**kw)
    ns.kwargs = lambda: dict(kw)
    return ns
