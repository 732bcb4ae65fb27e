"""lvae_b200 -- B200-native (sm_100a) implementation of the Ladder VAE hot path.

Public surface mirrors the reference's module API (models.lvae.LadderVAE, models.lvae_layers.*,
lib.nn / lib.stochastic / lib.likelihoods) on top of hand-written CUDA kernels reached through a
C ABI (include/lvae_b200.h, liblvae_b200.so).  There is no CPU or ATen fallback for the hot ops.
"""
from . import _capi  # noqa: F401
from . import ops  # noqa: F401
from .ops import inject, manual_seed  # noqa: F401
from .models.lvae import LadderVAE  # noqa: F401
from .models.lvae_layers import (BottomUpLayer, MergeLayer, ResBlockWithResampling,  # noqa: F401
                                 SkipConnectionMerger, TopDownLayer)
from .lib.nn import GateLayer2d, ResidualBlock, ResidualGatedBlock  # noqa: F401
from .lib.stochastic import NormalStochasticBlock2d  # noqa: F401
from .lib.likelihoods import BernoulliLikelihood, DiscretizedLogisticMixLikelihood  # noqa: F401

__all__ = ["LadderVAE", "TopDownLayer", "BottomUpLayer", "NormalStochasticBlock2d", "ResidualBlock",
           "ResidualGatedBlock", "GateLayer2d", "BernoulliLikelihood", "DiscretizedLogisticMixLikelihood",
           "inject", "manual_seed", "ops"]
