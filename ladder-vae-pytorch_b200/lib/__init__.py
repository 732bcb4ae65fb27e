"""Kernel-backed mirror of the reference's ``lib`` package (nn, stochastic, likelihoods).

When this directory shadows the reference's ``lib`` on sys.path (drop-in route), modules this mirror does
not replace -- ``lib.datasets``, the data loading the experiment layer imports (experiment/data.py:6) -- are still found
in the reference's own ``lib`` directory: every other ``lib`` directory on sys.path is appended to this package's search
path (ours stays first)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(os.path.abspath(_p or "."), "lib")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
