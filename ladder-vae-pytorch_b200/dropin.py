"""Run the reference's own entry points (main.py, evaluate.py) on the kernel-backed modules, unchanged:

    cd /path/to/ladder-vae-pytorch
    PYTHONPATH=/path/to/repo python -m lvae_b200.dropin main.py --dataset static_mnist --zdims 32 32 32 ...

`python main.py` puts the script's directory first on sys.path, and the reference's ``models`` is a regular package
(models/__init__.py), so prepending the mirror to PYTHONPATH cannot shadow it.  install() therefore registers the mirror
under the reference's module names in sys.modules BEFORE anything imports them -- ``models``, ``models.lvae``,
``models.lvae_layers``, ``lib``, ``lib.nn``, ``lib.stochastic``, ``lib.likelihoods`` (the five files SURVEY.md section 8
puts on the hot path) -- which wins over any path lookup.  ``lib.datasets`` (data loading, not replaced) still resolves to
the reference's file: the mirror's ``lib`` package searches the other ``lib`` directories on sys.path after its own.
experiment/experiment_manager.py:11 (`from models.lvae import LadderVAE`) then gets the sm_100a LadderVAE.
"""
from __future__ import annotations

import os
import runpy
import sys


def install() -> None:
    import lvae_b200  # noqa: F401  (the package: loads liblvae_b200.so; raises if it was not built)
    from lvae_b200 import lib as _lib, models as _models
    from lvae_b200.lib import likelihoods, nn, stochastic
    from lvae_b200.models import lvae, lvae_layers
    here = os.path.dirname(os.path.abspath(_lib.__file__))
    for p in list(sys.path):
        cand = os.path.join(os.path.abspath(p or "."), "lib")
        if os.path.isdir(cand) and os.path.abspath(cand) != here and cand not in _lib.__path__:
            _lib.__path__.append(cand)
    sys.modules.update({"models": _models, "models.lvae": lvae, "models.lvae_layers": lvae_layers, "lib": _lib,
                        "lib.nn": nn, "lib.stochastic": stochastic, "lib.likelihoods": likelihoods})


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m lvae_b200.dropin <reference script, e.g. main.py> [its arguments]")
    script = os.path.abspath(argv[0])
    sys.path.insert(0, os.path.dirname(script))       # what `python script.py` does
    install()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
