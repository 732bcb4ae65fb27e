"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/dropin_small.npz from the UNMODIFIED reference model driven through the
UNMODIFIED reference experiment layer (experiment/experiment_manager.py) on the boilr stand-in of tests/dropin:

    python -m oracle.make_golden_dropin          # build container only (needs /root/reference)

Recorded (float64): (a) the parameters after boilr-style data-dependent initialisation (experiment_manager.py:62-72;
Kaiming re-initialisation replaced by the seeded weights so that the fixture is platform independent); (b) one training
step as main.py runs it: LVAEExperiment.forward_pass (:322-367) -> backward -> torch.optim.Adamax (:76-81); (c) the
stand-in's test_procedure with K = 3 importance samples (call site evaluate.py:30).  tests/test_dropin_gpu.py replays
the same three stages on the kernel-backed model."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

from . import lvae_oracle as O
from . import ref_loader
from .make_golden import GOLDEN_DIR, make_inputs, small_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = dict(batch=4, weight_seed=21, input_seed=201, n_iw=3)


def dropin_cfg():
    return small_cfg(dropout=0.0)


def import_reference_experiment(ref_mods):
    """experiment.experiment_manager from /root/reference, bound to the reference's own models.lvae."""
    models = types.ModuleType("models")
    models.__path__ = []
    models.lvae = ref_mods["lvae"]
    saved = {k: sys.modules.get(k) for k in ("models", "models.lvae", "boilr", "boilr.nn", "boilr.models", "boilr.data",
                                             "boilr.utils", "boilr.nn.init", "multiobject", "multiobject.pytorch", "lib",
                                             "lib.datasets", "experiment", "experiment.data", "experiment.experiment_manager")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.modules["models"], sys.modules["models.lvae"] = models, ref_mods["lvae"]
    paths = [os.path.join(ROOT, "tests", "dropin"), ref_loader.REFERENCE_ROOT]
    for p in reversed(paths):
        sys.path.insert(0, p)
    try:
        import importlib
        em = importlib.import_module("experiment.experiment_manager")
        ddi = importlib.import_module("boilr.nn.init").data_dependent_init
        base = importlib.import_module("boilr").VAEExperimentManager
    finally:
        for p in paths:
            sys.path.remove(p)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return em, ddi, base


def param_stats(named):
    names, s, l2 = [], [], []
    for n, p in named:
        names.append(n)
        s.append(p.detach().double().sum().item())
        l2.append(p.detach().double().pow(2).sum().sqrt().item())
    return np.array(names), np.array(s), np.array(l2)


def main():
    if not ref_loader.reference_available():
        sys.exit("reference tree not present; golden files can only be made in the build container")
    ref = ref_loader.load_reference()
    em, ddi, _ = import_reference_experiment(ref)
    cfg = dropin_cfg()
    B = CASE["batch"]
    dt = torch.float64
    model = ref["lvae"].LadderVAE(**cfg.kwargs()).to(dt)
    model.load_state_dict(O.make_params(cfg, CASE["weight_seed"], dt))
    x, eps, _ = make_inputs(cfg, B, CASE["input_seed"], True, CASE["n_iw"] + 2)
    x = x.to(dt)
    blob = {"meta": np.array(json.dumps(dict(cfg=cfg.kwargs(), **CASE)))}
    # (a) data-dependent init: keep the seeded weights where boilr would draw Kaiming-normal ones
    orig = torch.nn.init.kaiming_normal_
    torch.nn.init.kaiming_normal_ = lambda t, *a, **k: t
    try:
        with ref_loader.EpsQueue([e.to(dt) for e in eps[0]]):
            ddi(model, {"x": x})
    finally:
        torch.nn.init.kaiming_normal_ = orig
    blob["ddi_names"], blob["ddi_sum"], blob["ddi_l2"] = param_stats(model.named_parameters())
    # (b) one training step through the reference's own forward_pass
    exp = em.LVAEExperiment.__new__(em.LVAEExperiment)
    exp.args = types.SimpleNamespace(beta_anneal=0, lr=3e-4, weight_decay=0.0)
    exp.device = torch.device("cpu")
    exp.model = model.train()
    exp.optimizer = exp._make_optimizer()
    exp.optimizer.zero_grad()
    with ref_loader.EpsQueue([e.to(dt) for e in eps[1]]):
        out = exp.forward_pass(x)
    for k in ("loss", "elbo", "kl", "l2", "recons"):
        blob["step_" + k] = np.float64(out[k].item())
    blob["step_elbo_sep"] = out["elbo_sep"].detach().numpy()
    blob["step_kl_avg_layerwise"] = out["kl_avg_layerwise"].detach().numpy()
    out["loss"].backward()
    exp.optimizer.step()
    _, blob["stepped_sum"], blob["stepped_l2"] = param_stats(model.named_parameters())
    # (c) importance-weighted bound through the stand-in's test_procedure (K full forward passes)
    model.eval()
    exp.dataloaders = types.SimpleNamespace(test=[(x, None)])
    flat = [e.to(dt) for k in range(2, 2 + CASE["n_iw"]) for e in eps[k]]
    with torch.no_grad(), ref_loader.EpsQueue(flat):
        res = exp.test_procedure(iw_samples=CASE["n_iw"])
    blob["iw_elbo"] = np.float64(res["elbo/elbo"])
    blob["iw_bound"] = np.float64(res["elbo/elbo_IW_%d" % CASE["n_iw"]])
    path = os.path.join(GOLDEN_DIR, "dropin_small.npz")
    np.savez_compressed(path, **blob)
    print("wrote %s (%.1f KB): loss %.6f, l2 %.6f, iw %.6f" % (path, os.path.getsize(path) / 1024, blob["step_loss"], blob["step_l2"],
                                                               blob["iw_bound"]))


if __name__ == "__main__":
    main()
