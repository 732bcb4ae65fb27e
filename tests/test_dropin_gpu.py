"""GPU half of the drop-in route (tests/test_dropin_cpu.py is the import-level half): the three things main.py /
evaluate.py do with the model, replayed on the kernel-backed LadderVAE through the boilr stand-in of tests/dropin and
compared with tests/golden/dropin_small.npz, which oracle/make_golden_dropin.py recorded from the UNMODIFIED reference
model driven by the UNMODIFIED reference experiment layer:
  (a) boilr-style data-dependent initialisation (experiment_manager.py:62-72): forward hooks on every Conv2d /
      ConvTranspose2d rewrite weights, biases and outputs -> the hooked, module-by-module path of the mirror;
  (b) one training step as boilr's loop runs it: forward_pass (:322-367) -> backward -> torch.optim.Adamax (:76-81),
      i.e. plain autograd + a stock torch optimizer on the mirror's parameters (no TrainEngine);
  (c) the importance-weighted bound through test_procedure (K full forward passes, call site evaluate.py:30)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import lvae_oracle as O
from lvae_test_helpers import load_golden, make_inputs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _standin():
    sys.path.insert(0, os.path.join(HERE, "dropin"))
    try:
        saved = {k: sys.modules.pop(k, None) for k in list(sys.modules) if k == "boilr" or k.startswith("boilr.")}
        import importlib
        boilr = importlib.import_module("boilr")
        ddi = importlib.import_module("boilr.nn.init").data_dependent_init
    finally:
        sys.path.remove(os.path.join(HERE, "dropin"))
        for k in [k for k in sys.modules if k == "boilr" or k.startswith("boilr.")]:
            sys.modules.pop(k)
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
    return boilr, ddi


def _stats(model):
    s, l2 = [], []
    for _, p in model.named_parameters():
        s.append(p.detach().double().sum().item())
        l2.append(p.detach().double().pow(2).sum().sqrt().item())
    return np.array(s), np.array(l2)


def _close(a, b, tol):
    scale = np.maximum(np.abs(b), 1e-3 * np.abs(b).max())
    return float(np.max(np.abs(a - b) / scale)) < tol, float(np.max(np.abs(a - b) / scale))


def test_dropin_init_step_and_iw_match_the_reference_run():
    import lvae_b200
    from dropin_experiment import forward_pass
    boilr, ddi = _standin()
    cfg, meta, g = load_golden("dropin_small")
    B, K = meta["batch"], meta["n_iw"]
    x, eps, _ = make_inputs(cfg, B, meta["input_seed"], True, K + 2)
    xd = x.float().cuda()
    model = lvae_b200.LadderVAE(**cfg.kwargs())
    model.load_state_dict(O.make_params(cfg, meta["weight_seed"]), strict=True)
    model = model.cuda()
    assert [n for n, _ in model.named_parameters()] == list(g["ddi_names"])
    # (a) data-dependent init through forward hooks (Kaiming re-initialisation pinned to the seeded weights, as in the fixture)
    orig = torch.nn.init.kaiming_normal_
    torch.nn.init.kaiming_normal_ = lambda t, *a, **k: t
    try:
        with lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]]):
            ddi(model, {"x": xd})
    finally:
        torch.nn.init.kaiming_normal_ = orig
    s, l2 = _stats(model)
    ok, err = _close(l2, g["ddi_l2"], 2e-3)
    assert ok, ("ddi l2", err)
    ok, err = _close(s, g["ddi_sum"], 5e-3)
    assert ok, ("ddi sum", err)
    assert not any(m._forward_hooks for m in model.modules())          # hooks removed: the fused path is back
    # (b) one training step the way boilr's loop drives it
    class Exp(boilr.VAEExperimentManager):
        def forward_pass(self, x, y=None):
            return forward_pass(self.model, x, self.device, self.args.beta_anneal)
    exp = Exp(types.SimpleNamespace(beta_anneal=0, lr=3e-4, weight_decay=0.0, seed=0))
    exp.device = torch.device("cuda")
    exp.model = model.train()
    exp.optimizer = torch.optim.Adamax(model.parameters(), lr=3e-4, weight_decay=0.0)
    exp.optimizer.zero_grad()
    with lvae_b200.inject(eps=[e.float().cuda() for e in eps[1]]):
        out = exp.forward_pass(xd)
    for k in ("loss", "elbo", "kl", "l2", "recons"):
        assert abs(float(out[k]) - float(g["step_" + k])) < 2e-4 * abs(float(g["step_" + k])) + 1e-5, (k, float(out[k]), float(g["step_" + k]))
    assert np.allclose(out["elbo_sep"].detach().cpu().numpy(), g["step_elbo_sep"], rtol=2e-4)
    assert np.allclose(out["kl_avg_layerwise"].detach().cpu().numpy(), g["step_kl_avg_layerwise"], rtol=2e-4, atol=1e-5)
    out["loss"].backward()
    exp.optimizer.step()
    s, l2 = _stats(model)
    ok, err = _close(l2, g["stepped_l2"], 2e-3)
    assert ok, ("stepped l2", err)
    # (c) IW bound, K full forward passes in eval mode
    model.eval()
    exp.dataloaders = types.SimpleNamespace(test=[(xd, None)])
    flat = [e.float().cuda() for k in range(2, 2 + K) for e in eps[k]]
    with torch.no_grad(), lvae_b200.inject(eps=flat):
        res = exp.test_procedure(iw_samples=K)
    assert abs(res["elbo/elbo"] - float(g["iw_elbo"])) < 5e-4 * abs(float(g["iw_elbo"]))
    assert abs(res["elbo/elbo_IW_%d" % K] - float(g["iw_bound"])) < 5e-4 * abs(float(g["iw_bound"]))
