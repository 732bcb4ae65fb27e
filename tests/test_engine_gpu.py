"""GPU: the step engines (flat arenas, fused Adamax, CUDA-graph replay, IW evaluator)."""
import numpy as np
import pytest
import torch

from oracle import lvae_oracle as O
from lvae_test_helpers import load_golden, make_inputs, rel_err

pytestmark = pytest.mark.gpu


def build(cfg, seed):
    import lvae_b200
    model = lvae_b200.LadderVAE(**cfg.kwargs())
    model.load_state_dict(O.make_params(cfg, seed), strict=True)
    return model.cuda()


def test_eager_engine_matches_oracle_training_steps():
    """Two full optimisation steps (grads into the arena, fused Adamax, weight re-pack) against the
    CPU oracle's torch.optim.Adamax steps on the same eps / dropout masks."""
    import lvae_b200
    from lvae_b200.engine import TrainEngine
    cfg, meta, g = load_golden("small_dmol_train_b4")
    cfg.dropout = 0.25
    B = meta["batch"]
    model = build(cfg, 21)
    eng = TrainEngine(model, B, use_graph=False)
    st = O.TrainState(cfg, O.make_params(cfg, 21, torch.float64))
    hist = []
    for step in range(2):
        x, eps, masks = make_inputs(cfg, B, 300 + step, True)
        with lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]], masks=[m.float().cuda() for m in masks]):
            out = eng.step(x.float().cuda())
        o2, t2 = st.step(x, eps[0], masks)
        hist.append({n: (st.P[n].grad.clone(), dict(model.named_parameters())[n].grad.double().cpu().clone())
                     for n in st.names if st.P[n].grad is not None})
        assert rel_err(out["loss"], t2["loss"]) < 1e-4
        assert rel_err(out["elbo"], t2["elbo"]) < 1e-4
    assert int(eng.step_count) == 2
    # Adamax normalises by max|g|, so entries whose true gradient is zero (conv biases in front of a
    # train-mode BatchNorm) move by +-lr on rounding noise: compare only entries with a real gradient.
    gmax = max(float(s["exp_inf"].max()) for s in st.opt.state.values())
    worst, info = 0.0, None
    for n, p in model.named_parameters():
        if p.requires_grad:
            sel = torch.stack([h[n][0].abs() > 1e-4 * gmax for h in hist]).all(0)   # real gradient in EVERY step
            if sel.any():
                d = (p.detach().double().cpu() - st.P[n].detach()).abs() * sel
                if float(d.max()) > worst:
                    worst = float(d.max())
                    i = int(d.flatten().argmax())
                    info = (n, i, [(float(h[n][0].flatten()[i]), float(h[n][1].flatten()[i])) for h in hist], gmax)
    assert worst < 2e-5, (worst, info)             # each step moves an entry by up to lr = 3e-4
    l2 = O.l2_norm(st.P, [n for n in st.names if st.P[n].requires_grad]).item()
    assert rel_err(out["l2"], l2) < 1e-5
    # eager forward after engine steps must see the updated weights (pack cache invalidation)
    x, eps, _ = make_inputs(cfg, B, 999, False)
    model.eval()
    with torch.no_grad(), lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]]):
        o = model(x.float().cuda())
    with torch.no_grad():
        o2 = O.forward({k: v.detach() for k, v in st.P.items()}, cfg, x, eps[0], None, False)
    assert rel_err(o["ll"], o2["ll"]) < 1e-4


def test_graph_engine_runs_and_draws_fresh_noise():
    from lvae_b200.engine import TrainEngine
    import lvae_b200
    lvae_b200.manual_seed(5)
    cfg = O.baseline_config("mnist3")
    model = build(cfg, 22)
    eng = TrainEngine(model, 8, use_graph=True)
    x, _, _ = make_inputs(cfg, 8, 1, True)
    xd = x.float().cuda()
    p0 = eng.arena.flat.clone()
    losses = []
    for _ in range(6):
        losses.append(float(eng.step(xd)["loss"]))
    assert all(np.isfinite(losses))
    assert len(set(losses)) == len(losses)                 # new eps / dropout masks on every replay
    assert int(eng.step_count) == 6 + 2                      # + 2 warm-up steps before capture
    assert float((eng.arena.flat - p0).abs().max()) > 1e-4
    assert eng.launches_per_step > 100
    assert losses[-1] < losses[0] + 50


def test_iw_evaluator():
    from lvae_b200.engine import IWEvaluator
    import lvae_b200
    lvae_b200.manual_seed(6)
    cfg = O.baseline_config("mnist3")
    model = build(cfg, 23)
    x, _, _ = make_inputs(cfg, 16, 2, False)
    xd = x.float().cuda()
    ev = IWEvaluator(model, 16, use_graph=True)
    b1 = ev.bound(xd, 1)
    b64 = ev.bound(xd, 64)
    ev2 = IWEvaluator(model, 16, use_graph=False)
    c64 = ev2.bound(xd, 64)
    assert tuple(b64.shape) == (16,) and torch.isfinite(b64).all()
    assert float(b64.mean()) > float(b1.mean()) - 1.0        # the bound tightens with K (up to noise)
    assert abs(float(b64.mean()) - float(c64.mean())) < 0.05 * abs(float(c64.mean()))
    # reference-style loop (bottom-up pass recomputed per sample) agrees statistically
    ev3 = IWEvaluator(model, 16, use_graph=False, reuse_bottomup=False)
    d64 = ev3.bound(xd, 64)
    assert abs(float(b64.mean()) - float(d64.mean())) < 0.05 * abs(float(d64.mean()))
