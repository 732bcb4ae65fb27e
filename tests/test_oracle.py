"""CPU: the oracle against the golden vectors written from the unmodified reference, and --
where /root/reference is mounted -- against the reference itself."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import lvae_oracle as O
from oracle import ref_loader
from lvae_test_helpers import golden_names, load_golden, make_inputs, rel_err

FAST = ["mnist3_train_b4", "mnist3_eval_b4", "small_dmol_train_b4", "small_dmol_eval_b4", "small_bern_bacdbac",
        "small_bern_cabdcabd_linear", "small_dmol_nobn_selu", "cifar15_train_b2"]
# (mnist12 / celeba20 golden files are exercised by the GPU parity tests; here they would only add minutes)


def run_oracle(cfg, meta, dtype):
    P = O.make_params(cfg, meta["weight_seed"], dtype)
    x, eps, masks = make_inputs(cfg, meta["batch"], meta["input_seed"], meta["training"], meta["n_iw"])
    x = x.to(dtype)
    cast = lambda lst: None if lst is None else [t.to(dtype) for t in lst]
    if meta["training"]:
        st = O.TrainState(cfg, P)
        out, terms = st.step(x, cast(eps[0]), cast(masks), update=False)
    else:
        with torch.no_grad():
            out = O.forward(P, cfg, x, cast(eps[0]), None, False)
            terms = O.loss_terms(out)
    return P, out, terms, x, eps


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_golden_f64(name):
    cfg, meta, g = load_golden(name)
    P, out, terms, x, eps = run_oracle(cfg, meta, torch.float64)
    assert rel_err(out["ll"], g["f64_ll"]) < 1e-10
    assert rel_err(out["kl_sep"], g["f64_kl_sep"]) < 1e-10
    assert rel_err(out["kl_avg_layerwise"], g["f64_kl_avg_layerwise"]) < 1e-10
    assert rel_err(out["kl_loss"], g["f64_kl_loss"]) < 1e-10
    assert rel_err(terms["loss"], g["f64_loss"]) < 1e-10
    assert rel_err([z.sum().item() for z in out["z"]], g["f64_z_sum"], 1e-6) < 1e-8
    assert rel_err([k.sum().item() for k in out["kl_spatial"]], g["f64_kl_spatial_sum"]) < 1e-10
    if meta["training"]:
        names = [str(n) for n in g["f64_grad_names"]]
        gl2 = np.array([P[n].grad.pow(2).sum().sqrt().item() if P[n].grad is not None else 0.0 for n in names])
        scale = g["f64_grad_l2"].max()
        assert np.abs(gl2 - g["f64_grad_l2"]).max() < 1e-8 * scale
        rn = [str(n) for n in g["f64_running_names"]]
        if rn:
            rs = np.array([P[n].double().sum().item() for n in rn])
            assert np.abs(rs - g["f64_running_sum"]).max() < 1e-9 * np.abs(g["f64_running_sum"]).max()
    if meta["n_iw"]:
        cols = []
        with torch.no_grad():
            for k in range(meta["n_iw"]):
                o = O.forward(P, cfg, x, [e for e in eps[k]], None, False)
                cols.append(o["ll"] - o["kl_sep"])
        iw = O.iw_bound(torch.stack(cols, 1))
        assert rel_err(iw, g["f64_iw_bound"]) < 1e-10


def test_oracle_f32_close_to_f64_golden():
    cfg, meta, g = load_golden("small_dmol_train_b4")
    P, out, terms, _, _ = run_oracle(cfg, meta, torch.float32)
    assert rel_err(out["ll"], g["f64_ll"]) < 1e-4
    assert rel_err(terms["loss"], g["f64_loss"]) < 1e-4


def test_state_dict_layout_counts():
    # parameter counts of the four BASELINE models
    expect = {"mnist3": 3209153, "mnist12": 11607809, "cifar15": 14467684, "celeba20": 19207460}
    for name, n in expect.items():
        cfg = O.baseline_config(name)
        shapes = O.param_shapes(cfg)
        tot = sum(int(np.prod(shapes[k])) for k in O.trainable_names(cfg))
        assert tot == n, (name, tot)


def test_iw_bound_and_adamax_restatements():
    torch.manual_seed(0)
    m = torch.randn(5, 7, dtype=torch.float64) * 30
    ref = torch.log(torch.exp(m - m.max(1, keepdim=True)[0]).sum(1)) + m.max(1)[0] - np.log(7)
    assert rel_err(O.iw_bound(m), ref) < 1e-12
    p = torch.randn(50, dtype=torch.float64, requires_grad=True)
    p2 = p.detach().clone()
    opt = torch.optim.Adamax([p], lr=3e-4)
    ea, ei = torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    for step in range(1, 4):
        g = torch.randn(50, dtype=torch.float64)
        p.grad = g.clone()
        opt.step()
        O.adamax_step(p2, g, ea, ei, step)
    assert rel_err(p2, p.detach()) < 1e-12


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["small_dmol_train_b4", "small_bern_bacdbac", "small_bern_cabdcabd_linear"])
def test_oracle_matches_live_reference(name):
    cfg, meta, g = load_golden(name)
    ref = ref_loader.load_reference()
    dtype = torch.float64
    model = ref["lvae"].LadderVAE(**cfg.kwargs()).to(dtype)
    P = O.make_params(cfg, meta["weight_seed"], dtype)
    assert list(model.state_dict().keys()) == list(P.keys())
    model.load_state_dict(P)
    model.train(meta["training"])
    x, eps, masks = make_inputs(cfg, meta["batch"], meta["input_seed"], meta["training"])
    mctx = ref_loader.DropoutMaskQueue(masks) if masks else contextlib.nullcontext()
    with ref_loader.EpsQueue(eps[0]), mctx:
        out = model(x)
    loss = (-out["ll"]).mean() + out["kl_loss"]
    loss.backward()
    st = O.TrainState(cfg, {k: v.clone() for k, v in P.items()})
    o2, t2 = st.step(x, eps[0], masks, update=False)
    assert rel_err(o2["ll"], out["ll"].detach()) < 1e-12
    assert rel_err(t2["loss"].detach(), loss.detach()) < 1e-12
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert float((st.P[n].grad - p.grad).abs().max()) < 1e-9 * gmax, n
