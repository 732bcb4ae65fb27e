"""GPU tests of the opt-in kernels that were written when no GPU minutes were left (so they are NOT part of the default
`-m gpu` run: set LVAE_TEST_EXPERIMENTAL=1).  Each compares an opt-in kernel with the validated default path on the same
inputs; the results must agree bit for bit (same bf16 operands, same MMA order) or to fp32 accumulation noise.

    LVAE_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LVAE_TEST_EXPERIMENTAL", "0") == "0",
                                 reason="opt-in kernels: set LVAE_TEST_EXPERIMENTAL=1")]

# B, H, W, training (c2 / h stored + Dropout2d mask + statistics), activation id (3 = ELU)
CHAIN_CASES = [(2, 16, 16, True, 3), (3, 32, 32, True, 3), (5, 8, 8, True, 3), (3, 4, 4, True, 3), (40, 2, 2, True, 3),
               (2, 16, 16, False, 3), (4, 8, 8, False, 1), (2, 64, 64, True, 3), (3, 16, 8, True, 3)]


@pytest.mark.parametrize("case", CHAIN_CASES)
def test_conv_gate_chain_matches_two_launches(case):
    """lvae_conv_gate_tc (conv2 + 1x1 gate conv + gate in one launch) against conv2 -> gate conv with the fused gate epilogue."""
    import lvae_b200  # noqa: F401
    from lvae_b200 import ops
    B, H, W, train, act = case
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    bf = torch.bfloat16
    a2 = torch.randn(B, H, W, 64, generator=g).to(bf).cuda()
    xres = torch.randn(B, H, W, 64, generator=g).to(bf).cuda()
    w2 = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    wg = (torch.randn(128, 64, 1, 1, generator=g) / 8).cuda()
    b2, bg = torch.randn(64, generator=g).cuda(), torch.randn(128, generator=g).cuda()
    m2 = ((torch.rand(B, 64, generator=g) > 0.2).float() / 0.8).cuda() if train else None
    w2p = ops.WeightPack(64, 64, 9, 2).get(w2, bf)
    wgp = ops.WeightPack(128, 64, 1, 2).get(wg, bf)
    acc_ref = torch.zeros(8, 2, 64, dtype=torch.float64, device="cuda")
    acc_new = torch.zeros_like(acc_ref)
    # validated path: two launches
    ops._gate_keep_h[0] = train
    c2_ref = ops._conv_tc(a2, None, w2p, b2, m2, None, 64, 3, False, False)
    h_ref, out_ref = ops._conv_tc(c2_ref, None, wgp, bg, None, None, 128, 1, False, False,
                                  stats_acc=acc_ref if train else None, gate=(xres, act))
    # one launch
    c2, h, out = ops._conv_gate_chain(a2, w2p, b2, m2, wgp, bg, xres, act, acc_new if train else None, train)
    torch.cuda.synchronize()
    ops._gate_keep_h[0] = True
    assert torch.equal(out, out_ref)
    if train:
        assert torch.equal(c2, c2_ref) and torch.equal(h, h_ref)
        assert torch.allclose(acc_new.sum(0), acc_ref.sum(0), rtol=1e-12, atol=1e-9)
    else:
        assert c2 is None and h is None


def test_conv_gate_chain_in_the_model(monkeypatch):
    """The whole training step with the chained kernel switched on gives the same loss and gradients as without."""
    import lvae_b200
    from lvae_b200 import ops
    from oracle import lvae_oracle as O
    from oracle.make_golden import make_inputs, small_cfg
    cfg = small_cfg(n_filters=64, z_dims=[32, 32, 32], dropout=0.2)
    x, eps, masks = make_inputs(cfg, 4, 5, True)
    results = []
    for chain in (False, True):
        monkeypatch.setattr(ops, "_gate_chain", [chain])
        model = lvae_b200.LadderVAE(**cfg.kwargs())
        model.load_state_dict(O.make_params(cfg, 3))
        model = model.cuda().train().set_compute_dtype(torch.bfloat16)
        ops.stats["gate_chain"] = 0
        with lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]], masks=[m.float().cuda() for m in masks]):
            out = model(x.float().cuda())
        loss = (-out["ll"]).mean() + out["kl_loss"]
        loss.backward()
        torch.cuda.synchronize()
        assert (ops.stats.get("gate_chain", 0) > 0) == chain
        results.append((float(loss), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = results
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    for n in g0:
        assert torch.allclose(g0[n], g1[n], rtol=1e-4, atol=1e-6), n


BWD_CASES = [(2, 16, 16, True, 3), (3, 32, 32, True, 3), (5, 8, 8, False, 3), (3, 4, 4, True, 3), (40, 2, 2, True, 1),
             (1, 2, 2, True, 3), (7, 4, 2, False, 4)]


@pytest.mark.parametrize("case", BWD_CASES)
def test_gate_bwd_dgrad_chain_matches_two_launches(case):
    """lvae_gate_bwd_dgrad_tc against lvae_gate_bwd followed by the tcgen05 1x1 data gradient (mask in its epilogue)."""
    import lvae_b200  # noqa: F401
    from lvae_b200 import ops
    B, H, W, masked, act = case
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    bf = torch.bfloat16
    gout = torch.randn(B, H, W, 64, generator=g).to(bf).cuda()
    h = torch.randn(B, H, W, 128, generator=g).to(bf).cuda()
    wg = (torch.randn(128, 64, 1, 1, generator=g) / 8).cuda()
    m2 = ((torch.rand(B, 64, generator=g) > 0.2).float() / 0.8).cuda() if masked else None
    wpb = ops.WeightPack(128, 64, 1, 3).get(wg, bf)
    dh_ref = torch.empty_like(h)
    ops.call("lvae_gate_bwd", gout.data_ptr(), h.data_ptr(), dh_ref.data_ptr(), B * H * W, 64, act, 1, ops._stream())
    if ops._pow2(H) and ops._pow2(W):
        dc2_ref = ops._conv_tc(dh_ref, None, wpb, None, m2, None, 64, 1, True, False)
    else:
        dc2_ref = None
    dh, dc2 = ops._gate_bwd_dgrad_chain(gout, h, wpb, m2, act)
    torch.cuda.synchronize()
    assert torch.equal(dh, dh_ref)
    if dc2_ref is not None:
        assert torch.equal(dc2, dc2_ref)
    # float64 reference of the data gradient from the bf16-rounded dh
    ref = torch.einsum("bhwk,kc->bhwc", dh_ref.double().cpu(), wg.to(bf).double().cpu().view(128, 64))
    if m2 is not None:
        ref = ref * m2.double().cpu().view(B, 1, 1, 64)
    err = (dc2.double().cpu() - ref).abs().max() / ref.abs().max()
    assert float(err) < 6e-3


def test_gate_chains_in_the_model_backward(monkeypatch):
    """Both chained kernels on: same loss and gradients as the default path."""
    import lvae_b200
    from lvae_b200 import ops
    from oracle import lvae_oracle as O
    from oracle.make_golden import make_inputs, small_cfg
    cfg = small_cfg(n_filters=64, z_dims=[32, 32, 32], dropout=0.2)
    x, eps, masks = make_inputs(cfg, 4, 5, True)
    results = []
    for on in (False, True):
        monkeypatch.setattr(ops, "_gate_chain", [on])
        monkeypatch.setattr(ops, "_gate_bwd_chain", [on])
        model = lvae_b200.LadderVAE(**cfg.kwargs())
        model.load_state_dict(O.make_params(cfg, 3))
        model = model.cuda().train().set_compute_dtype(torch.bfloat16)
        ops.stats["gate_bwd_chain"] = 0
        with lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]], masks=[m.float().cuda() for m in masks]):
            out = model(x.float().cuda())
        loss = (-out["ll"]).mean() + out["kl_loss"]
        loss.backward()
        torch.cuda.synchronize()
        assert (ops.stats.get("gate_bwd_chain", 0) > 0) == on
        results.append((float(loss), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = results
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    for n in g0:
        assert torch.allclose(g0[n], g1[n], rtol=1e-4, atol=1e-6), n
