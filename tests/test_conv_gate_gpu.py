"""GPU: the chained conv2 + 1x1 gate conv + gate kernel (lvae_conv_gate_tc, the tail of a gated residual block as ONE
launch; default since the round-2 A/B run: 16.7 vs 17.4 ms / step, IW 126 vs 114 images/s) against the two-launch path it
replaced (3x3 conv, then the gate conv with the fused gate epilogue) on the same inputs.  The tensors must agree bit for
bit (same bf16 operands, same MMA order); the per-channel statistics are per-CTA fp32 partial sums added into doubles, and
the two kernels cut the tiles into CTAs differently, so those agree to fp32 summation noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

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
        assert torch.allclose(acc_new.sum(0), acc_ref.sum(0), rtol=2e-6, atol=1e-3)
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
