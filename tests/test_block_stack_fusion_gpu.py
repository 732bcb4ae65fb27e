"""GPU: cross-block fusion of the backward pass (ops._pending_bn1 / lvae_bn_act_bwd2_gate).  Inside a stack of directly
adjacent gated residual blocks the BatchNorm1-backward apply of block k and the gate backward of block k-1 run as ONE
elementwise kernel.  Same arithmetic on the same rounded values: results must be bit-identical to the two launches."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P_hw_act", [(6 * 16 * 16, 256, 3), (3 * 8 * 8, 64, 3), (33 * 2 * 2, 4, 3), (5 * 4 * 4, 16, 1)])
def test_fused_kernel_is_bit_identical_to_the_two_launches(P_hw_act):
    import lvae_b200  # noqa: F401
    from lvae_b200 import ops
    from lvae_b200._capi import call
    P, hw, act = P_hw_act
    C = 64
    g = torch.Generator().manual_seed(P + act)
    bf = torch.bfloat16
    dy = torch.randn(P, C, generator=g).to(bf).cuda()
    x = torch.randn(P, C, generator=g).to(bf).cuda()
    add = torch.randn(P, C, generator=g).to(bf).cuda()
    h = torch.randn(P, 2 * C, generator=g).to(bf).cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    save = torch.stack([torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5]).cuda()
    acc = torch.zeros(8, 2, C, dtype=torch.float64)
    acc[:, 0] = torch.randn(8, C, generator=g).double() * P / 80
    acc[:, 1] = torch.randn(8, C, generator=g).double() * P / 80
    acc = acc.cuda()
    s = ops._stream()
    # two launches
    dx_ref, dh_ref = torch.empty_like(x), torch.empty_like(h)
    dg_ref, db_ref = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    call("lvae_bn_act_bwd2", dy.data_ptr(), x.data_ptr(), dx_ref.data_ptr(), save.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
         acc.data_ptr(), dg_ref.data_ptr(), db_ref.data_ptr(), None, add.data_ptr(), P, hw, C, act, 1, 1, 1, s)
    call("lvae_gate_bwd", dx_ref.data_ptr(), h.data_ptr(), dh_ref.data_ptr(), P, C, act, 1, s)
    # one launch
    dx, dh = torch.empty_like(x), torch.empty_like(h)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    call("lvae_bn_act_bwd2_gate", dy.data_ptr(), x.data_ptr(), dx.data_ptr(), save.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
         acc.data_ptr(), dg.data_ptr(), db.data_ptr(), None, add.data_ptr(), h.data_ptr(), dh.data_ptr(), P, hw, C, act, act, 1, s)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_ref) and torch.equal(dh, dh_ref)
    assert torch.equal(dg, dg_ref) and torch.equal(db, db_ref)


def test_training_step_with_and_without_block_stack_fusion(monkeypatch):
    """Engine step (gradient sinks, eager launches) with the hand-over on and off: same loss, same gradient arena bit for bit
    up to the weight-gradient reduce-add order; the hand-over actually happens; nothing is left pending."""
    import lvae_b200
    from lvae_b200 import ops
    from lvae_b200.configs import baseline_config
    from lvae_b200.engine import TrainEngine
    from bench import synthetic_batch
    cfg = baseline_config("mnist3")
    x = synthetic_batch(cfg, 8, 0).cuda()
    res = []
    for on in (False, True):
        monkeypatch.setattr(ops, "_stack_fusion", [on])
        torch.manual_seed(7)
        lvae_b200.manual_seed(11)
        model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
        model.set_compute_dtype(torch.bfloat16)
        eng = TrainEngine(model, 8, use_graph=False, wgrad_side_stream=False)
        ops.stats["bn1_gate_fused"] = 0
        losses = [float(eng.step(x)["loss"]) for _ in range(2)]
        torch.cuda.synchronize()
        assert (ops.stats.get("bn1_gate_fused", 0) > 0) == on
        assert not ops._pending_bn1
        res.append((losses, eng.arena.flat.detach().clone()))
    (l0, p0), (l1, p1) = res
    assert abs(l0[0] - l1[0]) <= 1e-6 * abs(l0[0])
    assert abs(l0[1] - l1[1]) <= 5e-5 * abs(l0[1])          # second step: after one update (atomics order in the reductions)
    assert (p0 - p1).abs().max().item() <= 2e-3              # Adamax moves every entry by <= lr per step


def test_plain_autograd_without_sinks_never_defers():
    """Without an engine (no gradient sinks) the hand-over must stay off: gamma / beta gradients are returned to autograd."""
    import lvae_b200
    from lvae_b200 import ops
    from lvae_b200.configs import baseline_config
    from bench import synthetic_batch
    cfg = baseline_config("mnist3")
    torch.manual_seed(3)
    model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda().train()
    model.set_compute_dtype(torch.bfloat16)
    x = synthetic_batch(cfg, 4, 0).cuda()
    ops.stats["bn1_gate_fused"] = 0
    out = model(x)
    ((-out["ll"]).mean() + out["kl_loss"]).backward()
    torch.cuda.synchronize()
    assert ops.stats.get("bn1_gate_fused", 0) == 0 and not ops._pending_bn1
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)
