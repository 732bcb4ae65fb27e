"""GPU: BatchNorm apply fused behind a grid barrier into the convolution that produces its input (forward: conv1 +
BN2 statistics + BN2 apply; backward: conv dgrad + BatchNorm-backward sums + BatchNorm-backward apply), used when a
launch has one tile per CTA (the <= 8x8 rungs at batch 256) -- against the same model with the separate apply passes.
Same kernels up to the point of the apply and the same per-channel statistics, so the results agree to bf16 rounding of a
handful of elements (the fused backward forms x_hat as fma(x, rstd, -mean * rstd), the separate pass as (x - mean) * rstd)."""
import pytest
import torch

from oracle import lvae_oracle as O
from lvae_test_helpers import make_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg_name,batch", [("mnist3", 8), ("cifar15", 3)])
def test_fused_bn_apply_matches_separate_passes(cfg_name, batch, monkeypatch):
    import lvae_b200
    from lvae_b200 import ops
    cfg = O.baseline_config(cfg_name)
    x, eps, masks = make_inputs(cfg, batch, 31, True)
    res = []
    for on in (False, True):
        monkeypatch.setattr(ops, "_bn_post", [on])
        model = lvae_b200.LadderVAE(**cfg.kwargs())
        model.load_state_dict(O.make_params(cfg, 5), strict=True)
        model = model.cuda().train().set_compute_dtype(torch.bfloat16)
        ops.stats["bn_post_fwd"] = ops.stats["bn_post_bwd"] = 0
        with lvae_b200.inject(eps=[e.float().cuda() for e in eps[0]], masks=[m.float().cuda() for m in masks]):
            out = model(x.float().cuda())
        loss = (-out["ll"]).mean() + out["kl_loss"]
        loss.backward()
        torch.cuda.synchronize()
        assert (ops.stats["bn_post_fwd"] > 0) == on and (ops.stats["bn_post_bwd"] > 0) == on
        if on:       # every gated block at these sizes: one fused apply forward, two backward
            assert ops.stats["bn_post_bwd"] == 2 * ops.stats["bn_post_fwd"]
        res.append(dict(loss=float(loss), ll=out["ll"].detach().double().cpu(), kl=out["kl_sep"].detach().double().cpu(),
                        grads={n: p.grad.detach().double().cpu() for n, p in model.named_parameters() if p.grad is not None},
                        bufs={n: b.detach().double().cpu() for n, b in model.named_buffers()}))
    a, b = res
    assert abs(a["loss"] - b["loss"]) < 2e-5 * abs(a["loss"]), (a["loss"], b["loss"])
    assert float((a["ll"] - b["ll"]).abs().max()) < 1e-4 * float(a["ll"].abs().max())
    assert float((a["kl"] - b["kl"]).abs().max()) < 1e-3 * float(a["kl"].abs().max())
    # running statistics and num_batches_tracked are written by the fused kernel
    for n in a["bufs"]:
        assert float((a["bufs"][n] - b["bufs"][n]).abs().max()) <= 1e-5 * max(1.0, float(a["bufs"][n].abs().max())), n
    gmax = max(float(g.abs().max()) for g in a["grads"].values())
    worst = 0.0
    for n, ga in a["grads"].items():
        gb = b["grads"][n]
        if float(ga.norm()) < 1e-6 * gmax * ga.numel() ** 0.5:
            continue
        cos = float((ga * gb).sum() / (ga.norm() * gb.norm() + 1e-300))
        rel = float((ga - gb).norm() / ga.norm())
        worst = max(worst, rel)
        assert cos > 0.9995 and rel < 3e-2, (n, cos, rel)
