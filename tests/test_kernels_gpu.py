"""GPU parity, kernel by kernel: every C-ABI op against a float64 CPU restatement
(torch functional ops / oracle functions) on the same seeded inputs.  Tolerances are for fp32
kernels: 2e-5 relative on outputs, 1e-4 on gradients (north_star asks 1e-4 / 1e-3)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lvae_test_helpers import rel_err

pytestmark = pytest.mark.gpu

TOL, GTOL = 2e-5, 1e-4


@pytest.fixture(scope="module")
def L():
    import lvae_b200
    lvae_b200._capi.device_check()
    return lvae_b200


def dev(t, requires_grad=False):
    return t.float().cuda().requires_grad_(requires_grad)


def to_nhwc_phys(t):
    """NCHW tensor whose memory is NHWC (what our modules exchange)."""
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


CONV_CASES = [
    # cin, cout, k, stride, pad, transposed, H, W, B, cin2
    (64, 64, 3, 1, 1, False, 8, 8, 3, 0),
    (64, 64, 3, 2, 1, False, 8, 8, 2, 0),
    (64, 64, 3, 2, 1, True, 4, 4, 2, 0),
    (64, 128, 1, 1, 0, False, 4, 4, 2, 0),
    (64, 64, 1, 1, 0, False, 4, 6, 2, 64),
    (1, 64, 5, 2, 2, False, 32, 32, 2, 0),
    (3, 16, 5, 1, 2, False, 9, 7, 2, 0),
    (32, 64, 3, 1, 1, False, 2, 2, 5, 0),
    (64, 100, 3, 1, 1, False, 6, 6, 2, 0),
    (64, 1, 3, 1, 1, False, 28, 28, 2, 0),
    (16, 16, 3, 1, 1, False, 5, 5, 3, 0),
    (6, 8, 3, 1, 1, False, 5, 5, 3, 0),
    (8, 8, 3, 2, 1, True, 3, 5, 2, 0),
    (64, 64, 3, 2, 1, False, 16, 16, 2, 0),   # dgrad: parity-class-major transposed gather (tap skipping)
    (64, 64, 3, 2, 1, True, 8, 8, 2, 0),      # forward: the same path
    (32, 32, 3, 2, 1, True, 8, 16, 3, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_bwd(L, case):
    cin, cout, k, stride, pad, transposed, H, W, B, cin2 = case
    from lvae_b200.lib.nn import Conv2d, ConvTranspose2d
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
    x2 = torch.randn(B, cin2, H, W, generator=g, dtype=torch.float64) if cin2 else None
    if transposed:
        mod = ConvTranspose2d(cin, cout, k, stride=stride, padding=pad, output_padding=1)
    else:
        mod = Conv2d(cin + cin2, cout, k, stride=stride, padding=pad)
    w = torch.randn(mod.weight.shape, generator=g, dtype=torch.float64) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    scale = (torch.rand(B, cout, generator=g, dtype=torch.float64) > 0.3).double() / 0.7
    # float64 reference
    xr = x.clone().requires_grad_(True)
    x2r = x2.clone().requires_grad_(True) if cin2 else None
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    inp = torch.cat([xr, x2r], 1) if cin2 else xr
    if transposed:
        yr = F.conv_transpose2d(inp, wr, br, stride=stride, padding=pad, output_padding=1)
    else:
        yr = F.conv2d(inp, wr, br, stride=stride, padding=pad)
    res = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    resr = res.clone().requires_grad_(True)
    outr = yr * scale.view(B, cout, 1, 1) + resr
    gy = torch.randn(outr.shape, generator=g, dtype=torch.float64)
    outr.backward(gy)
    # ours
    mod = mod.cuda()
    with torch.no_grad():
        mod.weight.copy_(w.float())
        mod.bias.copy_(b.float())
    xd, resd = dev(x, True), dev(res, True)
    x2d = dev(x2, True) if cin2 else None
    kw = dict(out_scale=dev(scale), res=resd)
    if cin2:
        kw["x2"] = x2d
    out = mod(xd, **kw)
    assert tuple(out.shape) == tuple(outr.shape)
    assert rel_err(out, outr) < TOL
    out.backward(dev(gy))
    assert rel_err(xd.grad, xr.grad) < GTOL
    if cin2:
        assert rel_err(x2d.grad, x2r.grad) < GTOL
    assert rel_err(mod.weight.grad, wr.grad) < GTOL
    assert rel_err(mod.bias.grad, br.grad) < GTOL
    assert rel_err(resd.grad, resr.grad) < 1e-6


@pytest.mark.parametrize("C,act,training", [(64, "elu", True), (64, "elu", False), (16, "relu", True),
                                            (8, "selu", True), (24, "leakyrelu", True), (64, None, True)])
def test_bn_act(L, C, act, training):
    from lvae_b200.lib.nn import BatchNorm2d
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(C)
    B, H, W = 5, 6, 7
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 2 + 0.5
    gamma = 1 + 0.2 * torch.randn(C, generator=g, dtype=torch.float64)
    beta = 0.3 * torch.randn(C, generator=g, dtype=torch.float64)
    rm = 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
    rv = 1 + 0.3 * torch.rand(C, generator=g, dtype=torch.float64)
    fn = {None: lambda t: t, "elu": F.elu, "relu": F.relu, "selu": F.selu, "leakyrelu": F.leaky_relu}[act]
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rmr, rvr = rm.clone(), rv.clone()
    yr = fn(F.batch_norm(xr, rmr, rvr, gr, br, training, 0.1, 1e-5))
    gy = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    yr.backward(gy)
    bn = BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(gamma.float()); bn.bias.copy_(beta.float())
        bn.running_mean.copy_(rm.float()); bn.running_var.copy_(rv.float())
    bn.train(training)
    xd = dev(x, True)
    y = ops.bn_act(xd, bn, ops.ACT_IDS[act])
    assert rel_err(y, yr) < TOL
    y.backward(dev(gy))
    assert rel_err(xd.grad, xr.grad) < GTOL
    assert rel_err(bn.weight.grad, gr.grad) < GTOL
    assert rel_err(bn.bias.grad, br.grad) < GTOL
    assert rel_err(bn.running_mean, rmr) < TOL and rel_err(bn.running_var, rvr) < TOL
    assert int(bn.num_batches_tracked) == (1 if training else 0)
    # second call: the self-cleaning accumulators must be back to zero
    y2 = ops.bn_act(dev(x), bn, ops.ACT_IDS[act])
    assert rel_err(y2, yr) < (TOL if training else TOL)


def test_plain_activation_and_gate(L):
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 16, 5, 4, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = F.elu(xr)
    gy = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    yr.backward(gy)
    xd = dev(x, True)
    y = ops.bn_act(xd, None, 3)
    y.backward(dev(gy))
    assert rel_err(y, yr) < TOL and rel_err(xd.grad, xr.grad) < GTOL
    h = torch.randn(3, 32, 5, 4, generator=g, dtype=torch.float64)
    r = torch.randn(3, 16, 5, 4, generator=g, dtype=torch.float64)
    hr, rr = h.clone().requires_grad_(True), r.clone().requires_grad_(True)
    a, b = hr.chunk(2, 1)
    outr = F.elu(a) * torch.sigmoid(b) + rr
    outr.backward(gy)
    hd, rd = dev(h, True), dev(r, True)
    out = ops.gate(hd, rd, 3)
    out.backward(dev(gy))
    assert rel_err(out, outr) < TOL and rel_err(hd.grad, hr.grad) < GTOL and rel_err(rd.grad, rr.grad) < 1e-6


def test_upsample_crop_pad(L):
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 8, 5, 3, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    gy = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    yr.backward(gy)
    xd = dev(x, True)
    y = ops.upsample2x(xd)
    y.backward(dev(gy))
    assert rel_err(y, yr) < TOL and rel_err(xd.grad, xr.grad) < GTOL
    img = torch.rand(2, 3, 28, 26, generator=g)
    p = ops.pad_image(img.cuda(), (32, 32))
    assert rel_err(p, F.pad(img, [3, 3, 2, 2])) == 0   # img is float32 already
    act = torch.randn(2, 8, 32, 32, generator=g, dtype=torch.float64)
    ar = act.clone().requires_grad_(True)
    cr = ar[:, :, 2:30, 3:29]
    gc = torch.randn(cr.shape, generator=g, dtype=torch.float64)
    cr.backward(gc)
    ad = dev(act, True)
    c = ops.crop(ad, (28, 26))
    c.backward(dev(gc))
    assert rel_err(c, cr.float()) == 0 and rel_err(ad.grad, ar.grad.float()) == 0


@pytest.mark.parametrize("cin,H,W,B", [(3, 32, 32, 5), (1, 32, 32, 3), (3, 64, 64, 2), (3, 28, 20, 3), (1, 7, 9, 2), (3, 36, 70, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_conv_kernels(L, cin, H, W, B, dtype):
    """The stem (models/lvae.py:75: Conv2d c -> 64, 5x5, stride 2, padding 2) has its own forward and weight-gradient kernels
    (16x16 / 8x16 output tiles with the input patch in shared memory): image sizes that are not multiples of the tiles, several
    tiles per image, both channel counts, both activation types, against F.conv2d in float64."""
    from lvae_b200 import _capi
    from lvae_b200.lib.nn import Conv2d
    g = torch.Generator().manual_seed(cin * 1000 + H * 10 + W)
    x = torch.randn(B, cin, H, W, generator=g)
    mod = Conv2d(cin, 64, 5, stride=2, padding=2)
    w = torch.randn(mod.weight.shape, generator=g) / math.sqrt(cin * 25)
    b = torch.randn(64, generator=g)
    xq = x.to(dtype)
    wq = w.to(dtype).double() if dtype == torch.bfloat16 else w.double()      # the bf16 path rounds the packed weights
    wr, br = wq.clone().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.conv2d(xq.double(), wr, br, stride=2, padding=2)
    gy = torch.randn(yr.shape, generator=g).to(dtype)
    yr.backward(gy.double())
    mod = mod.cuda()
    with torch.no_grad():
        mod.weight.copy_(w)
        mod.bias.copy_(b)
    n0 = _capi.launch_count()
    xd = xq.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    out = mod(xd)
    assert out.dtype == dtype and tuple(out.shape) == tuple(yr.shape)
    out.backward(gy.cuda())
    tol, gtol = (TOL, GTOL) if dtype == torch.float32 else (6e-3, 2e-3)   # bf16: output rounding; gradients accumulate in fp32
    assert rel_err(out, yr) < tol
    assert rel_err(mod.weight.grad, wr.grad) < gtol
    assert rel_err(mod.bias.grad, br.grad) < gtol
    assert _capi.launch_count() > n0


@pytest.mark.parametrize("B,C,H,W", [(3, 64, 16, 16), (2, 64, 5, 7), (2, 8, 1, 1), (1, 16, 2, 9), (2, 4, 4, 4)])
def test_upsample_bf16_matches_rounded_exact_result(L, B, C, H, W):
    """bf16 activations: the x2 bilinear weights are exact binary fractions, so forward and backward must equal the fp64 result
    rounded to bf16 (up to a tie: one bf16 ulp on a handful of elements).  C % 8 == 0 takes the 16-byte kernels (closed-form
    backward stencil), C = 4 the generic ones: both are checked against the same reference, including the clamped borders."""
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + H * 10 + W)
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    gy = torch.randn(B, C, 2 * H, 2 * W, generator=g).bfloat16()
    xr = x.double().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    yr.backward(gy.double())
    xd = x.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
    y = ops.upsample2x(xd)
    y.backward(gy.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    for ours, ref in ((y, yr.detach()), (xd.grad, xr.grad)):
        assert ours.dtype == torch.bfloat16
        want = ref.bfloat16().double()
        d = (ours.double().cpu() - want).abs()
        ulp = want.abs().clamp(min=2.0 ** -120) * 2.0 ** -7            # >= one bf16 ulp of the reference value
        assert bool((d <= ulp).all()), float((d / ulp).max())
        assert float((d > 0).double().mean()) < 5e-3


@pytest.mark.parametrize("Z,hw,analytical,broadcast", [(32, (8, 8), False, False), (32, (2, 2), False, True),
                                                       (32, (4, 4), True, False), (6, (3, 5), False, False),
                                                       (8, (4, 4), False, True), (64, (4, 4), False, False),
                                                       # several CTAs per sample (deterministic cross-CTA finish)
                                                       (32, (32, 32), False, False), (32, (16, 16), True, True),
                                                       (6, (24, 20), False, False)])
def test_stochastic_core(L, Z, hw, analytical, broadcast):
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(Z + hw[0])
    B = 5
    q = torch.randn(B, 2 * Z, *hw, generator=g, dtype=torch.float64)
    p = torch.randn(1 if broadcast else B, 2 * Z, *hw, generator=g, dtype=torch.float64)
    eps = torch.randn(B, Z, *hw, generator=g, dtype=torch.float64)
    qr, pr = q.clone().requires_grad_(True), p.clone().requires_grad_(True)
    qm, ql = qr.chunk(2, 1)
    pm, pl = pr.chunk(2, 1)
    z = qm + (ql / 2).exp() * eps
    logp = O.normal_log_prob(z, pm, pl).sum((1, 2, 3))
    logq = O.normal_log_prob(z, qm, ql).sum((1, 2, 3))
    kl_an = O.normal_kl(qm, ql, pm, pl)
    kl = kl_an.sum((1, 2, 3)) if analytical else logq - logp
    kls = kl_an.sum(1)
    w = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in (z, kl, kls, logp, logq)]
    (z * w[0]).sum().add((kl * w[1]).sum()).add((kls * w[2]).sum()).add((logp * w[3]).sum()).add((logq * w[4]).sum()).backward()
    qd, pd = dev(q, True), dev(p, True)
    zz, _, klo, klso, lpo, lqo = ops.stochastic_core(qd, pd, eps=dev(eps), analytical=analytical)
    assert rel_err(zz, z) < TOL and rel_err(klo, kl) < TOL and rel_err(klso, kls) < TOL
    assert rel_err(lpo, logp) < TOL and rel_err(lqo, logq) < TOL
    ((zz * dev(w[0])).sum() + (klo * dev(w[1])).sum() + (klso * dev(w[2])).sum() + (lpo * dev(w[3])).sum()
     + (lqo * dev(w[4])).sum()).backward()
    assert rel_err(qd.grad, qr.grad) < GTOL
    assert rel_err(pd.grad, pr.grad) < GTOL


def test_stochastic_core_other_z_kinds(L):
    """Mode (z = mu_q), forced latent and sampling from the prior (q absent) against the fp64 closed forms."""
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, Z, hw = 4, 32, (16, 16)
    q = torch.randn(B, 2 * Z, *hw, generator=g, dtype=torch.float64)
    p = torch.randn(B, 2 * Z, *hw, generator=g, dtype=torch.float64)
    qm, ql = q.chunk(2, 1)
    pm, pl = p.chunk(2, 1)
    forced = torch.randn(B, Z, *hw, generator=g, dtype=torch.float64)
    for kind, zref in (("mode", qm), ("forced", forced)):
        zz, _, klo, klso, lpo, lqo = ops.stochastic_core(dev(q), dev(p), forced=dev(forced) if kind == "forced" else None,
                                                         use_mode=kind == "mode")
        logp = O.normal_log_prob(zref, pm, pl).sum((1, 2, 3))
        logq = O.normal_log_prob(zref, qm, ql).sum((1, 2, 3))
        assert rel_err(zz, zref) < TOL and rel_err(lpo, logp) < TOL and rel_err(lqo, logq) < TOL, kind
        assert rel_err(klo, logq - logp) < TOL and rel_err(klso, O.normal_kl(qm, ql, pm, pl).sum(1)) < TOL, kind
    eps = torch.randn(B, Z, *hw, generator=g, dtype=torch.float64)
    zz, _, klo, klso, lpo, lqo = ops.stochastic_core(None, dev(p), eps=dev(eps))
    zref = pm + (pl / 2).exp() * eps
    assert klo is None and klso is None and lqo is None
    assert rel_err(zz, zref) < TOL and rel_err(lpo, O.normal_log_prob(zref, pm, pl).sum((1, 2, 3))) < TOL


@pytest.mark.parametrize("Lr,B,fb", [(15, 256, 1.0), (3, 7, 0.5), (12, 1000, 0.0), (20, 64, 2.0)])
def test_kl_bookkeeping(L, Lr, B, fb):
    """lvae_kl_bookkeeping (+ backward) against the oracle's free_bits_kl and the sums of models/lvae.py:192-198."""
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(Lr * 1000 + B)
    kl = (torch.rand(Lr, B, generator=g, dtype=torch.float64) * 2 * max(fb, 0.5)).requires_grad_(True)     # about half below the clamp
    lp = torch.randn(Lr, B, generator=g, dtype=torch.float64).requires_grad_(True)
    klbl = kl.t()                                           # (batch, layers), the reference's layout
    ref = {"kl_sep": klbl.sum(1), "kl": klbl.sum(1).mean(), "kl_avg_layerwise": klbl.mean(0),
           "kl_loss": O.free_bits_kl(klbl, fb).sum(), "logp": lp.mean(1).sum()}
    w = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) for k, v in ref.items()}
    sum((ref[k] * w[k]).sum() for k in ref).backward()
    for as_rows in (True, False):
        if as_rows:                                         # rows of one (3,L,B) matrix, as LadderVAE.topdown_pass hands them over
            rows = torch.zeros(3, Lr, B, device="cuda")
            rows[0].copy_(kl.detach())
            rows[1].copy_(lp.detach())
            rows.requires_grad_(True)
            kls, lps = [rows[0, i] for i in range(Lr)], [rows[1, i] for i in range(Lr)]
            out = ops.kl_bookkeeping(kls, lps, fb, rows=rows.detach())
        else:                                               # unrelated vectors: gathered first
            kd, ld = dev(kl.detach(), True), dev(lp.detach(), True)
            out = ops.kl_bookkeeping([kd[i] for i in range(Lr)], [ld[i] for i in range(Lr)], fb)
        for k in ref:
            assert rel_err(out[k], ref[k]) < 1e-5, (k, as_rows)
        sum((out[k] * dev(w[k])).sum() for k in ref).backward()
        gk, gl = (rows.grad[0], rows.grad[1]) if as_rows else (kd.grad, ld.grad)
        assert rel_err(gk, kl.grad) < 1e-5 and rel_err(gl, lp.grad) < 1e-5, as_rows


@pytest.mark.parametrize("B,hw,broadcast", [(5, (16, 16), False), (3, (8, 8), False), (2, (32, 32), False), (4, (4, 4), False),
                                            (3, (2, 2), True)])
def test_stochastic_core_with_philox_noise(L, B, hw, broadcast):
    """The training path draws eps inside the kernel (Philox), so it cannot be fed the oracle's eps: instead the noise is
    recovered from the sample, eps = (z - mu_q) / sigma_q, and every other output (log p, log q, Monte-Carlo KL, analytic KL
    per pixel, the bf16 copy of z) is checked against the fp64 closed forms for THAT noise.  Covers one and several CTAs per sample and the batch-broadcast prior."""
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    Z = 32
    g = torch.Generator().manual_seed(B * 100 + hw[0])
    q = torch.randn(B, 2 * Z, *hw, generator=g, dtype=torch.float64) * 0.7
    p = torch.randn(1 if broadcast else B, 2 * Z, *hw, generator=g, dtype=torch.float64) * 0.7
    L.manual_seed(21)
    zz, zlp, klo, klso, lpo, lqo = ops.stochastic_core(dev(q), dev(p), lowp_copy=True)
    qm, ql = q.chunk(2, 1)
    pm, pl = p.chunk(2, 1)
    z = zz.double().cpu()
    eps = (z - qm) / (ql / 2).exp()
    if eps.numel() >= 20000:
        assert abs(float(eps.mean())) < 0.03 and abs(float(eps.std()) - 1) < 0.03   # it IS standard normal noise
    logp = O.normal_log_prob(z, pm, pl).sum((1, 2, 3))
    logq = O.normal_log_prob(z, qm, ql).sum((1, 2, 3))
    assert rel_err(lpo, logp) < TOL and rel_err(lqo, logq) < 5 * TOL      # log q from the recovered eps: (z - mu) / sigma rounding
    assert rel_err(klo, logq - logp) < 2e-4
    assert rel_err(klso, O.normal_kl(qm, ql, pm, pl).sum(1)) < TOL
    assert tuple(zlp.shape) == (B, 64) + tuple(hw)
    assert float((zlp[:, :Z].double().cpu() - z).abs().max()) <= 2 ** -8 * float(z.abs().max()) + 1e-6     # bf16 rounding of z
    assert float(zlp[:, Z:].abs().max()) == 0.0                                                            # zero-padded channels


def test_stochastic_philox_statistics(L):
    from lvae_b200 import ops
    L.manual_seed(7)
    q = torch.zeros(64, 64, 16, 16, device="cuda")
    p = torch.zeros(64, 64, 16, 16, device="cuda")
    z1 = ops.stochastic_core(q, p)[0]
    z2 = ops.stochastic_core(q, p)[0]
    assert abs(float(z1.mean())) < 5e-3 and abs(float(z1.std()) - 1) < 5e-3
    assert float((z1 - z2).abs().max()) > 1          # fresh draws per call
    k = float((z1 ** 4).mean())
    assert abs(k - 3) < 0.1                           # normal kurtosis
    L.manual_seed(7)
    z3 = ops.stochastic_core(q, p)[0]
    assert torch.equal(z1, z3)                        # reproducible from the seed


def test_bernoulli(L):
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 4, 1, 28, 28
    logits = torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 8
    logits[0, 0, 0, :4] = torch.tensor([20.0, -20.0, 30.0, -120.0])          # saturation (SURVEY trap 2)
    x = (torch.rand(B, C, H, W, generator=g) < 0.3).double()
    lr = logits.float().requires_grad_(True)                                   # fp32: saturation is dtype-specific
    llr = O.bernoulli_log_lik(x.float(), torch.sigmoid(lr))
    gll = torch.randn(B, generator=g)
    llr.backward(gll)
    ld = dev(logits, True)
    prob, ll = ops.bernoulli_loglik(ld, x.float().cuda())
    assert rel_err(ll, llr) < TOL
    assert rel_err(prob, torch.sigmoid(lr)) < TOL
    ll.backward(gll.cuda())
    assert rel_err(ld.grad, lr.grad) < GTOL
    s = ops.bernoulli_sample(torch.full((8, 1, 64, 64), 0.3, device="cuda"))
    assert abs(float(s.mean()) - 0.3) < 0.01
    # 3 colour channels: NHWC params vs NCHW image indexing
    logits3 = torch.randn(2, 3, 5, 4, generator=g, dtype=torch.float64)
    x3 = (torch.rand(2, 3, 5, 4, generator=g) < 0.5).double()
    ll3 = ops.bernoulli_loglik(dev(logits3), x3.float().cuda())[1]
    assert rel_err(ll3, O.bernoulli_log_lik(x3, torch.sigmoid(logits3))) < TOL


@pytest.mark.parametrize("H,W,B", [(32, 32, 3), (16, 16, 2), (5, 7, 2)])
def test_dmol(L, H, W, B):
    from oracle import lvae_oracle as O
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(H)
    l = torch.randn(B, 100, H, W, generator=g, dtype=torch.float64)
    l[:, 20:30] -= 4            # some log-scales below the -7 clamp / narrow logistics
    l[:, 50:60] *= 3
    x = torch.randint(0, 256, (B, 3, H, W), generator=g).double() / 255
    x[0, :, 0, 0] = 0.0
    x[0, :, 0, 1] = 1.0
    lr = l.clone().requires_grad_(True)
    llr = O.dmol_log_lik(x, lr)
    gll = torch.randn(B, generator=g, dtype=torch.float64)
    llr.backward(gll)
    ld = dev(l, True)
    ll = ops.dmol_loglik(ld, x.float().cuda())
    assert rel_err(ll, llr) < TOL
    ll.backward(dev(gll))
    assert rel_err(ld.grad, lr.grad) < 1e-3       # north_star gradient tolerance; inputs here are extreme
    s = ops.dmol_sample(dev(l))
    assert tuple(s.shape) == (B, 3, H, W) and float(s.min()) >= 0 and float(s.max()) <= 1


def test_dmol_sampler_distribution(L):
    from lvae_b200 import ops
    # one dominant component with a tight scale: samples must concentrate at its mean
    l = torch.zeros(4, 100, 8, 8)
    l[:, 3] = 20.0
    for c in range(3):
        l[:, 10 + 30 * c + 3] = (-0.5, 0.0, 0.5)[c]
        l[:, 10 + 30 * c + 10 + 3] = -6.0
    s = ops.dmol_sample(l.cuda()) * 2 - 1
    m = s.mean((0, 2, 3)).cpu()
    assert torch.allclose(m, torch.tensor([-0.5, 0.0, 0.5]), atol=0.01)


def test_adamax_l2_iw(L):
    from oracle import lvae_oracle as O
    from lvae_b200 import ops, _capi
    g = torch.Generator().manual_seed(9)
    n = 100003
    p = torch.randn(n, generator=g, dtype=torch.float64)
    pr = p.clone()
    ea, ei = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    pd = dev(p)
    m, u = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros((), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for t in range(1, 4):
        gr = torch.randn(n, generator=g, dtype=torch.float64)
        O.adamax_step(pr, gr, ea, ei, t)
        gd = dev(gr)
        _capi.call("lvae_adamax_step", pd.data_ptr(), gd.data_ptr(), m.data_ptr(), u.data_ptr(), n, 3e-4, 0.9, 0.999,
                   1e-8, 0.0, step.data_ptr(), 1.0, None, s)
    assert int(step) == 3
    assert rel_err(pd, pr) < 1e-6
    acc = torch.zeros((), dtype=torch.float64, device="cuda")
    out = torch.zeros((), device="cuda")
    _capi.call("lvae_l2_norm", pd.data_ptr(), n, acc.data_ptr(), out.data_ptr(), s)
    assert rel_err(out, pr.pow(2).sum().sqrt()) < 1e-6 and float(acc) == 0.0
    # the fused form: one more step, with the norm of the updated parameters from the same pass
    gr = torch.randn(n, generator=g, dtype=torch.float64)
    O.adamax_step(pr, gr, ea, ei, 4)
    _capi.call("lvae_adamax_step_l2", pd.data_ptr(), dev(gr).data_ptr(), m.data_ptr(), u.data_ptr(), n, 3e-4, 0.9, 0.999,
               1e-8, 0.0, step.data_ptr(), 1.0, None, acc.data_ptr(), out.data_ptr(), s)
    assert int(step) == 4 and rel_err(pd, pr) < 1e-6
    assert rel_err(out, pr.pow(2).sum().sqrt()) < 1e-6 and float(acc) == 0.0
    # IW bound: streaming logsumexp over K samples split over R "ranks"
    B, K, R = 37, 12, 3
    ll = torch.randn(K, B, generator=g, dtype=torch.float64) * 50 - 800
    kl = torch.rand(K, B, generator=g, dtype=torch.float64) * 40
    ref = O.iw_bound((ll - kl).t())
    states = torch.zeros(R, B, 2, device="cuda")
    for r in range(R):
        for j, k in enumerate(range(r, K, R)):
            ops.iw_lse_update(dev(ll[k]), dev(kl[k]), states[r], j == 0)
    out = ops.iw_lse_combine(states, K)
    assert rel_err(out, ref) < 1e-5


def test_dropout_masks(L):
    from lvae_b200 import ops
    L.manual_seed(1)
    ops.prepare_masks(10, 64, 64, 0.2, torch.device("cuda", 0))
    m = [ops.next_mask(64, 64, 0.2, torch.device("cuda", 0)) for _ in range(10)]
    ops.clear_masks()
    allm = torch.stack(m)
    vals = set(np.round(allm.unique().cpu().numpy(), 5).tolist())
    assert vals == {0.0, 1.25}
    assert abs(float((allm > 0).float().mean()) - 0.8) < 0.01
    assert not torch.equal(m[0], m[1])


def test_bad_arguments_raise(L):
    from lvae_b200 import _capi
    with pytest.raises(RuntimeError, match="bn_stats"):
        _capi.call("lvae_bn_stats", None, None, 0, 64, 0, None)
    with pytest.raises(RuntimeError, match="ldw"):
        t = torch.zeros(16, device="cuda")
        _capi.call("lvae_conv2d_gather", t.data_ptr(), None, t.data_ptr(), None, None, None, None, t.data_ptr(),
                   1, 1, 1, 4, 0, 1, 1, 3, 3, 1, 1, 1, 0, 0, 0, None)
