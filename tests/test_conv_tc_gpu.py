"""GPU parity of the tcgen05 / TMA / TMEM convolution kernel against a float64 convolution of the
same bf16-rounded operands.  With fp32 output the only error left is fp32 accumulation order
(tolerance 2e-5 relative); with bf16 output add one bf16 rounding (4e-3)."""
import math

import pytest
import torch
import torch.nn.functional as F

from lvae_test_helpers import rel_err

pytestmark = pytest.mark.gpu


def phys_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


CASES = [
    # cin, cout, k, H, W, B, cin2, out_fp32
    (64, 64, 3, 16, 16, 2, 0, True),
    (64, 64, 3, 16, 16, 2, 0, False),
    (64, 64, 3, 32, 32, 3, 0, True),
    (64, 64, 3, 8, 8, 5, 0, True),
    (64, 64, 3, 4, 4, 3, 0, True),       # 48 pixels: partial tile, batch dimension out of bounds in the box
    (64, 64, 3, 2, 2, 40, 0, True),      # 160 pixels: two tiles, second partial
    (64, 64, 3, 64, 64, 1, 0, True),
    (64, 64, 3, 16, 8, 3, 0, True),      # non-square halo tiles (one tile column)
    (64, 64, 3, 32, 16, 2, 0, False),
    (64, 128, 1, 16, 16, 2, 0, True),    # gate conv
    (64, 64, 1, 8, 8, 4, 64, True),      # merge conv over two inputs
    (64, 100, 3, 32, 32, 2, 0, True),    # DMoL head (N padded to 112)
    (64, 32, 3, 8, 8, 4, 0, True),
    (128, 64, 3, 8, 8, 4, 0, True),      # two k-blocks from one tensor
]


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_forward_backward(case):
    import lvae_b200
    from lvae_b200.lib.nn import Conv2d
    from lvae_b200 import ops
    cin, cout, k, H, W, B, cin2, out_fp32 = case
    g = torch.Generator().manual_seed(sum(case))
    bf = lambda t: t.to(torch.bfloat16)
    x = bf(torch.randn(B, cin, H, W, generator=g))
    x2 = bf(torch.randn(B, cin2, H, W, generator=g)) if cin2 else None
    mod = Conv2d(cin + cin2, cout, k, padding=k // 2).cuda()
    mod.spec.out_fp32 = out_fp32
    w = bf(torch.randn(mod.weight.shape, generator=g) / math.sqrt((cin + cin2) * k * k))
    b = torch.randn(cout, generator=g)
    scale = ((torch.rand(B, cout, generator=g) > 0.3).float() / 0.7)
    with torch.no_grad():
        mod.weight.copy_(w.float())
        mod.bias.copy_(b)
    # float64 reference on the same bf16-rounded operands
    xr = x.double().requires_grad_(True)
    x2r = x2.double().requires_grad_(True) if cin2 else None
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    inp = torch.cat([xr, x2r], 1) if cin2 else xr
    yr = F.conv2d(inp, wr, br, padding=k // 2) * scale.double().view(B, cout, 1, 1)
    use_res = cout == 64 and not out_fp32
    res = bf(torch.randn(B, cout, H, W, generator=g)) if use_res else None
    if use_res:
        yr = yr + res.double()
    gy = bf(torch.randn(yr.shape, generator=g))
    yr.backward(gy.double())

    xd = phys_nhwc(x.cuda()).requires_grad_(True)
    x2d = phys_nhwc(x2.cuda()).requires_grad_(True) if cin2 else None
    kw = dict(out_scale=scale.cuda())
    if cin2:
        kw["x2"] = x2d
    if use_res:
        kw["res"] = phys_nhwc(res.cuda())
    before = dict(ops.stats)
    y = mod(xd, **kw)
    assert y.dtype == (torch.float32 if out_fp32 else torch.bfloat16)
    assert rel_err(y.float(), yr) < (2e-5 if out_fp32 else 6e-3)
    y.backward(gy.cuda().to(y.dtype))
    torch.cuda.synchronize()
    assert ops.stats["tc_fwd"] == before["tc_fwd"] + 1
    if cout % 64 == 0:
        assert ops.stats["tc_dgrad"] == before["tc_dgrad"] + 1
    if cin == 64 and cout in (64, 128):
        assert ops.stats["tc_wgrad"] == before["tc_wgrad"] + 1     # tensor-core weight gradient taken
    # dgrad output is bf16: one rounding
    assert rel_err(xd.grad.float(), xr.grad) < 6e-3
    if cin2:
        assert rel_err(x2d.grad.float(), x2r.grad) < 6e-3
    # the Dropout2d-masked gradient is re-rounded to bf16 ahead of the TMA-fed dgrad / wgrad
    assert rel_err(mod.weight.grad, wr.grad) < 5e-3
    assert rel_err(mod.bias.grad, br.grad) < 5e-3


def test_tc_path_is_taken_and_matches_cuda_core_path():
    import lvae_b200
    from lvae_b200.lib.nn import Conv2d
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(0)
    mod = Conv2d(64, 64, 3, padding=1).cuda()
    x = phys_nhwc(torch.randn(4, 64, 16, 16, generator=g).to(torch.bfloat16).cuda())
    y_tc = mod(x)
    ops.set_tensor_cores(False)
    try:
        y_cc = mod(x)
    finally:
        ops.set_tensor_cores(True)
    assert rel_err(y_tc.float(), y_cc.float()) < 1e-2
    assert not torch.equal(y_tc, torch.zeros_like(y_tc))


S2_CASES = [
    # transposed, small grid H, W, batch
    (False, 8, 8, 4),
    (False, 4, 4, 8),
    (False, 2, 2, 40),      # partial last tile
    (False, 16, 8, 2),
    (True, 8, 8, 4),
    (True, 4, 4, 8),
    (True, 1, 1, 70),       # 1x1 -> 2x2
    (True, 8, 16, 3),
]


@pytest.mark.parametrize("case", S2_CASES)
def test_tc_stride2_conv_forward_backward(case):
    """Stride-2 3x3 64->64 Conv2d / ConvTranspose2d (the resampling pre_convs) on the tcgen05 kernel: TMA element strides
    for the gather side, four parity-class launches for the transposed side."""
    import lvae_b200
    from lvae_b200.lib.nn import Conv2d, ConvTranspose2d
    from lvae_b200 import ops
    transposed, hs, ws, B = case
    g = torch.Generator().manual_seed(7 + hs * 13 + ws + B + int(transposed))
    bf = lambda t: t.to(torch.bfloat16)
    H, W = (hs, ws) if transposed else (2 * hs, 2 * ws)
    x = bf(torch.randn(B, 64, H, W, generator=g))
    if transposed:
        mod = ConvTranspose2d(64, 64, 3, padding=1, stride=2, output_padding=1).cuda()
    else:
        mod = Conv2d(64, 64, 3, padding=1, stride=2).cuda()
    w = bf(torch.randn(mod.weight.shape, generator=g) / 24.0)
    b = torch.randn(64, generator=g)
    with torch.no_grad():
        mod.weight.copy_(w.float())
        mod.bias.copy_(b)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    if transposed:
        yr = F.conv_transpose2d(xr, wr, br, stride=2, padding=1, output_padding=1)
    else:
        yr = F.conv2d(xr, wr, br, stride=2, padding=1)
    gy = bf(torch.randn(yr.shape, generator=g))
    yr.backward(gy.double())
    xd = phys_nhwc(x.cuda()).requires_grad_(True)
    before = dict(ops.stats)
    y = mod(xd)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(yr.shape)
    assert ops.stats["tc_fwd"] == before["tc_fwd"] + 1, "stride-2 conv did not take the tcgen05 path"
    assert rel_err(y.float(), yr) < 6e-3
    y.backward(gy.cuda())
    torch.cuda.synchronize()
    assert ops.stats["tc_dgrad"] == before["tc_dgrad"] + 1
    assert ops.stats["tc_wgrad"] == before["tc_wgrad"] + 1, "stride-2 weight gradient did not take the tcgen05 path"
    assert rel_err(xd.grad.float(), xr.grad) < 6e-3
    assert rel_err(mod.weight.grad, wr.grad) < 5e-3
    assert rel_err(mod.bias.grad, br.grad) < 5e-3


@pytest.mark.parametrize("cout,H,W,B,out_fp32", [(1, 28, 28, 3, True), (2, 7, 5, 4, True), (4, 8, 8, 2, False), (1, 3, 3, 1, True)])
def test_narrow_head_conv(cout, H, W, B, out_fp32):
    """64 -> N <= 4 3x3 conv (Bernoulli parameter_net) takes the bandwidth-bound narrow kernel in the bf16 pipeline."""
    import lvae_b200
    from lvae_b200.lib.nn import Conv2d
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(cout * 100 + H)
    x = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16)
    mod = Conv2d(64, cout, 3, padding=1).cuda()
    mod.spec.out_fp32 = out_fp32
    xr = x.double().requires_grad_(True)
    wr, br = mod.weight.detach().double().cpu().requires_grad_(True), mod.bias.detach().double().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=1)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy.double())
    xd = phys_nhwc(x.cuda()).requires_grad_(True)
    n0 = ops.stats.get("narrow_fwd", 0)
    y = mod(xd)
    assert ops.stats.get("narrow_fwd", 0) == n0 + 1
    assert y.dtype == (torch.float32 if out_fp32 else torch.bfloat16)
    assert rel_err(y.float(), yr) < (1e-5 if out_fp32 else 6e-3)
    y.backward(gy.cuda().to(y.dtype))
    torch.cuda.synchronize()
    assert rel_err(xd.grad.float(), xr.grad) < 8e-3
    assert rel_err(mod.weight.grad, wr.grad) < 8e-3


@pytest.mark.parametrize("B,Hs,Ws,h,w", [(5, 32, 32, 28, 28), (3, 16, 16, 13, 9), (2, 32, 32, 32, 30), (65, 32, 32, 28, 28)])
def test_narrow_head_conv_reads_a_cropped_view_in_place(B, Hs, Ws, h, w):
    """Under no_grad ops.crop returns a strided view of the NHWC buffer and the 64 -> 1 head conv (row-segment kernel) reads the
    window in place: same result as the copying crop followed by the conv on a dense tensor, and as the fp64 reference; the
    zero padding is that of the WINDOW, not of the surrounding tensor."""
    import lvae_b200  # noqa: F401
    from lvae_b200.lib.nn import Conv2d
    from lvae_b200 import ops
    g = torch.Generator().manual_seed(B + Hs + h)
    x = torch.randn(B, 64, Hs, Ws, generator=g).to(torch.bfloat16)
    mod = Conv2d(64, 1, 3, padding=1).cuda()
    mod.spec.out_fp32 = True
    xd = phys_nhwc(x.cuda())
    y0, x0 = (Hs - h) // 2, (Ws - w) // 2
    ref = F.conv2d(x.double()[:, :, y0:y0 + h, x0:x0 + w], mod.weight.detach().double().cpu(), mod.bias.detach().double().cpu(), padding=1)
    with torch.no_grad():
        v = ops.crop(xd, (h, w))
        assert v.data_ptr() != xd.data_ptr() or (y0 == 0 and x0 == 0)
        assert v.permute(0, 2, 3, 1).untyped_storage().data_ptr() == xd.permute(0, 2, 3, 1).untyped_storage().data_ptr()   # a view
        y_view = mod(v)
        dense = phys_nhwc(x[:, :, y0:y0 + h, x0:x0 + w].contiguous().cuda())
        y_dense = mod(dense)
    torch.cuda.synchronize()
    assert torch.equal(y_view, y_dense)
    assert rel_err(y_view.float(), ref) < 1e-5
    # with grad enabled the crop is the copying autograd node
    c = ops.crop(xd.requires_grad_(True), (h, w))
    assert c.permute(0, 2, 3, 1).is_contiguous()
