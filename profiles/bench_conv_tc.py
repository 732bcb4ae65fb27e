"""Device-time microbenchmark of the tcgen05 conv kernel per shape (CUDA-graph replay, so no host
launch or tensor-map encode cost is inside the timed region; rotating buffers > L2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lvae_b200  # noqa: E402
from lvae_b200 import _capi, ops  # noqa: E402

B, C = 256, 64
one = len(sys.argv) > 1 and sys.argv[1] == "--one"
wgrad = "--wgrad" in sys.argv
fuse_mode = 2 if "--bnb" in sys.argv else (1 if "--stats" in sys.argv else 0)   # fused epilogue reductions (N = 64 only)
import ctypes  # noqa: E402
shapes = [(int(os.environ.get("HW", "16")), 3, 64)] if one else [(32, 3, 64), (16, 3, 64), (8, 3, 64), (4, 3, 64), (16, 1, 128)] if "--wgrad" in sys.argv else [(32, 3, 64), (16, 3, 64), (8, 3, 64), (4, 3, 64), (2, 3, 64), (16, 1, 128), (32, 3, 100)]
s = torch.cuda.current_stream()
for HW, k, N in shapes:
    f32_out = N == 100                       # the DMoL head writes fp32 likelihood parameters (as in the model)
    per = B * HW * HW * (C * 2 + N * (4 if f32_out else 2))
    nbuf = max(2, int(300e6 // per) + 1)
    xs = [torch.randn(B, HW, HW, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    ys = [torch.empty(B, HW, HW, N, device="cuda", dtype=torch.float32 if f32_out else torch.bfloat16) for _ in range(nbuf)]
    w = torch.randn(N, C, k, k, device="cuda") / 24
    wp = ops.WeightPack(N, C, k * k, 2).get(w, torch.bfloat16)
    bias = torch.zeros(N, device="cuda")

    dw = torch.zeros(N, C, k, k, device="cuda")
    db = torch.zeros(N, device="cuda")
    ws = torch.empty(148 * 512 * 128, device="cuda")

    def launch_wgrad(i):
        _capi.call("lvae_conv2d_wgrad_tc", xs[i % nbuf].data_ptr(), None, ys[i % nbuf].data_ptr(), dw.data_ptr(), db.data_ptr(),
                   ws.data_ptr(), B, HW, HW, N, k, 0, 0, 0, 0, torch.cuda.current_stream().cuda_stream)

    acc = torch.zeros(8, 2, 64, dtype=torch.float64, device="cuda")
    save = torch.cat([torch.zeros(64), torch.ones(64)]).cuda()
    gamma, beta = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    fz = _capi.ConvFuse()
    if fuse_mode == 1:
        fz.stats_acc = acc.data_ptr()
    elif fuse_mode == 2:
        fz.bnb_save, fz.bnb_gamma, fz.bnb_beta, fz.bnb_acc, fz.bnb_act = save.data_ptr(), gamma.data_ptr(), beta.data_ptr(), acc.data_ptr(), 3

    def launch_fused(i):
        if fuse_mode == 2:
            fz.bnb_x = xs[(i + 1) % nbuf].data_ptr()
        _capi.call("lvae_conv2d_tc_ex", xs[i % nbuf].data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None,
                   ys[i % nbuf].data_ptr(), None, 0, B, HW, HW, C, N, k, 0, 0, ctypes.addressof(fz), torch.cuda.current_stream().cuda_stream)

    def launch(i):
        if wgrad:
            return launch_wgrad(i)
        if fuse_mode and N == 64:
            return launch_fused(i)
        _capi.call("lvae_conv2d_tc", xs[i % nbuf].data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None,
                   ys[i % nbuf].data_ptr(), None, 0, B, HW, HW, C, N, k, 0, 1 if f32_out else 0, torch.cuda.current_stream().cuda_stream)
    launch(0)
    torch.cuda.synchronize()
    if one:
        torch.cuda.profiler.start()
        for i in range(3):
            launch(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    n = 40
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 2.0 * B * HW * HW * C * N * k * k
    print("H=W=%2d k=%d N=%3d: %7.2f us/launch  %7.1f TFLOP/s  %6.0f GB/s (in+out)  tiles/CTA %.2f" % (
        HW, k, N, us, fl / us / 1e6, per / us / 1e3, B * HW * HW / 128 / 148))
