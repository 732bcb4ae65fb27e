"""Launch floor of the tcgen05 conv kernel: 1x1 (one tap, 4 MMAs) against 3x3 (nine taps, 36 MMAs) on tiny tensors, graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200  # noqa: F401
from lvae_b200 import _capi, ops
B, C = 256, 64
for HW in (2, 4, 8):
    for k, N in ((1, 64), (3, 64)):
        nbuf = 16
        xs = [torch.randn(B, HW, HW, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty(B, HW, HW, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        w = torch.randn(N, C, k, k, device="cuda") / 24
        wp = ops.WeightPack(N, C, k * k, 2).get(w, torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        def launch(i):
            _capi.call("lvae_conv2d_tc", xs[i % nbuf].data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None, ys[i % nbuf].data_ptr(),
                       None, 0, B, HW, HW, C, N, k, 0, 0, torch.cuda.current_stream().cuda_stream)
        for i in range(3): launch(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(64): launch(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("HW=%d k=%d N=%d: %.2f us per launch" % (HW, k, N, e0.elapsed_time(e1) * 1e3 / 320))
