"""Device-time microbenchmark of the HBM-bound passes (CUDA-graph replay, rotating buffers > L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200 import _capi

B, C = 256, 64
bf = torch.bfloat16
def S():
    return torch.cuda.current_stream().cuda_stream
for HW in (32, 16, 8, 4):
    P = B * HW * HW
    per = P * C * 2
    nbuf = max(2, int(400e6 // (3 * per)) + 1)
    xs = [torch.randn(P, C, device="cuda").to(bf) for _ in range(nbuf)]
    ys = [torch.empty(P, C, device="cuda", dtype=bf) for _ in range(nbuf)]
    ds = [torch.randn(P, C, device="cuda").to(bf) for _ in range(nbuf)]
    hs = [torch.randn(P, 2 * C, device="cuda").to(bf) for _ in range(nbuf)]
    dhs = [torch.empty(P, 2 * C, device="cuda", dtype=bf) for _ in range(nbuf)]
    acc = torch.zeros(8, 2, C, dtype=torch.float64, device="cuda")      # 8-way striped accumulators
    acc[0, 0] = 0.1 * P; acc[0, 1] = 1.5 * P
    save = torch.zeros(2, C, device="cuda"); save[1] = 1.0
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    kern = {
        "bn_act_fwd2 (r+w)": (2 * per, lambda i: _capi.call("lvae_bn_act_fwd2", xs[i].data_ptr(), ys[i].data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), save.data_ptr(), rm.data_ptr(), rv.data_ptr(), None, P, C, 3, 1, 0.1, 1e-5, 1, 1, S())),
        "bn_act_bwd2 apply (2r+w)": (3 * per, lambda i: _capi.call("lvae_bn_act_bwd2", ds[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), save.data_ptr(), gamma.data_ptr(), beta.data_ptr(), acc.data_ptr(), dg.data_ptr(), db.data_ptr(), None, None, P, HW * HW, C, 3, 1, 1, 1, S())),
        "bn_act_bwd2 reduce+apply (4r+w)": (5 * per, lambda i: _capi.call("lvae_bn_act_bwd2", ds[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), save.data_ptr(), gamma.data_ptr(), beta.data_ptr(), acc.data_ptr(), dg.data_ptr(), db.data_ptr(), None, None, P, HW * HW, C, 3, 1, 1, 0, S())),
        "bn_stats (r)": (per, lambda i: _capi.call("lvae_bn_stats", xs[i].data_ptr(), acc.data_ptr(), P, C, 1, S())),
        "gate_fwd_stats (3r+w)": (4 * per, lambda i: _capi.call("lvae_gate_fwd_stats", hs[i].data_ptr(), xs[i].data_ptr(), ys[i].data_ptr(), acc.data_ptr(), P, C, 3, 1, S())),
        "gate_bwd (3r+2w)": (5 * per, lambda i: _capi.call("lvae_gate_bwd", ds[i].data_ptr(), hs[i].data_ptr(), dhs[i].data_ptr(), P, C, 3, 1, S())),
    }
    for name, (nbytes, fn) in kern.items():
        fn(0); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n = 30
        with torch.cuda.graph(g):
            for i in range(n):
                fn(i % nbuf)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print("HW=%2d %-34s %7.2f us  %7.0f GB/s" % (HW, name, us, nbytes / us / 1e3))
