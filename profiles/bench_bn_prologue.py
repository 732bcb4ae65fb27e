import os, sys
sys.path.insert(0, "/root/repo")
import torch
import lvae_b200
from lvae_b200 import _capi
B, C = 256, 64
bf = torch.bfloat16
def S(): return torch.cuda.current_stream().cuda_stream
for HW in (16, 8, 4, 2):
    P = B * HW * HW
    per = P * C * 2
    nbuf = max(2, int(400e6 // (3 * per)) + 1)
    nbuf = min(nbuf, 64)
    xs = [torch.randn(P, C, device="cuda").to(bf) for _ in range(nbuf)]
    ys = [torch.empty(P, C, device="cuda", dtype=bf) for _ in range(nbuf)]
    acc = torch.zeros(8, 2, C, dtype=torch.float64, device="cuda"); acc[0, 0] = 0.1 * P; acc[0, 1] = 1.5 * P
    save = torch.zeros(2, C, device="cuda"); save[1] = 1.0
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
    kern = {
        "bn_act_fwd2 (stats prologue)": lambda i: _capi.call("lvae_bn_act_fwd2", xs[i].data_ptr(), ys[i].data_ptr(), acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), save.data_ptr(), rm.data_ptr(), rv.data_ptr(), None, P, C, 3, 1, 0.1, 1e-5, 1, 1, S()),
        "bn_act_fwd2 eval (running stats)": lambda i: _capi.call("lvae_bn_act_fwd2", xs[i].data_ptr(), ys[i].data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), save.data_ptr(), rm.data_ptr(), rv.data_ptr(), None, P, C, 3, 0, 0.1, 1e-5, 1, 1, S()),
        "bn_act_fwd (mean/rstd given)": lambda i: _capi.call("lvae_bn_act_fwd", xs[i].data_ptr(), ys[i].data_ptr(), save[0].data_ptr(), save[1].data_ptr(), gamma.data_ptr(), beta.data_ptr(), P, C, 3, 1, 1, S()),
    }
    for name, fn in kern.items():
        for i in range(3): fn(i % nbuf)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nbuf): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("HW=%2d %-36s %6.2f us" % (HW, name, e0.elapsed_time(e1) * 1e3 / (5 * nbuf)))
