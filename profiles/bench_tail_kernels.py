"""Device-time microbenchmark of the single-launch "tail" kernels of a step (CUDA-graph replay, rotating buffers): the stem
convolution and its weight gradient, the generic BatchNorm-backward apply of the stem's ungated block, the backward of the
x2 up-sampling, the batched weight re-pack, and the Bernoulli head of the IW evaluator (64 -> 1 narrow conv on a cropped view).

    python profiles/bench_tail_kernels.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200 import _capi, ops
from lvae_b200.engine import PackTable
from lvae_b200.configs import baseline_config

bf = torch.bfloat16


def S():
    return torch.cuda.current_stream().cuda_stream


def timeit(name, fn, nbuf, nbytes=None, n=20):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    extra = "  %7.0f GB/s" % (nbytes / us / 1e3) if nbytes else ""
    print("%-58s %8.2f us%s" % (name, us, extra), flush=True)


NB = 6
# --- stem: Conv2d(3 -> 64, 5x5, s2, p2) on the (256,32,32,3) CIFAR batch, bf16 --------------------------------------------
B, Hi, Ho = 256, 32, 16
conv = lvae_b200.lib.nn.Conv2d(3, 64, 5, stride=2, padding=2).cuda()
xs = [torch.randn(B, Hi, Hi, 3, device="cuda").to(bf) for _ in range(NB)]
ys = [torch.empty(B, Ho, Ho, 64, device="cuda", dtype=bf) for _ in range(NB)]
gys = [torch.randn(B, Ho, Ho, 64, device="cuda").to(bf) for _ in range(NB)]
wp = conv.spec.pack_fwd.get(conv.weight, bf)
ld = conv.spec.pack_fwd.ld
gw = torch.zeros_like(conv.weight)
gb = torch.zeros_like(conv.bias)
timeit("stem conv forward  (256,32,32,3) -> (256,16,16,64) bf16",
       lambda i: _capi.call("lvae_conv2d_gather", xs[i].data_ptr(), None, wp.data_ptr(), conv.bias.data_ptr(), None, None, None,
                            ys[i].data_ptr(), B, Hi, Hi, 3, 0, Ho, Ho, 64, ld, 5, 5, 2, 2, 0, 1, S()), NB)
timeit("stem conv weight gradient (same shape)",
       lambda i: _capi.call("lvae_conv2d_wgrad", xs[i].data_ptr(), None, gys[i].data_ptr(), None, None, gw.data_ptr(), gb.data_ptr(),
                            B, Hi, Hi, 3, 0, Ho, Ho, 64, 5, 5, 2, 2, 1, S()), NB)

# --- generic BatchNorm-backward (reduce + apply + params) of the stem's ungated block: (256,16,16,64) bf16 --------------
P, C = B * Ho * Ho, 64
x2 = [torch.randn(P, C, device="cuda").to(bf) for _ in range(NB)]
d2 = [torch.randn(P, C, device="cuda").to(bf) for _ in range(NB)]
o2 = [torch.empty(P, C, device="cuda", dtype=bf) for _ in range(NB)]
mean, rstd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
acc = torch.zeros(8 * 2 * C, dtype=torch.float64, device="cuda")
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
timeit("bn_act_bwd generic (reduce + apply + params), 16x16",
       lambda i: _capi.call("lvae_bn_act_bwd", d2[i].data_ptr(), x2[i].data_ptr(), o2[i].data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                            gamma.data_ptr(), beta.data_ptr(), acc.data_ptr(), dg.data_ptr(), db.data_ptr(), P, C, 3, 1, 1, S()), NB,
       nbytes=5 * P * C * 2)

# --- backward of the x2 bilinear up-sampling: (256,32,32,64) -> (256,16,16,64) bf16 ----------------------------------------
gu = [torch.randn(B, 32, 32, 64, device="cuda").to(bf) for _ in range(NB)]
du = [torch.empty(B, 16, 16, 64, device="cuda", dtype=bf) for _ in range(NB)]
timeit("upsample2x_bwd (256,32,32,64) -> (256,16,16,64) bf16",
       lambda i: _capi.call("lvae_upsample2x_bwd", gu[i].data_ptr(), du[i].data_ptr(), B, 16, 16, 64, 1, S()), NB,
       nbytes=(B * 32 * 32 * 64 + B * 16 * 16 * 64) * 2)
uu = [torch.randn(B, 16, 16, 64, device="cuda").to(bf) for _ in range(NB)]
timeit("upsample2x_fwd (256,16,16,64) -> (256,32,32,64) bf16",
       lambda i: _capi.call("lvae_upsample2x_fwd", uu[i].data_ptr(), gu[i].data_ptr(), B, 16, 16, 64, 1, S()), NB,
       nbytes=(B * 32 * 32 * 64 + B * 16 * 16 * 64) * 2)

# --- the batched weight re-pack of the CIFAR-15 model ------------------------------------------------------------------
cfg = baseline_config("cifar15")
model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
model.set_compute_dtype(bf)
packs = PackTable(model, bf)
timeit("pack_weights, CIFAR-15 model (%d descriptors)" % packs.n, lambda i: packs.repack(), 1)

# --- Bernoulli head of the IW evaluator: 3x3 64 -> 1 on the centred 28x28 crop of a (1000,32,32,64) tensor --------------
Bi = 1000
hx = [torch.randn(Bi, 32, 32, 64, device="cuda").to(bf) for _ in range(3)]
hw_ = torch.randn(1, 64, 3, 3, device="cuda")
hb = torch.zeros(1, device="cuda")
hy = torch.empty(Bi, 28, 28, 1, device="cuda")
timeit("conv3x3_narrow1 (1000,28,28,64) window -> 1, fp32 out",
       lambda i: _capi.call("lvae_conv3x3_narrow_ex", hx[i][:, 2:, 2:].data_ptr(), hw_.data_ptr(), hb.data_ptr(), hy.data_ptr(), Bi, 28, 28, 1,
                            1, 32 * 64, 32 * 32 * 64, S()), 3, nbytes=Bi * 28 * 28 * 64 * 2)

# --- batch sum of the top prior's gradient: (256, 2x2x64) fp32 -> (2x2x64) ----------------------------------------------
dpb = [torch.randn(256, 256, device="cuda") for _ in range(3)]
dps = torch.empty(256, device="cuda")
timeit("sum_batch (256, 256) -> (256)", lambda i: _capi.call("lvae_sum_batch", dpb[i].data_ptr(), dps.data_ptr(), 256, 256, 0, S()), 3)
