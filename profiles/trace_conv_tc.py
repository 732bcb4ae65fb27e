"""Per-tile clock trace of CTA 0 of conv_tc_kernel (halo mode) at B=256, HxW, 3x3 64->64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200 import _capi, ops
HW = int(os.environ.get("HW", "32"))
B, C, N, k = int(os.environ.get("B", "256")), 64, 64, 3
x = torch.randn(B, HW, HW, C, device="cuda").to(torch.bfloat16)
y = torch.empty(B, HW, HW, N, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, C, k, k, device="cuda") / 24
wp = ops.WeightPack(N, C, k * k, 2).get(w, torch.bfloat16)
bias = torch.zeros(N, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def launch():
    _capi.call("lvae_conv2d_tc", x.data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None, y.data_ptr(), None, 0,
               B, HW, HW, C, N, k, 0, 0, s)
launch(); torch.cuda.synchronize()
dbg = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
_capi.lib().lvae_conv2d_tc_debug(dbg.data_ptr())
launch(); torch.cuda.synchronize()
_capi.lib().lvae_conv2d_tc_debug(None)
t = dbg.view(-1, 8).cpu()
t0 = int(t[0, 0])
print("tile  mma_ready  operands  mma_issued | epi_wait  acc_ready  epi_done | prod_free   (cycles from first stamp)")
for i in range(t.shape[0]):
    if int(t[i, 2]) == 0:
        break
    print("%3d  %9d %9d %9d | %9d %9d %9d | %9d" % ((i,) + tuple(int(v) - t0 if int(v) else -1 for v in t[i, :7])))
