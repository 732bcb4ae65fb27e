"""Per-tile clock64 trace of CTA 0 of conv_gate_tc_kernel in eval mode (the IW evaluator's shape: B = 1000, 16x16 by default).
Epilogue warp 2: top of the tile, 3x3 accumulator ready, c2 packed, c2 staged, gate accumulator ready, gate evaluated, store
issued; MMA warp: tile start, operands landed, gate GEMM issued."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200  # noqa: F401
from lvae_b200 import _capi, ops
HW, B = int(os.environ.get("HW", "16")), int(os.environ.get("B", "1000"))
bf = torch.bfloat16
a2 = torch.randn(B, HW, HW, 64, device="cuda").to(bf)
xres = torch.randn(B, HW, HW, 64, device="cuda").to(bf)
w2 = torch.randn(64, 64, 3, 3, device="cuda") / 24
wg = torch.randn(128, 64, 1, 1, device="cuda") / 8
b2, bg = torch.randn(64, device="cuda"), torch.randn(128, device="cuda")
w2p = ops.WeightPack(64, 64, 9, 2).get(w2, bf)
wgp = ops.WeightPack(128, 64, 1, 2).get(wg, bf)
def launch():
    return ops._conv_gate_chain(a2, w2p, b2, None, wgp, bg, xres, 3, None, False)
for _ in range(3):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    launch()
e1.record()
torch.cuda.synchronize()
print("conv_gate eval B=%d %dx%d: %.1f us per launch (eager, back to back)" % (B, HW, HW, e0.elapsed_time(e1) * 50))
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
_capi.lib().lvae_conv_gate_tc_debug(dbg.data_ptr())
launch(); torch.cuda.synchronize()
_capi.lib().lvae_conv_gate_tc_debug(None)
t = dbg.view(-1, 16).cpu()
t0 = int(t[0, 8])
print("tile | epi: top  acc1_rdy c2_packed c2_staged acc2_rdy gate_done stored | mma: start operands gate_issued   (cycles)")
for i in range(t.shape[0]):
    if int(t[i, 0]) == 0:
        break
    v = [int(x) - t0 if int(x) else -1 for x in t[i]]
    print("%3d | %8d %8d %8d %8d %8d %8d %8d | %8d %8d %8d" % (i, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[8], v[9], v[10]))
