"""One launch each of the kernels BASELINE.json's north_star names, at the bench shapes, inside a cudaProfilerStart/Stop
range -- the target of the `ncu --set full` captures summarised under profiles/ncu_*_r02.txt:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/ncu_r02 \
        python profiles/ncu_targets.py

A gated residual block (lib/nn.py:78-126) forward + backward at B = 256, 64 channels, 16x16 and 32x32 in bf16 covers
bn_act_fwd2 / bn_act_bwd2, conv_tc_kernel<1|2>, conv_gate_tc_kernel, gate_bwd_kernel and wgrad_tc_kernel; the stochastic
block core (B = 256, 16x16, Z = 32) and the DMoL likelihood (B = 256, 32x32) are launched through ops directly.
Never a timing of record: a profiler is attached."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lvae_b200  # noqa: E402,F401
from lvae_b200 import ops  # noqa: E402
from lvae_b200.lib.nn import ResidualGatedBlock  # noqa: E402

torch.manual_seed(0)
lvae_b200.manual_seed(3)
dev = torch.device("cuda")
B = int(os.environ.get("B", "256"))
sides = [int(s) for s in os.environ.get("SIDES", "16,32").split(",")]


def block_pass(blk, x):
    ops.new_forward_epoch()
    ops.prepare_masks(2, x.shape[0], 64, 0.2, dev)
    y = blk(x)
    y.backward(torch.ones_like(y))


def stoch_pass(q, p):
    z, _, kl, kls, lp, lq = ops.stochastic_core(q, p)
    (kl.sum() + z.sum() * 0.01).backward()


def dmol_pass(l, x):
    ll = ops.dmol_loglik(l, x)
    ll.sum().backward()


work = []
for s in sides:
    blk = ResidualGatedBlock(64, "elu", batchnorm=True, block_type="bacdbacd", dropout=0.2).to(dev).train()
    x = torch.randn(B, s, s, 64, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
    work.append((block_pass, (blk, x)))
q = (torch.randn(B, 16, 16, 64, device=dev) * 0.3).permute(0, 3, 1, 2).requires_grad_(True)
p = (torch.randn(B, 16, 16, 64, device=dev) * 0.3).permute(0, 3, 1, 2).requires_grad_(True)
work.append((stoch_pass, (q, p)))
l = (torch.randn(B, 32, 32, 100, device=dev) * 0.5).permute(0, 3, 1, 2).requires_grad_(True)
xi = torch.randint(0, 256, (B, 3, 32, 32), device=dev).float() / 255.0
work.append((dmol_pass, (l, xi)))

for fn, args in work:           # warm-up: allocations, packs, function attributes
    fn(*args)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for fn, args in work:
    fn(*args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu targets done")
