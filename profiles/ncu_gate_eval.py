"""One eval-mode (no_grad) gated residual block forward at the IW evaluator's shape (B = 1000, 64 channels, 16x16, bf16) inside a
cudaProfilerStart/Stop range: target of an `ncu --set full --import-source on -k regex:conv_gate_tc_kernel` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lvae_b200  # noqa: E402,F401
from lvae_b200 import ops  # noqa: E402
from lvae_b200.lib.nn import ResidualGatedBlock  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda")
B, s = int(os.environ.get("B", "1000")), int(os.environ.get("SIDE", "16"))
blk = ResidualGatedBlock(64, "elu", batchnorm=True, block_type="bacdbacd", dropout=0.2).to(dev).eval()
x = torch.randn(B, s, s, 64, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
with torch.no_grad():
    for _ in range(2):
        y = blk(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = blk(x)
    e1.record()
    torch.cuda.synchronize()
    print("eval block forward (BN1, conv1, BN2, conv2+gate): %.1f us per block" % (e0.elapsed_time(e1) * 100))
    torch.cuda.profiler.start()
    y = blk(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", ops.stats.get("gate_chain", 0))
