"""One bf16 forward + backward of the 3-layer MNIST model: how often conv_backward_raw found the bf16 gradient copy written by
lvae_stoch_bwd_ex (expected: once per conv_in_q and per non-top conv_in_p = 3 + 2) and how often it had to cast."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200 import ops
from lvae_b200.configs import baseline_config

hits, misses = [0], [0]
take = ops._lowp_grad_take
def counted(g, dt):
    r = take(g, dt)
    (hits if r is not None else misses)[0] += 1
    return r
ops._lowp_grad_take = counted
cfg = baseline_config("mnist3")
torch.manual_seed(0)
m = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
m.set_compute_dtype(torch.bfloat16)
x = (torch.rand(8, 1, 28, 28) < 0.2).float().cuda()
out = m(x)
((-out["ll"]).mean() + out["kl_loss"]).backward()
torch.cuda.synchronize()
print("bf16 gradient copies picked up: %d, casts: %d, loss %.4f" % (hits[0], misses[0], float((-out["ll"]).mean() + out["kl_loss"])))
