"""Which Python lines launch the ATen glue kernels of one training step?  (188 launches / 0.75 ms of summed kernel time in
profiles/launches_r01_final_summary.csv: gradient-accumulation adds, dtype copies, fills, means ...)

    python profiles/trace_glue.py [--config cifar15] [--batch 256] [--dtype bf16]

One eager step under torch.profiler with Python stacks; every CPU op that is not one of our C-ABI calls and that launches a
CUDA kernel is attributed to the innermost stack frame inside this repository (forward) or to the autograd node that ran it
(backward).  Output: a table sorted by launch count -- the work list for fusing the glue away.
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lvae_b200  # noqa: E402
from lvae_b200.configs import baseline_config  # noqa: E402
from lvae_b200.engine import TrainEngine  # noqa: E402
from bench import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--config", default="cifar15")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--top", type=int, default=40)
args = ap.parse_args()

cfg = baseline_config(args.config)
torch.manual_seed(42)
model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
if args.dtype == "bf16":
    model.set_compute_dtype(torch.bfloat16)
eng = TrainEngine(model, args.batch, use_graph=False, wgrad_side_stream=0)
x = synthetic_batch(cfg, args.batch, 0).cuda()
for _ in range(2):
    eng.step(x)
torch.cuda.synchronize()

acts = [torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]
with torch.profiler.profile(activities=acts, with_stack=True, record_shapes=True) as prof:
    eng.step(x)
    torch.cuda.synchronize()


def repo_frame(stack):
    for fr in stack or []:
        if ROOT in fr and "profiles/" not in fr:
            return fr.replace(ROOT + "/", "")
    return (stack[0] if stack else "<no python frame: autograd engine thread>")


table = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
        continue
    kernels = [k for k in getattr(ev, "kernels", [])]
    if not kernels:
        continue
    # only leaf ops (an aten::to that calls aten::copy_ reports the kernel twice otherwise)
    if any(c.name.startswith("aten::") and getattr(c, "kernels", []) for c in getattr(ev, "cpu_children", [])):
        continue
    where = repo_frame(ev.stack)
    shape = str(ev.input_shapes[:2]) if ev.input_shapes else ""
    key = (ev.name, where)
    table[key][0] += len(kernels)
    table[key][1] += sum(k.duration for k in kernels)
    table[key][2][shape] += 1

rows = sorted(table.items(), key=lambda kv: -kv[1][0])
print("%-22s %5s %9s  %s" % ("aten op", "n", "dev us", "launched from (innermost frame in the repo) / most common input shapes"))
tot_n = tot_us = 0
for (name, where), (n, us, shapes) in rows:
    tot_n += n
    tot_us += us
for (name, where), (n, us, shapes) in rows[:args.top]:
    print("%-22s %5d %9.1f  %s   %s" % (name, n, us, where, shapes.most_common(1)[0][0]))
print("total: %d glue launches, %.1f us of device time" % (tot_n, tot_us))
