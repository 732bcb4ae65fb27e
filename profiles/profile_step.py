"""Per-kernel device-time breakdown of one training step (CUPTI via torch.profiler; eager launches,
so every kernel keeps its name).  Usage: python profiles/profile_step.py [--dtype bf16|f32] [--config cifar15]
[--batch 256] [--ncu]  (--ncu: bracket the step with cudaProfilerStart/Stop for `ncu --profile-from-start off`)."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lvae_b200  # noqa: E402
from lvae_b200.engine import TrainEngine  # noqa: E402
from lvae_b200.configs import baseline_config
from bench import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--config", default="cifar15")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--no-side-stream", action="store_true")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--hist", default="", help="comma list of kernel-name substrings: print a duration histogram for each")
args = ap.parse_args()

cfg = baseline_config(args.config)
torch.manual_seed(42)
model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
if args.dtype == "bf16":
    model.set_compute_dtype(torch.bfloat16)
eng = TrainEngine(model, args.batch, use_graph=False, wgrad_side_stream=not args.no_side_stream)
x = synthetic_batch(cfg, args.batch, 0).cuda()
for _ in range(2):
    eng.step(x)
torch.cuda.synchronize()
if args.ncu:
    torch.cuda.profiler.start()
    eng.step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    eng.step(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
durs = collections.defaultdict(list)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0]
        agg[name][0] += 1
        agg[name][1] += ev.device_time
        durs[name].append(ev.device_time)
tot = sum(v[1] for v in agg.values())
print("config %s batch %d dtype %s: %d kernels, %.2f ms of device time in one eager step" % (
    args.config, args.batch, args.dtype, sum(v[0] for v in agg.values()), tot / 1e3))
print("%-72s %6s %10s %6s %9s" % ("kernel", "n", "total us", "%", "avg us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print("%-72s %6d %10.1f %5.1f%% %9.2f" % (k[:72], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))

edges = [0, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 48, 64, 1e9]
for pat in [h for h in args.hist.split(",") if h]:
    for name, d in durs.items():
        if pat in name:
            print("histogram of %s (%d launches, %.1f us)" % (name[:60], len(d), sum(d)))
            for lo, hi in zip(edges[:-1], edges[1:]):
                sel = [t for t in d if lo <= t < hi]
                if sel:
                    print("   %5.0f-%-5.0f us: n=%4d total %8.1f us" % (lo, min(hi, 9999), len(sel), sum(sel)))
