"""Per-kernel device-time breakdown of ONE importance-sample pass of the IW evaluator (CUPTI, eager launches).
Usage: python profiles/profile_iw.py [--dtype bf16|f32] [--batch 1000]"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200.engine import IWEvaluator
from lvae_b200.configs import baseline_config
from bench import synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=1000)
ap.add_argument("--top", type=int, default=24)
args = ap.parse_args()
cfg = baseline_config("mnist12")
torch.manual_seed(42)
model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
if args.dtype == "bf16":
    model.set_compute_dtype(torch.bfloat16)
ev = IWEvaluator(model, args.batch, use_graph=False)
x = synthetic_batch(cfg, args.batch, 0).cuda()
ev.bound(x, 2)
torch.cuda.synchronize()
with torch.no_grad():
    ev._bottomup()
    ev._one_sample()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        ev._one_sample()
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0]
        agg[name][0] += 1
        agg[name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("mnist12 batch %d dtype %s: %d kernels, %.2f ms of device time in one sample pass" % (args.batch, args.dtype, sum(v[0] for v in agg.values()), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print("%-72s %6d %10.1f %5.1f%% %9.2f" % (k[:72], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
# The same pass as ONE CUDA-graph replay (what the evaluator runs): start-to-start intervals.  A kernel launched under
# programmatic dependent launch starts early and waits for its predecessor, so its own duration includes that wait; the time
# until the NEXT kernel starts is the better measure of its cost on the dependent chain.
evg = IWEvaluator(model, args.batch, use_graph=True)
evg.bound(x, 3)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    evg.graph.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
nxt = collections.defaultdict(lambda: [0, 0.0])
for a, b in zip(evs, evs[1:]):
    name = a.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0]
    nxt[name][0] += 1
    nxt[name][1] += b.time_range.start - a.time_range.start
span = evs[-1].time_range.end - evs[0].time_range.start
print("\ngraph replay of one sample pass: %d kernels, span %.2f ms; time until the next kernel starts, per kernel:" % (len(evs), span / 1e3))
for k, v in sorted(nxt.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print("%-72s %6d %10.1f %5.1f%% %9.2f" % (k[:72], v[0], v[1], 100 * v[1] / span, v[1] / v[0]))
