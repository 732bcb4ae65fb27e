"""Timeline analysis of ONE CUDA-graph-replayed training step (CUPTI): span, busy time, idle gaps."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lvae_b200
from lvae_b200.engine import TrainEngine
from lvae_b200.configs import baseline_config
from bench import synthetic_batch

cfg = baseline_config("cifar15")
torch.manual_seed(42)
model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
model.set_compute_dtype(torch.bfloat16)
eng = TrainEngine(model, 256, use_graph=True)
x = synthetic_batch(cfg, 256, 0).cuda()
for _ in range(4):
    eng.step(x)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    eng.step(x)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print("kernels %d  span %.2f ms  sum of durations %.2f ms" % (len(evs), (t1 - t0) / 1e3, sum(e.time_range.end - e.time_range.start for e in evs) / 1e3))
# union coverage + gaps
cur_end, busy, gaps = evs[0].time_range.start, 0.0, []
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, e.name[:50]))
        busy += en - s
        cur_end = en
    elif en > cur_end:
        busy += en - cur_end
        cur_end = en
print("busy (union) %.2f ms  idle %.2f ms in %d gaps (avg %.2f us)" % (busy / 1e3, (t1 - t0 - busy) / 1e3, len(gaps), (sum(g for g, _ in gaps) / max(1, len(gaps)))))
hist = collections.Counter()
for g, _ in gaps:
    hist[min(int(g), 10)] += 1
print("gap histogram (us: count):", sorted(hist.items()))
# per-kernel start-to-start intervals by name (critical-path feel)
agg = collections.defaultdict(lambda: [0, 0.0])
for a, b in zip(evs, evs[1:]):
    name = a.name.replace("(anonymous namespace)::", "").split("(")[0][:60]
    agg[name][0] += 1
    agg[name][1] += b.time_range.start - a.time_range.start
print("%-62s %6s %10s %8s" % ("kernel (time until the NEXT kernel starts)", "n", "total us", "avg us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-62s %6d %10.1f %8.2f" % (k, v[0], v[1], v[1] / v[0]))
