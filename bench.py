#!/usr/bin/env python
"""Benchmark of the Ladder-VAE hot path (BASELINE.json metric: train images/s, CIFAR10 15-layer LVAE,
batch 256 per GPU, data-parallel over N B200s).

    python bench.py --gpus N --steps K --warmup W              # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)
    python bench.py --workload iw ...                          # IW-1000 evals/s on the 12-layer MNIST model

For N > 1 launch with torch.distributed.run (one rank per GPU).  Rank 0 prints ONE JSON line.
A "step" is one ELBO training step (zero grads, forward, loss, backward, gradient all-reduce,
Adamax) over one synthetic batch.  `value` is timed with the batch already resident in HBM;
`e2e` is timed through the public TrainEngine.step(x_host) call with the pinned-host -> device
copy of the batch and the device -> host read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {"train": ("cifar15", 256), "iw": ("mnist12", 1000), "mnist12": ("mnist12", 128), "mnist3": ("mnist3", 64),
           "celeba20": ("celeba20", 64)}
# algorithmic conv GFLOP per image, forward / total
CONV_GFLOP = {"mnist3": (1.042, 3.126), "mnist12": (2.913, 8.737), "cifar15": (3.655, 10.962), "celeba20": (14.759, 44.266)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def synthetic_batch(cfg, batch, seed):
    """Bernoulli(0.15) binary images or uint8-grid uniform RGB, numpy seeded."""
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    shp = (batch, cfg.color_ch) + tuple(cfg.img_shape)
    if cfg.likelihood_form == "bernoulli":
        x = (rng.random_sample(shp) < 0.15).astype(np.float32)
    else:
        x = rng.randint(0, 256, size=shp).astype(np.float32) / 255.0
    return torch.from_numpy(x)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_train(cfg_name, cpu_batch, steps, warmup=1):
    """The reference's CPU path (oracle port, same op sequence in PyTorch CPU ops) on a bounded sample."""
    import torch
    from oracle import lvae_oracle as O
    cfg = O.baseline_config(cfg_name)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    st = O.TrainState(cfg, O.make_params(cfg, 42))
    x = synthetic_batch(cfg, cpu_batch, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st.step(x)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return dict(value=cpu_batch / t, unit="images/s", cores=torch.get_num_threads(), kind="port",
                sample="%d training steps of batch %d (same model, fp32, torch CPU ops via the oracle port), median"
                       % (steps, cpu_batch), s_per_step=t)


def cpu_baseline_iw(cpu_batch, forwards):
    import torch
    from oracle import lvae_oracle as O
    cfg = O.baseline_config("mnist12")
    torch.set_num_threads(os.cpu_count() or 1)
    P = O.make_params(cfg, 42)
    x = synthetic_batch(cfg, cpu_batch, 0)
    times = []
    with torch.no_grad():
        for i in range(1 + forwards):
            t0 = time.perf_counter()
            O.forward(P, cfg, x, None, None, False)
            if i >= 1:
                times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return dict(value=cpu_batch / t / 1000.0, unit="images/s with a 1000-sample bound", cores=torch.get_num_threads(), kind="port",
                sample="%d eval-mode forwards of batch %d, extrapolated to K=1000 (the reference recomputes the full "
                       "forward per sample)" % (forwards, cpu_batch), s_per_forward=t)


def run_reference(args, rank):
    if rank != 0:
        return
    name, batch = CONFIGS[args.workload]
    t0 = time.perf_counter()
    if args.workload == "iw":
        cb = cpu_baseline_iw(32, max(1, args.steps))
        metric, unit = "IW-1000 evals/s (MNIST 12-layer LVAE)", "images/s with a 1000-sample bound"   # = our arm's
        cfgd = {"workload": "importance-weighted bound K=1000, binarized-MNIST-shaped 12-layer LVAE", "cpu_batch": 32}
        ms = cb["s_per_forward"] * 1e3
    else:
        cb = cpu_baseline_train(name, 16, max(1, args.steps), max(1, min(args.warmup, 1)))
        metric, unit = "train images/s (CIFAR10 15-layer LVAE)", "images/s"
        cfgd = {"workload": "ELBO training step, %s, fp32" % name, "cpu_batch": 16, "per_gpu_batch": batch}
        ms = cb["s_per_step"] * 1e3
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "iw" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def conv_roofline(torch, pk, dtype="bf16"):
    """Dominant kernel: 3x3 64->64 stride-1 conv (335 of 538 forward convs) at the most
    common CIFAR-15 shape (B=256, 16x16).  Timed alone with CUDA events on the launch stream over
    rotating buffers larger than L2."""
    from lvae_b200 import _capi, ops
    B, H, W, C, k = 256, 16, 16, 64, 3
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    nbuf = 12 if dtype == "bf16" else 6             # (in + out) x nbuf > 126 MB L2
    xs = [torch.randn(B, H, W, C, device="cuda").to(tdt) for _ in range(nbuf)]
    ys = [torch.empty(B, H, W, C, device="cuda", dtype=tdt) for _ in range(nbuf)]
    bias = torch.zeros(C, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    w = torch.randn(C, C, k, k, device="cuda") / 24.0
    if dtype == "bf16":
        pack = ops.WeightPack(C, C, k * k, 2)
        wp = pack.get(w, torch.bfloat16)
        name = "conv_tc_kernel 3x3 64->64 B=256 16x16 (TMA + tcgen05.mma + TMEM, bf16 in / fp32 accumulate)"

        def launch(i):
            _capi.call("lvae_conv2d_tc", xs[i % nbuf].data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None,
                       ys[i % nbuf].data_ptr(), None, 0, B, H, W, C, C, k, 0, 0, s)
    else:
        pack = ops.WeightPack(C, C, k * k, 0)
        wp = pack.get(w, torch.float32)
        name = "conv_gather_kernel<float> 3x3 64->64 B=256 16x16 (fp32 CUDA-core implicit GEMM)"

        def launch(i):
            _capi.call("lvae_conv2d_gather", xs[i % nbuf].data_ptr(), None, wp.data_ptr(), bias.data_ptr(), None, None, None,
                       ys[i % nbuf].data_ptr(), B, H, W, C, 0, H, W, C, C, k, k, 1, 1, 0, 0, s)
    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    n = 48
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 2.0 * B * H * W * C * C * k * k
    achieved = flops / (us * 1e-6) / 1e12
    bytes_alg = 2.0 * B * H * W * C * (2 if dtype == "bf16" else 4)
    return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_burst"], "traffic": NCU_CONV_DRAM_BYTES if dtype == "bf16" else None,
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/ncu_conv_tc_r01_final.txt)",
            "us_per_launch": us, "flops_per_launch": flops,
            "algorithmic_bytes_per_launch": bytes_alg, "hbm_gbs_at_this_rate": bytes_alg / (us * 1e-6) / 1e9,
            "peak_source": "%s bf16 burst (kernel timed alone)" % pk["src"]}


def hbm_rooflines(workload):
    """Secondary roofline entries for the HBM-bound kernels north_star names (fused stochastic block, likelihoods): the
    microbenchmarks of profiles/bench_hbm_kernels.py (CUDA-graph replay over rotating buffers larger than L2, CUDA events on
    the capturing stream, algorithmic bytes of SURVEY.md 8d over the measured copy bandwidth).  Never fatal: an exception
    is reported in place of the numbers; the table it prints goes to stderr."""
    import contextlib
    import importlib.util
    try:
        spec = importlib.util.spec_from_file_location("bench_hbm_kernels", os.path.join(ROOT, "profiles", "bench_hbm_kernels.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        peak, src = mod.hbm_peak()
        rows = []
        with contextlib.redirect_stdout(sys.stderr):
            if workload == "iw":
                mod.bench_stochastic(rows, peak, src, B=1000)
                mod.bench_bernoulli(rows, peak, src)
            else:
                mod.bench_stochastic(rows, peak, src)
                mod.bench_dmol(rows, peak, src)
        keep = ("kernel", "shape", "us_per_launch", "achieved", "peak", "unit", "frac", "bound")
        return [{k: r[k] for k in keep} for r in rows]
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


# dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full` capture (profiles/ncu_conv_tc_r01_final.txt):
# 8.53 MB read (8.39 MB activations + weights), 0 written (the output is still dirty in L2 when the kernel ends)
NCU_CONV_DRAM_BYTES = 8526336


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--iw-samples", type=int, default=1000)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"],
                    help="bf16: bf16 activations, tcgen05 convs, fp32 accumulation (default); f32: exact fp32 CUDA-core path")
    ap.add_argument("--iw-full-forward", action="store_true",
                    help="IW: recompute the bottom-up pass for every sample like the reference's loop (default: once per batch)")
    ap.add_argument("--no-side-stream", action="store_true", help="keep weight-gradient kernels on the main stream")
    ap.add_argument("--side-streams", type=int, default=2, help="number of side streams the weight-gradient kernels rotate over")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-rooflines", action="store_true", help="skip the stochastic / likelihood kernel microbenchmarks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import lvae_b200
    from lvae_b200 import _capi
    from lvae_b200.engine import IWEvaluator, TrainEngine
    from lvae_b200.configs import baseline_config        # product-side config table (the oracle is only the cpu_baseline leg)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    cfg_name, batch = CONFIGS[args.workload]
    batch = args.batch or batch
    cfg = baseline_config(cfg_name)
    torch.manual_seed(42)
    lvae_b200.manual_seed(1234 + rank)
    model = lvae_b200.LadderVAE(**cfg.kwargs()).cuda()
    if args.dtype == "bf16":
        model.set_compute_dtype(torch.bfloat16)
    x_host = synthetic_batch(cfg, batch, rank).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    if args.workload == "iw":
        K = args.iw_samples
        ev = IWEvaluator(model, batch, use_graph=not args.no_graph, reuse_bottomup=not args.iw_full_forward)
        x_dev = x_host.cuda()
        k_warm = max(8, world)
        for _ in range(args.warmup):
            ev.bound(x_dev, k_warm)
        barrier()
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            res = ev.bound(x_dev, K)
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        value = batch / (ms * 1e-3)                                  # images whose K-sample bound completes per second
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = ev.bound(x_host, K)
            res_host = res.cpu()
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        _, k_local = lvae_b200.engine.shard_samples(K, rank, world)
        launches = (ev.launches_per_sample * k_local + ev.launches_bottomup) * args.steps
        line = {"metric": "IW-%d evals/s (MNIST 12-layer LVAE)" % K, "value": value, "unit": "images/s with a %d-sample bound" % K,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "importance-weighted bound, K=%d samples sharded over %d GPU(s), test batch %d, "
                                       "binarized-MNIST-shaped 12-layer LVAE, eval mode" % (K, world, batch),
                           "bottom_up_pass": "per sample (as the reference)" if args.iw_full_forward else "once per image batch (eval mode is deterministic)",
                           "l2_policy": "per-sample working set exceeds L2"},
                "sample_forwards_per_s": value * K,
                "e2e": {"value": batch / (e2e_ms * 1e-3), "unit": "images/s with a %d-sample bound" % K,
                        "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": res_host.numel() * 4},
                "gpu_launches": launches, "clocks": clocks}
        if rank == 0 and world == 1 and not args.no_hbm_rooflines:
            line["roofline_hbm"] = hbm_rooflines("iw")
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline_iw(32, 3).items() if k != "s_per_forward"}
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    engine = TrainEngine(model, batch, use_graph=not args.no_graph, wgrad_side_stream=0 if args.no_side_stream else args.side_streams)
    for _ in range(args.warmup):
        engine.step(x_host)
    torch.cuda.synchronize()
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- device-resident timing: `value` ----
    barrier()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        out = engine.step(None)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = batch * world / (ms * 1e-3)
    final_loss = float(out["loss"])

    # ---- end to end through the public API with host buffers: `e2e` ----
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = engine.step(x_host)                       # pinned host -> device copy inside
        loss_host.copy_(out["loss"], non_blocking=False)  # device -> host read of the loss
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    gflop = CONV_GFLOP[cfg_name][1]
    line = {"metric": "train images/s (%s LVAE)" % {"cifar15": "CIFAR10 15-layer"}.get(cfg_name, cfg_name),
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": "ELBO training step (zero grad, forward, loss, backward, gradient all-reduce, Adamax), "
                                   "%s, 10-component DMoL, dropout 0.2, train-mode BatchNorm" % cfg_name,
                       "per_gpu_batch": batch, "global_batch": batch * world, "parallelism": "dp%d" % world,
                       "cuda_graph": not args.no_graph,
                       "l2_policy": "inputs larger than L2: each step streams %.1f GB of activations" % mem_gb},
            "e2e": {"value": batch * world / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "wall_ms_per_step": e2e_wall_ms},
            "gpu_launches": engine.launches_per_step * args.steps,
            "launches_per_step": engine.launches_per_step,
            "clocks": clocks,
            "whole_step_conv_tflops": value / world * gflop / 1e3,
            "whole_step_frac_of_bf16_sustained": value / world * gflop / 1e3 / pk["tf_sustained"],
            "peak_mem_gb": mem_gb, "loss": final_loss}
    if rank == 0:
        line["roofline"] = conv_roofline(torch, pk, args.dtype)
        if world == 1 and not args.no_hbm_rooflines:
            line["roofline_hbm"] = hbm_rooflines(args.workload)
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline is a single-GPU-run item (rank 0, N = 1 only)
            cb = cpu_baseline_train(cfg_name, 16, 6)          # ~5-10 s of CPU work on the box's host cores
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
