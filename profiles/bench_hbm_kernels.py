"""Device-time microbenchmark of the HBM-bound kernels BASELINE.json's north_star names: the fused stochastic block
(lib/stochastic.py:45-96), the discretized-mixture-of-logistics and Bernoulli likelihoods (lib/likelihoods.py:291-388).

    python profiles/bench_hbm_kernels.py                  # one B200; prints a table and one JSON line per kernel

Each kernel is replayed from a CUDA graph over rotating buffers larger than L2 and timed with CUDA events on the
capturing stream.  `achieved` = ALGORITHMIC bytes per launch (stochastic forward 20 B / latent element
+ 4 B per pixel of kl_spatial, backward 40 B; DMoL forward 412 B / pixel, backward 816 B (fp32 gradient) or 672 B
(bf16 128-channel gradient operand); Bernoulli 8 / 12 B per pixel) / launch time; `peak` = the measured copy bandwidth of
MEASURED_PEAKS.json (fallback 6650 GB/s, B200_PROFILING.md).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lvae_b200  # noqa: E402,F401
from lvae_b200 import _capi, ops  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def S():
    return torch.cuda.current_stream().cuda_stream


def time_graph(fn, nbuf, n=24):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


def nbufs(bytes_per_launch):
    return max(2, int(300e6 // max(bytes_per_launch, 1)) + 1)


def report(rows, name, shape, nbytes, us, peak, src):
    gbs = nbytes / us / 1e3
    rows.append({"kernel": name, "shape": shape, "us_per_launch": us, "algorithmic_bytes": nbytes, "achieved": gbs,
                 "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bound": "hbm", "peak_source": src})
    print("%-26s %-26s %8.2f us  %7.0f GB/s  %5.1f %% of %s peak" % (name, shape, us, gbs, 100 * gbs / peak, src))


def bench_stochastic(rows, peak, src, B=256, Z=32, sides=(16, 8, 4, 2)):
    dev = torch.device("cuda")
    for hw_side in sides:
        hw = hw_side * hw_side
        n_el = B * hw * Z
        fwd_bytes = 20 * n_el + 4 * B * hw
        bwd_bytes = 40 * n_el
        nb = nbufs(bwd_bytes)
        q = [torch.randn(B, hw, 2 * Z, device=dev) * 0.3 for _ in range(nb)]
        p = [torch.randn(B, hw, 2 * Z, device=dev) * 0.3 for _ in range(nb)]
        z = [torch.empty(B, hw, Z, device=dev) for _ in range(nb)]
        gz = [torch.randn(B, hw, Z, device=dev) for _ in range(nb)]
        dq = [torch.empty(B, hw, 2 * Z, device=dev) for _ in range(nb)]
        dp = [torch.empty(B, hw, 2 * Z, device=dev) for _ in range(nb)]
        kl = torch.empty(B, device=dev)
        kls = torch.empty(B, hw, device=dev)
        logp = torch.empty(B, device=dev)
        logq = torch.empty(B, device=dev)
        g1 = torch.ones(B, device=dev)
        rng = ops.rng_state(dev)
        ws = ops._stoch_workspace(B, dev)

        def fwd(i):
            _capi.call("lvae_stoch_fwd", q[i].data_ptr(), p[i].data_ptr(), 0, None, None, rng.data_ptr(), 7 + i,
                       z[i].data_ptr(), None, Z, kl.data_ptr(), kls.data_ptr(), logp.data_ptr(), logq.data_ptr(), B, hw, Z,
                       0, 0, ws.data_ptr(), S())

        def bwd(i):
            _capi.call("lvae_stoch_bwd", q[i].data_ptr(), p[i].data_ptr(), 0, z[i].data_ptr(), gz[i].data_ptr(),
                       g1.data_ptr(), None, None, None, dq[i].data_ptr(), dp[i].data_ptr(), B, hw, Z, 0, 1, S())

        shape = "B=%d %dx%d Z=%d" % (B, hw_side, hw_side, Z)
        report(rows, "stoch_fwd (Philox eps)", shape, fwd_bytes, time_graph(fwd, nb), peak, src)
        report(rows, "stoch_bwd", shape, bwd_bytes, time_graph(bwd, nb), peak, src)


def bench_dmol(rows, peak, src, B=256, side=32):
    dev = torch.device("cuda")
    hw = side * side
    npix = B * hw
    nb = nbufs(816 * npix)
    l = [torch.randn(B, hw, 100, device=dev) * 0.5 for _ in range(nb)]
    x = [torch.randint(0, 256, (B, 3, hw), device=dev).float() / 255.0 for _ in range(nb)]
    dl = [torch.empty(B, hw, 100, device=dev) for _ in range(nb)]
    dlp = [torch.empty(B, hw, 128, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    ll = torch.zeros(B, device=dev)
    g = -torch.ones(B, device=dev) / B
    shape = "B=%d %dx%d" % (B, side, side)
    tag = ""
    report(rows, "dmol_fwd" + tag, shape, 412 * npix,
           time_graph(lambda i: _capi.call("lvae_dmol_fwd", l[i].data_ptr(), x[i].data_ptr(), ll.data_ptr(), B, hw, S()), nb),
           peak, src)
    report(rows, "dmol_bwd fp32 grad" + tag, shape, 816 * npix,
           time_graph(lambda i: _capi.call("lvae_dmol_bwd", l[i].data_ptr(), x[i].data_ptr(), g.data_ptr(), dl[i].data_ptr(),
                                           None, B, hw, S()), nb), peak, src)
    report(rows, "dmol_bwd bf16x128 grad" + tag, shape, (412 + 4 + 256) * npix,
           time_graph(lambda i: _capi.call("lvae_dmol_bwd", l[i].data_ptr(), x[i].data_ptr(), g.data_ptr(), None,
                                           dlp[i].data_ptr(), B, hw, S()), nb), peak, src)


def bench_bernoulli(rows, peak, src, B=1000, side=28):
    dev = torch.device("cuda")
    hw = side * side
    npix = B * hw
    nb = nbufs(12 * npix)
    lg = [torch.randn(B, hw, 1, device=dev) for _ in range(nb)]
    x = [(torch.rand(B, 1, hw, device=dev) < 0.15).float() for _ in range(nb)]
    pr = [torch.empty(B, hw, 1, device=dev) for _ in range(nb)]
    dl = [torch.empty(B, hw, 1, device=dev) for _ in range(nb)]
    ll = torch.empty(B, device=dev)
    g = -torch.ones(B, device=dev) / B
    shape = "B=%d %dx%d" % (B, side, side)
    # forward also writes the probabilities (the module API returns them): 8 B read + 4 B written per pixel
    report(rows, "bernoulli_fwd", shape, 12 * npix,
           time_graph(lambda i: _capi.call("lvae_bernoulli_fwd", lg[i].data_ptr(), x[i].data_ptr(), pr[i].data_ptr(),
                                           ll.data_ptr(), B, hw, 1, S()), nb), peak, src)
    report(rows, "bernoulli_bwd", shape, 12 * npix,
           time_graph(lambda i: _capi.call("lvae_bernoulli_bwd", pr[i].data_ptr(), x[i].data_ptr(), g.data_ptr(), None,
                                           dl[i].data_ptr(), B, hw, 1, S()), nb), peak, src)


def main():
    _capi.device_check()
    peak, src = hbm_peak()
    rows = []
    bench_stochastic(rows, peak, src)
    bench_dmol(rows, peak, src)
    bench_bernoulli(rows, peak, src)
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
