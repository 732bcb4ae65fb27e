"""CPU, world_size 2 over gloo: the host-side logic of the two multi-GPU paths -- bucketed gradient
all-reduce of the flat arena (data-parallel training) and the cross-rank merge of the
importance-weighted bound's running logsumexp states (sample sharding)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _lse_update(state, elbo):
    """CPU restatement of lvae_iw_lse_update (csrc/optim.cu) for the test."""
    m, s = state[:, 0], state[:, 1]
    nm = torch.maximum(m, elbo)
    s = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(s), s * torch.exp(m - nm)) + torch.exp(elbo - nm)
    return torch.stack([nm, s], 1)


def _lse_combine(states, k_total):
    """CPU restatement of lvae_iw_lse_combine."""
    m = states[:, :, 0].max(0)[0]
    s = (states[:, :, 1] * torch.exp(states[:, :, 0] - m)).sum(0)
    return m + torch.log(s) - math.log(k_total)


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lvae_b200
    from lvae_b200.engine import all_reduce_buckets, bucket_ranges, gather_states, shard_samples
    # --- data-parallel gradient all-reduce over buckets
    g = torch.Generator().manual_seed(100 + rank)
    n = 100_003
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    buckets = bucket_ranges(n, bucket_bytes=64 * 1024)
    assert buckets[0][0] == 0 and buckets[-1][1] == n and all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    all_reduce_buckets(flat, buckets)
    other = torch.randn(n, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    assert torch.allclose(flat, mine + other, atol=1e-6)
    # --- sample-sharded IW bound
    K, B = 11, 7
    gg = torch.Generator().manual_seed(5)
    elbo = torch.randn(K, B, generator=gg, dtype=torch.float64) * 40 - 700       # same on every rank
    start, cnt = shard_samples(K, rank, world)
    state = torch.stack([torch.full((B,), -math.inf, dtype=torch.float64), torch.zeros(B, dtype=torch.float64)], 1)
    for k in range(start, start + cnt):
        state = _lse_update(state, elbo[k])
    states = gather_states(state)
    assert tuple(states.shape) == (world, B, 2)
    bound = _lse_combine(states, K)
    ref = torch.logsumexp(elbo, 0) - math.log(K)
    assert torch.allclose(bound, ref, atol=1e-9)
    torch.save(bound, os.path.join(tmp, "bound_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "bound_0.pt"))
    b = torch.load(os.path.join(tmp_path, "bound_1.pt"))
    assert torch.equal(a, b)


def test_shard_samples_partition():
    import sys
    sys.path.insert(0, ROOT)
    from lvae_b200.engine import shard_samples
    for k in (1, 7, 100, 1000):
        for world in (1, 2, 3, 8):
            parts = [shard_samples(k, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == k
            pos = 0
            for s, c in parts:
                assert s == pos
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _worker_overlap(rank, world, port, tmp):
    """The overlapped reduction's host logic end to end on gloo: parameters announce their gradients in backward order
    (ops._param_grad_buffer -> ops._grad_note), each bucket is all-reduced by the flush callback once its last gradient
    has been announced AND the next gradient site has begun (ops.grad_site: the announcing site's own launches are then
    behind it), the rest at the end -- and the arena ends up as the sum over ranks."""
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bisect
    import lvae_b200  # noqa: F401
    from lvae_b200 import ops
    from lvae_b200.engine import bucket_ranges_aligned
    g = torch.Generator().manual_seed(7)
    sizes = [int(s) for s in torch.randint(1, 4000, (60,), generator=g)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in sizes]          # arena (= gradient-ready) order
    offs, off = {}, 0
    for p_, s in zip(params, sizes):
        offs[id(p_)] = off
        off += (s + 3) // 4 * 4
    grad = torch.zeros(off)
    buckets = bucket_ranges_aligned(sorted(offs.values()), off, bucket_bytes=32 * 1024)
    assert len(buckets) >= 4 and buckets[0][0] == 0 and buckets[-1][1] == off
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    assert all(s in set(offs.values()) for s, _ in buckets)        # every cut on a tensor boundary: no tensor straddles two buckets
    starts = [s for s, _ in buckets]
    bucket_of = {pid: bisect.bisect_right(starts, o) - 1 for pid, o in offs.items()}
    silent = {id(params[17])}                                             # one gradient autograd accumulates by itself
    counts = [0] * len(buckets)
    for p_ in params:
        if id(p_) not in silent:
            counts[bucket_of[id(p_)]] += 1
    kq = bucket_of[id(params[17])]
    counts[kq] = -1                                                       # its bucket waits for the end
    flushed, written = [], set()

    def flush(k):
        s, e = buckets[k]
        # everything of bucket k must have been written by now (except the silent one, which only the final flush covers)
        for p_ in params:
            if bucket_of[id(p_)] == k and id(p_) not in silent:
                assert id(p_) in written, "bucket %d flushed before all of its gradients were issued" % k
        flushed.append(k)
        dist.all_reduce(grad[s:e])

    ops.grad_track_begin(bucket_of, counts, flush)
    gr = torch.Generator().manual_seed(1000 + rank)
    local = torch.zeros(off)
    for p_ in params:
        p_._lvae_grad_sink = grad[offs[id(p_)]:offs[id(p_)] + p_.numel()]
        v = torch.randn(p_.numel(), generator=gr)
        local[offs[id(p_)]:offs[id(p_)] + p_.numel()] = v
        if id(p_) in silent:
            p_._lvae_grad_sink.copy_(v)                                   # written without an announcement
            continue
        # a gradient site: flush what earlier sites completed, announce (twice, like weight + bias of one conv: the second
        # announcement must not trigger the flush of a bucket this very site completes), THEN "launch" the gradient kernel
        ops.grad_site()
        sink, sunk = ops._param_grad_buffer(p_)
        sink2, _ = ops._param_grad_buffer(p_)
        assert sunk and sink2 is sink
        sink.copy_(v)
        written.add(id(p_))
    early = list(flushed)
    ops.grad_track_end()
    for k in range(len(buckets)):
        if k not in flushed:
            flush(k)
    assert early == sorted(early) and kq not in early and len(early) >= len(buckets) - 2
    assert sorted(flushed) == list(range(len(buckets)))
    both = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(both, local)
    assert torch.allclose(grad, sum(both), atol=1e-6)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_bucket_flush_two_rank_gloo(tmp_path):
    mp.spawn(_worker_overlap, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def test_grad_ready_order_and_iw_shard_offsets():
    import sys
    sys.path.insert(0, ROOT)
    import lvae_b200
    from lvae_b200.configs import baseline_config
    from lvae_b200.engine import grad_ready_order
    m = lvae_b200.LadderVAE(**baseline_config("mnist3").kwargs())
    names = {id(p): n for n, p in m.named_parameters()}
    order = [names[id(p)] for p in grad_ready_order(m)]
    assert sorted(order) == sorted(names.values()) and len(set(order)) == len(order)
    assert order[0].startswith("likelihood.") and order[-1] == "first_bottom_up.0.weight"
    first = {g: min(i for i, n in enumerate(order) if n.startswith(g)) for g in
             ("likelihood.", "final_top_down.", "top_down_layers.0.", "top_down_layers.2.", "bottom_up_layers.2.",
              "bottom_up_layers.0.", "first_bottom_up.")}
    seq = [first[g] for g in ("likelihood.", "final_top_down.", "top_down_layers.0.", "top_down_layers.2.",
                              "bottom_up_layers.2.", "bottom_up_layers.0.", "first_bottom_up.")]
    assert seq == sorted(seq)                       # the order LadderVAE's backward finishes them in


def test_bucket_cuts_and_sample_shards_property():
    """Randomised invariants of the two partitioners the multi-GPU paths rest on (hypothesis): gradient buckets tile the arena,
    are cut on tensor boundaries only and close as soon as they reach the size target; sample shards tile [0, K) in rank order
    with sizes that differ by at most one."""
    import sys
    sys.path.insert(0, ROOT)
    from hypothesis import given, settings, strategies as st
    from lvae_b200.engine import bucket_ranges_aligned, bucket_ranges, shard_samples

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=5000), min_size=1, max_size=60), st.integers(min_value=4, max_value=40000))
    def buckets(sizes, bucket_bytes):
        starts, pos = [], 0
        for n in sizes:
            starts.append(pos)
            pos += n
        numel = pos
        cuts = bucket_ranges_aligned(starts, numel, bucket_bytes)
        per = max(1, bucket_bytes // 4)
        assert cuts[0][0] == 0 and cuts[-1][1] == numel
        for (a, b), (c, _) in zip(cuts, cuts[1:]):
            assert b == c                                  # contiguous, no gap, no overlap
        sset = set(starts)
        for a, b in cuts:
            assert a in sset and a < b                     # every cut is a tensor boundary; no empty bucket
        for a, b in cuts[:-1]:
            assert b - a >= per                            # a bucket closes only once it has reached the target ...
            last_start = max(s for s in starts if s < b)
            assert last_start - a < per                    # ... and does so at the first tensor boundary past it
        plain = bucket_ranges(numel, bucket_bytes)
        assert plain[0][0] == 0 and plain[-1][1] == numel and all(b - a <= per for a, b in plain)

    @settings(max_examples=200, deadline=None)
    @given(st.integers(min_value=0, max_value=5000), st.integers(min_value=1, max_value=64))
    def shards(k, world):
        parts = [shard_samples(k, r, world) for r in range(world)]
        pos = 0
        for s, c in parts:
            assert s == pos and c >= 0
            pos += c
        assert pos == k
        sizes = [c for _, c in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    buckets()
    shards()
