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
