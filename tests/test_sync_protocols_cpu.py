"""Model checks of the mbarrier protocols added to csrc/conv_tcgen05.cu without a GPU at hand: random interleavings of
the participating warps over a small Python model of mbarrier phases (arrival count + transaction bytes) must never
deadlock and must hand every consumer exactly the producer's sequence.  This checks the PROTOCOL (counts, parities,
who arrives where), not the PTX; the kernels' own parity tests run on the GPU (`profiles/ab_switches.sh`, `ab_cta2.sh`).

* ring: dynamic tile scheduler (`draw_tile` / `take_tile`): one producer lane publishes tile indices through TC_RING slots,
  nine consumer warps (MMA + 8 epilogue) read each slot; `published[i]` counts 1, `consumed[i]` counts 9.
* pair: cta_group::2 kernels: the leader's `full[s]` collects both CTAs' TMA bytes (the peer's may land before the leader
  arms it), `tcgen05.commit` multicast arrives on `empty[s]` and `tmem_full[buf]` of both CTAs, the sixteen epilogue warps of
  both CTAs arrive on the leader's `tmem_empty[buf]`.
* chain: conv_gate_tc_kernel (csrc/conv_gate_tcgen05.cu): the 3x3 MMAs of tile i+1 are issued before the gate MMAs of tile i,
  with ONE staged-c2 buffer (reused for `out`) and ONE gate accumulator; the gate GEMM must always find tile i's c2 staged.
* gate_dgrad: gate_dgrad_tc_kernel (csrc/gate_dgrad_tcgen05.cu): eight warps produce operand tile i+1 into the other buffer
  before draining accumulator i.  (The model shows why the CTA barrier in front of the `a_ready` arrive is essential: without
  it a slow warp can miss a phase of `a_free`.)
"""
import random


class MBar:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.phase = count, count, 0, 0

    def _flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.init

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        self._flip()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._flip()

    def ready(self, parity):            # mbarrier.try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


def _run(gens, rnd, background=None):
    alive, steps = list(range(len(gens))), 0
    while alive:
        if background is not None:
            next(background)
        i = rnd.choice(alive)
        try:
            next(gens[i])
        except StopIteration:
            alive.remove(i)
        steps += 1
        assert steps < 10 ** 6, "deadlock"


def _ring(seed, n_tiles, ring=4, ncons=9):
    rnd = random.Random(seed)
    published = [MBar(1) for _ in range(ring)]
    consumed = [MBar(ncons) for _ in range(ring)]
    ids, counter, seen = [None] * ring, [0], [[] for _ in range(ncons)]

    def producer():
        ri = ph = 0
        while True:
            t = counter[0]
            counter[0] += 1
            t = t if t < n_tiles else -1
            while not consumed[ri].ready(ph ^ 1):
                yield
            ids[ri] = t
            published[ri].arrive()
            yield
            ri += 1
            if ri == ring:
                ri, ph = 0, ph ^ 1
            if t < 0:
                return
            for _ in range(rnd.randint(0, 3)):
                yield

    def consumer(k):
        ri = ph = 0
        while True:
            while not published[ri].ready(ph):
                yield
            t = ids[ri]
            yield
            consumed[ri].arrive()
            ri += 1
            if ri == ring:
                ri, ph = 0, ph ^ 1
            seen[k].append(t)
            if t < 0:
                return
            for _ in range(rnd.randint(0, 5)):
                yield

    _run([producer()] + [consumer(k) for k in range(ncons)], rnd)
    assert all(s == list(range(n_tiles)) + [-1] for s in seen)


def test_dynamic_scheduler_ring_protocol():
    for seed in range(150):
        _ring(seed, random.Random(seed).randint(1, 40))


def _pair(seed, iters, stages):
    rnd = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]                                   # leader only
    empty = [[MBar(1) for _ in range(stages)] for _ in range(2)]              # per CTA
    tmem_full = [[MBar(1) for _ in range(2)] for _ in range(2)]               # per CTA
    tmem_empty = [MBar(16) for _ in range(2)]                                 # leader only
    stage_tile = [[None] * stages for _ in range(2)]
    acc = [[None] * 2 for _ in range(2)]
    done, inflight = [[], []], []

    def producer(r):
        st = ph = 0
        for it in range(iters):
            while not empty[r][st].ready(ph ^ 1):
                yield
            if r == 0:
                full[st].arrive(expect_tx=2)
            inflight.append(("tma", r, st, 2 * it + r))
            yield
            st += 1
            if st == stages:
                st, ph = 0, ph ^ 1

    def mma():
        st = ph = 0
        for it in range(iters):
            buf, use = it & 1, it >> 1
            while not tmem_empty[buf].ready((use & 1) ^ 1):
                yield
            while not full[st].ready(ph):
                yield
            tiles = (stage_tile[0][st], stage_tile[1][st])
            assert tiles == (2 * it, 2 * it + 1)                               # both CTAs' operands of THIS tile pair
            inflight.append(("commit", st, buf, tiles))
            yield
            st += 1
            if st == stages:
                st, ph = 0, ph ^ 1

    def epilogue(r, w):
        for it in range(iters):
            buf, use = it & 1, it >> 1
            while not tmem_full[r][buf].ready(use & 1):
                yield
            assert acc[r][buf] == 2 * it + r
            yield
            tmem_empty[buf].arrive()
            if w == 0:
                done[r].append(2 * it + r)
            for _ in range(rnd.randint(0, 4)):
                yield

    def asynchronous():
        commits = []
        while True:
            if inflight and rnd.random() < 0.5:
                tmas = [e for e in inflight if e[0] == "tma"]
                commits = [e for e in inflight if e[0] == "commit"]
                ev = rnd.choice(tmas) if tmas and (not commits or rnd.random() < 0.5) else commits[0]   # commits retire in order
                inflight.remove(ev)
                if ev[0] == "tma":
                    _, r, st, tile = ev
                    stage_tile[r][st] = tile
                    full[st].complete_tx(1)
                else:
                    _, st, buf, tiles = ev
                    acc[0][buf], acc[1][buf] = tiles
                    for r in (0, 1):
                        empty[r][st].arrive()
                        tmem_full[r][buf].arrive()
            yield

    gens = [producer(0), producer(1), mma()] + [epilogue(r, w) for r in (0, 1) for w in range(8)]
    _run(gens, rnd, asynchronous())
    assert done[0] == [2 * i for i in range(iters)] and done[1] == [2 * i + 1 for i in range(iters)]


def test_cta_pair_barrier_protocol():
    for seed in range(120):
        _pair(seed, random.Random(seed).randint(1, 15), random.Random(seed + 1).randint(2, 4))


def _chain(seed, n_tiles, stages, taps):
    """conv_gate_tc_kernel: 3x3 MMAs of tile i+1 issued before the gate MMAs of tile i; one staged-c2 buffer and one gate
    accumulator; the epilogue runs phase 1 (stage c2) and phase 2 (drain the gate accumulator) of a tile back to back."""
    rnd = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    a1_full, a1_empty = [MBar(1), MBar(1)], [MBar(8), MBar(8)]
    c2_staged, a2_full, a2_empty = MBar(1), MBar(1), MBar(8)
    stage_tile = [None] * stages
    acc1, acc2, s_c2 = [None, None], [None], [None]
    inflight, outs = [], []

    def producer():
        st = ph = 0
        for tile in range(n_tiles):
            for t in range(taps):
                while not empty[st].ready(ph ^ 1):
                    yield
                full[st].arrive(expect_tx=1)
                inflight.append(("tma", st, (tile, t)))
                yield
                st += 1
                if st == stages:
                    st, ph = 0, ph ^ 1

    def mma():
        st = ph = 0

        def gate(j):
            while not a2_empty.ready((j & 1) ^ 1):
                yield
            while not c2_staged.ready(j & 1):
                yield
            assert s_c2[0] == ("c2", j)                    # the staged tile is tile j's c2, not yet overwritten by `out`
            inflight.append(("gate", j))
            yield

        for it in range(n_tiles):
            buf, use = it & 1, it >> 1
            while not a1_empty[buf].ready((use & 1) ^ 1):
                yield
            for t in range(taps):
                while not full[st].ready(ph):
                    yield
                assert stage_tile[st] == (it, t)
                inflight.append(("conv", st, buf, it, t == taps - 1))
                yield
                st += 1
                if st == stages:
                    st, ph = 0, ph ^ 1
            if it > 0:
                yield from gate(it - 1)
        if n_tiles > 0:
            yield from gate(n_tiles - 1)

    def epilogue(w):
        for it in range(n_tiles):
            buf, use = it & 1, it >> 1
            while not a1_full[buf].ready(use & 1):
                yield
            assert acc1[buf] == it
            yield
            a1_empty[buf].arrive()
            if w == 0:
                s_c2[0] = ("c2", it)                       # (all warps write; modelled once, after the staging barrier)
                c2_staged.arrive()
            yield
            while not a2_full.ready(it & 1):
                yield
            assert acc2[0] == it
            yield
            a2_empty.arrive()
            if w == 0:
                s_c2[0] = ("out", it)
                outs.append(it)
            for _ in range(rnd.randint(0, 4)):
                yield

    def asynchronous():
        while True:
            if inflight and rnd.random() < 0.6:
                tmas = [e for e in inflight if e[0] == "tma"]
                pipe = [e for e in inflight if e[0] != "tma"]              # the tensor pipe retires in issue order
                ev = rnd.choice(tmas) if tmas and (not pipe or rnd.random() < 0.5) else pipe[0]
                inflight.remove(ev)
                if ev[0] == "tma":
                    stage_tile[ev[1]] = ev[2]
                    full[ev[1]].complete_tx(1)
                elif ev[0] == "conv":
                    _, st, buf, it, last = ev
                    empty[st].arrive()
                    if last:
                        acc1[buf] = it
                        a1_full[buf].arrive()
                else:
                    assert s_c2[0] == ("c2", ev[1])
                    acc2[0] = ev[1]
                    a2_full.arrive()
            yield

    _run([producer(), mma()] + [epilogue(w) for w in range(8)], rnd, asynchronous())
    assert outs == list(range(n_tiles))


def test_conv_gate_chain_protocol():
    for seed in range(120):
        r = random.Random(seed)
        _chain(seed, r.randint(1, 12), r.randint(2, 5), r.choice([1, 9]))


def _gate_dgrad(seed, n_tiles):
    """gate_dgrad_tc_kernel: eight warps produce operand tile i+1 (double buffered) before they drain the accumulator of
    tile i; the MMA warp needs a produced operand and a drained accumulator."""
    rnd = random.Random(seed)
    a_ready, a_free = [MBar(1), MBar(1)], [MBar(1), MBar(1)]
    acc_full, acc_empty = [MBar(1), MBar(1)], [MBar(8), MBar(8)]
    s_a, acc, inflight, outs = [None, None], [None, None], [], []

    def mma():
        for it in range(n_tiles):
            buf, use = it & 1, it >> 1
            while not acc_empty[buf].ready((use & 1) ^ 1):
                yield
            while not a_ready[buf].ready(use & 1):
                yield
            assert s_a[buf] == it
            inflight.append((buf, it))
            yield

    cta_bar = {"n": 0, "gen": 0}                         # bar.sync 1, 256 among the eight warps

    def bar_sync():
        gen = cta_bar["gen"]
        cta_bar["n"] += 1
        if cta_bar["n"] == 8:
            cta_bar["n"], cta_bar["gen"] = 0, gen + 1
        while cta_bar["gen"] == gen:
            yield

    def warp(w):
        def produce(j):
            buf, use = j & 1, j >> 1
            while not a_free[buf].ready((use & 1) ^ 1):
                yield
            s_a[buf] = j                                   # every warp writes its share of the operand tile
            yield from bar_sync()                          # ... and only then may thread 0 publish it
            if w == 0:
                a_ready[buf].arrive()
            yield

        if n_tiles > 0:
            yield from produce(0)
        for it in range(n_tiles):
            buf, use = it & 1, it >> 1
            if it + 1 < n_tiles:
                yield from produce(it + 1)
            while not acc_full[buf].ready(use & 1):
                yield
            assert acc[buf] == it
            yield
            acc_empty[buf].arrive()
            if w == 0:
                outs.append(it)
            for _ in range(rnd.randint(0, 3)):
                yield

    def asynchronous():
        while True:
            if inflight and rnd.random() < 0.5:
                buf, it = inflight.pop(0)                  # in issue order
                assert s_a[buf] == it                      # the operand buffer was not overwritten while the MMAs read it
                acc[buf] = it
                a_free[buf].arrive()
                acc_full[buf].arrive()
            yield

    _run([mma()] + [warp(w) for w in range(8)], rnd, asynchronous())
    assert outs == list(range(n_tiles))


def test_gate_dgrad_chain_protocol():
    for seed in range(120):
        _gate_dgrad(seed, random.Random(seed).randint(1, 14))
