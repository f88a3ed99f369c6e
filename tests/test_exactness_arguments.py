"""Arguments the device code relies on, checked on the CPU (no GPU, no product code; section 3 is at the end).

1. Order-independent sums (kmeans.cu `update_body_warp` / `stats_body` fast paths): the reference
   adds a cluster's members sequentially (f32 centroid sums kmeans.rs:388-418, f64 loss :266-280).  If every term is an
   integer multiple of 2^g and sum|term| < 2^(g+p) (p = 24 for f32, 53 for f64) no addition of ANY association rounds,
   so a parallel tree reduction returns the sequential result bit for bit.  Outside the condition it does not.

2. Convergence polling through progress words (kmeans.cu `PollWords`, `lloyd_train`): after
   enqueuing iteration `it` the host waits until the problem has reported iteration it - 1, then reads the active
   bit.  A thread-per-rank simulation checks (a) that a single rank always stops, at most one no-op iteration late,
   however late it reads; (b) WHY sharded runs keep the blocking poll: a rank that reads late sees a later
   iteration's bit, enqueues fewer iterations than its peer, and the peer's next collective never completes; (c) that
   reporting the TICK OF CONVERGENCE instead of the current bit would make the decision independent of read timing
   (the design noted for sharded runs in DESIGN.md section 8; not built)."""
import threading
import time

import numpy as np


def _tree_sum(v, dtype):
    v = v.astype(dtype)
    while len(v) > 1:
        if len(v) & 1:
            v = np.concatenate([v, np.zeros(1, dtype)])
        v = (v[0::2] + v[1::2]).astype(dtype)
    return v[0]


def _seq_sum(v, dtype):
    acc = dtype(0)
    for t in v.astype(dtype):
        acc = dtype(acc + t)
    return acc


def _granule(v):
    """largest g with every non-zero term a multiple of 2^g (kmeans.cu pow2_granule)"""
    g = None
    for t in v:
        if t == 0:
            continue
        m, e = np.frexp(np.float64(abs(t)))        # t = m * 2^e, 0.5 <= m < 1
        k = 0
        while m != np.floor(m):
            m *= 2
            k += 1
        g = (e - k) if g is None else min(g, e - k)
    return g


def test_integer_valued_f32_terms_sum_identically_in_any_order():
    rng = np.random.default_rng(1)
    for _ in range(50):
        n = int(rng.integers(2, 3000))
        v = rng.integers(0, 256, n).astype(np.float32)          # SIFT / u8 columns: g = 0, sum < 2^24
        assert v.sum(dtype=np.float64) < 2 ** 24
        s = _seq_sum(v, np.float32)
        assert _tree_sum(v, np.float32) == s
        assert _tree_sum(rng.permutation(v), np.float32) == s
        assert s == np.float32(v.sum(dtype=np.float64))


def test_general_granule_condition_for_the_f64_loss():
    rng = np.random.default_rng(2)
    for _ in range(30):
        n = int(rng.integers(2, 2000))
        g = int(rng.integers(-30, 10))
        v = rng.integers(0, 1 << 20, n).astype(np.float64) * 2.0 ** g      # f32 distances widened to f64
        assert _granule(v) >= g and np.abs(v).sum() < 2.0 ** (g + 53)
        s = _seq_sum(v, np.float64)
        assert _tree_sum(v, np.float64) == s and _tree_sum(rng.permutation(v), np.float64) == s


def test_outside_the_condition_the_order_matters():
    """why the kernels TEST the condition per cluster and otherwise run the sequential chain"""
    rng = np.random.default_rng(3)
    differs = 0
    for _ in range(20):
        v = (rng.standard_normal(2000) * 100).astype(np.float32)            # arbitrary f32 residuals
        differs += int(_tree_sum(v, np.float32) != _seq_sum(v, np.float32))
    assert differs > 0
    big = np.full(70000, 255.0, np.float32)                                 # integers, but sum >= 2^24
    assert big.sum(dtype=np.float64) >= 2 ** 24
    assert _seq_sum(big, np.float32) != np.float32(big.sum(dtype=np.float64)) or _tree_sum(big, np.float32) != _seq_sum(big, np.float32)


class _Rank:
    """one rank of the simulation: a `device` thread executes enqueued iterations in order (each takes `iter_s`);
    iteration j contains the collective of iteration j, i.e. it needs every rank's device to reach it (a barrier);
    its epilogue posts conv (the tick at which the problem became inactive, 0 while active) and then the progress
    word (tick << 1 | active).  The host loop is lloyd_train's; `late_by` makes the host read only once the device is
    that many iterations further than it had to wait for (pre-emption, a slow graph launch, a CPU quota ...) --
    scripted in ticks, not in seconds, so that the tests do not depend on the machine's speed."""

    def __init__(self, world, barrier_for, converge_at, max_iters, mode, late_by=0, iter_s=0.0):
        self.world, self.barrier_for, self.converge_at, self.max_iters = world, barrier_for, converge_at, max_iters
        self.mode, self.late_by, self.iter_s = mode, late_by, iter_s
        self.queue, self.cv = [], threading.Condition()
        self.word = self.conv = self.enqueued = self.executed_active = 0
        self.hist = {0: (0, 0)}                        # tick -> (word, conv) as posted at that tick
        self.stop = self.hung = False

    def device(self):
        active, tick = True, 0
        while True:
            with self.cv:
                while not self.queue and not self.stop:
                    self.cv.wait(0.001)
                if not self.queue:
                    return
                j = self.queue.pop(0)
            try:
                self.barrier_for(j).wait(timeout=4.0)  # the exchange inside the captured iteration
            except threading.BrokenBarrierError:
                self.hung = True                       # a peer never enqueued iteration j: NCCL would wait forever
                return
            time.sleep(self.iter_s)
            if active:
                self.executed_active += 1
                if j >= self.converge_at:              # kmeans.rs:704 on identical models: same j on every rank
                    active = False
            tick += 1
            if not active and self.conv == 0:
                self.conv = tick                       # written BEFORE the word (fence in between on the device)
            self.hist[tick] = ((tick << 1) | int(active), self.conv)
            self.word = (tick << 1) | int(active)

    def host(self):
        def launch(j):
            with self.cv:
                self.queue.append(j)
                self.enqueued += 1
                self.cv.notify()
        launch(1)
        done, it = self.max_iters == 1, 2
        while it <= self.max_iters and not done:
            launch(it)
            want = it - 1
            t0 = time.monotonic()
            while (self.word >> 1) < want and not self.hung:   # (the real host spins; here it yields the interpreter lock)
                time.sleep(0.0001)
                assert time.monotonic() - t0 < 20, "progress word never arrived"
            seen = min(want + self.late_by, self.enqueued)     # the tick whose posting this (late) read observes
            t1 = time.monotonic()
            while (self.word >> 1) < seen and not self.hung and time.monotonic() - t1 < 6.0:
                time.sleep(0.0001)
            word, conv = self.hist[min(seen, self.word >> 1)]
            if self.mode == "active_bit":
                done = (word & 1) == 0                         # what is built (single rank)
            else:
                done = conv != 0 and conv <= want              # independent of WHEN the host reads
            it += 1
        t0 = time.monotonic()
        while (self.word >> 1) < self.enqueued and not self.hung and time.monotonic() - t0 < 8:
            time.sleep(0.0001)                                 # the final synchronise of lloyd_train
        self.stop = True


def _simulate(world, converge_at, max_iters, mode="active_bit", late=None, iter_s=0.0):
    barriers, lock = {}, threading.Lock()

    def barrier_for(j):
        with lock:
            if j not in barriers:
                barriers[j] = threading.Barrier(world)
            return barriers[j]

    ranks = [_Rank(world, barrier_for, converge_at, max_iters, mode, (late or [0] * world)[r], iter_s) for r in range(world)]
    threads = [threading.Thread(target=r.device) for r in ranks] + [threading.Thread(target=r.host) for r in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
        assert not t.is_alive(), "simulation did not end"
    return ranks


def test_single_rank_always_stops_at_most_one_noop_late_whatever_the_read_timing():
    for converge_at, max_iters, late in ((7, 50, 0), (7, 50, 1), (1, 50, 0), (50, 50, 0), (60, 50, 1), (3, 4, 0), (1, 1, 0)):
        r, = _simulate(1, converge_at, max_iters, late=[late], iter_s=0.0005)
        last_real = min(converge_at, max_iters)
        assert not r.hung and r.executed_active == last_real            # extra iterations were no-ops
        assert last_real <= r.enqueued <= min(max_iters, last_real + 1)


def test_two_ranks_reading_the_current_bit_can_diverge_which_is_why_sharded_runs_poll_blocking():
    # rank 0 reads one iteration late: waiting for iteration 4 it already sees iteration 5's bit and stops after
    # enqueuing 5; rank 1 sees that bit one check later, after enqueuing 6 -- whose collective rank 0 never joins
    ranks = _simulate(2, 5, 50, mode="active_bit", late=[1, 0], iter_s=0.0005)
    assert ranks[0].enqueued == 5 and ranks[1].enqueued == 6 and ranks[1].hung


def test_reporting_the_tick_of_convergence_is_independent_of_read_timing():
    for late in ([1, 0], [0, 1, 0], [1, 1]):
        ranks = _simulate(len(late), 5, 50, mode="conv_tick", late=late, iter_s=0.0005)
        assert not any(r.hung for r in ranks)
        assert {r.enqueued for r in ranks} == {6} and all(r.executed_active == 5 for r in ranks)


# ---- 3. the k + 1 rule of every top-k path (DESIGN.md section 5 "Ties at the k-th distance"; search.cu) ----------
# The GPU selects k + 1 candidates per list.  Argument: when the (k+1)-th smallest distance differs from the k-th, the
# reference's BinaryHeap result (flat/index.rs:82-177, restated in the oracle) is the unique set of the k smallest --
# whatever order the rows were pushed in -- so any selection algorithm returns it; only when they are equal does the
# result depend on the heap's sift order, and only then the slot is replayed with the heap's own operations.
def test_topk_set_is_order_free_unless_the_boundary_is_tied():
    from oracle import binding as ob
    rng = np.random.default_rng(4)
    tied_boundaries = 0
    for trial in range(200):
        n, k = int(rng.integers(5, 400)), int(rng.integers(1, 20))
        d = rng.integers(0, 40, n).astype(np.float32)                      # small integers: ties everywhere
        if trial % 3 == 0:
            d = d + rng.random(n).astype(np.float32)                        # and some tie-free inputs
        rid = np.arange(n, dtype=np.uint64)
        ids, dd = ob.flat_topk(d, rid, k)
        srt = np.sort(d, kind="stable")
        kk = min(k, n)
        assert np.array_equal(np.sort(dd), srt[:kk])                        # the distance multiset is always exact
        if kk < n and srt[kk] == srt[kk - 1]:
            tied_boundaries += 1                                            # heap order decides: the replay's domain
            continue
        expect = set(np.flatnonzero(d <= srt[kk - 1]).tolist()) if kk else set()
        assert set(ids.tolist()) == expect and len(expect) == kk
        perm = rng.permutation(n)                                           # push order does not matter here
        ids2, _ = ob.flat_topk(d[perm], rid[perm], k)
        assert set(ids2.tolist()) == expect
    assert tied_boundaries > 20                                             # the tied case is common on such data
