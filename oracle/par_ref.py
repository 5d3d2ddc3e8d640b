"""ORACLE -- TEST INFRASTRUCTURE ONLY.  The oracle port (``np_ref``) on ALL host cores: the batch is split over worker
processes and the handful of whole-tensor reductions the reference's algorithm contains are combined across them.

Why this exists: ``bench.py --impl reference`` / ``cpu_baseline`` must time the reference's algorithm "with all the host
threads it can use".  The reference is op-by-op eager Python (TF-Eager; ``README.md:24`` equates it with PyTorch eager);
numpy runs each op on one core and multi-threaded torch-CPU loses to it on tensors this small (measured in round 1:
2.5e5 vs 1.0e7 element-steps/s on 128 cores).  What does scale is data parallelism over the batch axis -- exactly the
sharding of SURVEY 8(e): trajectories are independent except for the single shared step size, whose inputs are four
reductions per attempted step (``tfdiffeq/misc.py:257-263`` max|y0|, max|y1|, mean(ratio^2); ``dopri5.py:100`` finite
check) and three L2 norms in ``_select_initial_step`` (``misc.py:170-175``).  Every worker runs the unchanged
``np_ref.odeint`` on its shard with those five helper reductions (``_absmax, _mean, _l2, _numel, _nonfinite``) replaced by
group-wide versions over shared memory, so all workers take the same step sequence as the single-process oracle (to
summation order: the mean is a sum of per-shard sums).

Only ``bench.py``'s reference arm / ``cpu_baseline`` and ``tests/`` may use this module.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _Group(object):
    """All-reduce of one double among `n` forked workers through shared memory: rank r writes vals[parity][r], publishes
    its operation counter in seq[r], spins until every seq[q] has reached the counter, combines in rank order."""

    def __init__(self, n):
        self.n = n
        self.vals = np.frombuffer(mp.RawArray("d", 2 * n), dtype=np.float64).reshape(2, n)
        self.seq = np.frombuffer(mp.RawArray("q", n), dtype=np.int64)
        self.rank = None
        self.count = 0

    def _exchange(self, v):
        self.count += 1
        s, par, r = self.count, self.count & 1, self.rank
        self.vals[par, r] = v
        self.seq[r] = s
        seq = self.seq
        spins = 0
        while seq.min() < s:
            spins += 1
            if spins > 2000:
                time.sleep(0)          # oversubscribed host: yield instead of burning the core a peer needs
        return self.vals[par].copy()

    def sum(self, v):
        return float(self._exchange(v).sum()) if self.n > 1 else float(v)

    def max_nan(self, v):
        if self.n == 1:
            return float(v)
        a = self._exchange(v)
        return float("nan") if np.isnan(a).any() else float(a.max())


def _install(group, np_ref):
    """Replace np_ref's five reduction helpers by group-wide ones (numpy arrays only)."""
    def absmax(x):
        dt = x.dtype
        loc = np.max(np.abs(x)) if x.size else -np.inf
        return dt.type(group.max_nan(float(loc)))

    def mean(x):
        s = group.sum(float(np.sum(x, dtype=np.float64)))
        n = group.sum(float(x.size))
        return x.dtype.type(s / n)

    def l2(x):
        return x.dtype.type(np.sqrt(group.sum(float(np.sum(x * x, dtype=np.float64)))))

    def numel(x):
        return int(round(group.sum(float(x.size))))

    def nonfinite(x):
        return group.sum(1.0 if np.any(~np.isfinite(x)) else 0.0) > 0
    np_ref._absmax, np_ref._mean, np_ref._l2, np_ref._numel, np_ref._nonfinite = absmax, mean, l2, numel, nonfinite


def _bounds(n, world, rank):
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _worker(rank, group, problem, pkw, y0, t, kw, reps, times, counts, out):
    for p in (HERE, os.path.join(os.path.dirname(HERE), "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["OMP_NUM_THREADS"] = "1"
    import np_ref
    from problems import PROBLEMS
    group.rank = rank
    _install(group, np_ref)
    lo, hi = _bounds(y0.shape[0], group.n, rank)
    f = PROBLEMS[problem](backend="numpy", dtype=np.dtype(y0.dtype).name, **pkw)
    shard = np.ascontiguousarray(y0[lo:hi])
    for i in range(reps):
        group.sum(0.0)                                     # start line
        t0 = time.perf_counter()
        st = np_ref.Stats()
        sol = np_ref.odeint(f, shard, t, stats=st, **kw)
        group.sum(0.0)                                     # finish line: the slowest worker defines the time
        dt = time.perf_counter() - t0
        if rank == 0:
            times[i] = dt
            # a fixed grid "accepts" every cell (the oracle's fixed-grid driver keeps no accept counter)
            counts[0], counts[1], counts[2] = (st.n_acc or (st.n_rej == 0 and len(t) - 1) or 0), st.n_rej, st.nfe
        if out is not None and i == reps - 1:
            out[:, lo:hi] = sol


def solve(problem, y0, t, nproc=None, reps=1, want_solution=False, pkw=None, **kw):
    """Run ``np_ref.odeint(PROBLEMS[problem], y0, t, **kw)`` ``reps`` times on ``nproc`` processes (batch axis 0 split
    contiguously).  Returns ``dict(seconds=[...], n_acc, n_rej, nfe, nproc, solution or None)``."""
    y0 = np.ascontiguousarray(y0)
    t = np.asarray(t, dtype=np.float64)
    nproc = int(nproc or os.cpu_count() or 1)
    nproc = max(1, min(nproc, y0.shape[0]))
    ctx = mp.get_context("fork")
    group = _Group(nproc)
    times = np.frombuffer(mp.RawArray("d", reps), dtype=np.float64)
    counts = np.frombuffer(mp.RawArray("q", 3), dtype=np.int64)
    out = None
    if want_solution:
        shape = (len(t),) + y0.shape
        out = np.frombuffer(mp.RawArray("d" if y0.dtype == np.float64 else "f", int(np.prod(shape))), dtype=y0.dtype).reshape(shape)
    procs = [ctx.Process(target=_worker, args=(r, group, problem, pkw or {}, y0, t, kw, reps, times, counts, out))
             for r in range(nproc)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    if any(p.exitcode != 0 for p in procs):
        raise RuntimeError("a parallel-oracle worker failed: exit codes %s" % [p.exitcode for p in procs])
    return dict(seconds=[float(x) for x in times], n_acc=int(counts[0]), n_rej=int(counts[1]), nfe=int(counts[2]),
                nproc=nproc, solution=None if out is None else np.array(out))
