"""CPU, world_size 2 over gloo: the host-side logic of the sharded (N > 1) path.

No kernels run here.  What is checked is the arithmetic contract the in-kernel exchange implements
(tfdiffeq_b200/comm.py:combine_partials == `group_combine` in csrc/b2ode.cu): splitting the batch across
ranks, reducing {sum err^2, max|y0|, max|y1|, bad} per rank, gathering and combining in rank order gives every
rank the same accept / dt decision as the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import np_ref
    from problems import PROBLEMS
    from tfdiffeq_b200.comm import combine_partials, shard_bounds
    rng = np.random.default_rng(0)
    y0_full = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((1001, 3))     # deliberately not divisible
    lo, hi = shard_bounds(y0_full.shape[0], world, rank)
    f = PROBLEMS["lorenz"](backend="numpy")
    rtol, atol, dt, t0 = 1e-7, 1e-9, 0.013, 0.0
    res = []
    for y0 in (y0_full, y0_full * np.where(np.arange(1001)[:, None] == 7, np.nan, 1.0)):
        y_loc = (y0[lo:hi],)
        f0 = (f(t0, y_loc[0]),)
        y1, f1, err, k = np_ref.runge_kutta_step(lambda t, y: (f(t, y[0]),), y_loc, f0, t0, dt, np_ref.DOPRI5)
        with np.errstate(all="ignore"):
            mine = (float(np.sum(err[0].astype(np.float64) ** 2)), float(np.max(np.abs(y_loc[0]))),
                    float(np.max(np.abs(y1[0]))), float(np.any(~np.isfinite(y_loc[0]))))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        tot = combine_partials(gathered)
        n_global = y0.size
        tol = atol + rtol * (np.nan if (tot[1] != tot[1] or tot[2] != tot[2]) else max(tot[1], tot[2]))
        msr = tot[0] / (tol * tol) / n_global
        res.append((msr, tot[3]))
    q.put((rank, lo, hi, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_error_norm_equals_unsharded_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_ref
    from problems import PROBLEMS
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the batch exactly
    assert out[0][1] == 0 and out[0][2] == out[1][1] and out[1][2] == 1001
    assert out[0][2] - out[0][1] == 501 and out[1][2] - out[1][1] == 500
    # both ranks computed the identical group result ...
    assert out[0][3] == out[1][3] or (np.isnan(out[0][3][1][0]) and np.isnan(out[1][3][1][0]))
    # ... and it equals the unsharded oracle's mean-square error ratio
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((1001, 3))
    f = PROBLEMS["lorenz"](backend="numpy")
    f0 = (f(0.0, y0),)
    y1, f1, err, k = np_ref.runge_kutta_step(lambda t, y: (f(t, y[0]),), (y0,), f0, 0.0, 0.013, np_ref.DOPRI5)
    msr = np_ref.compute_error_ratio(err, [1e-7], [1e-9], (y0,), y1)[0]
    assert abs(out[0][3][0][0] - msr) <= 1e-12 * msr
    assert out[0][3][0][1] == 0.0
    # NaN in one shard poisons the tolerance on every rank and raises the non-finite flag everywhere
    assert np.isnan(out[0][3][1][0]) and np.isnan(out[1][3][1][0])
    assert out[0][3][1][1] == 1.0 and out[1][3][1][1] == 1.0


@pytest.mark.parametrize("n,w", [(10, 3), (8, 8), (5, 8), (65536, 8), (0, 2)])
def test_shard_bounds_tile(n, w):
    from tfdiffeq_b200.comm import shard_bounds
    edges = [shard_bounds(n, w, r) for r in range(w)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for a, b in zip(edges[:-1], edges[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1
