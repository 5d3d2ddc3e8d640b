"""Multi-GPU (run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`): the sharded,
shared-step solve over a SharedStepGroup must take the same step sequence as one GPU integrating the whole
batch, and return the same values for its shard (up to the summation order of the error norm)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import tfdiffeq_b200 as tfd
    from problems import PROBLEMS
    from tfdiffeq_b200.comm import SharedStepGroup, shard_bounds
    group = SharedStepGroup()
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((4099, 3))
    t = torch.arange(61, dtype=torch.float64) * 0.01
    f = PROBLEMS["lorenz"](backend="torch", device=dev)
    lo, hi = shard_bounds(y0.shape[0], world, rank)
    res = {}
    for method, kw in (("dopri5", {}), ("dopri8", dict(rtol=1e-9, atol=1e-9))):
        full = tfd.odeint(f, torch.tensor(y0, device=dev), t, method=method, **kw)
        s_full = dict(tfd.last_stats)
        for rep in range(2):                                   # twice: mailbox sequence numbers persist across solves
            part = tfd.odeint(f, torch.tensor(y0[lo:hi], device=dev), t, method=method,
                              options={"shared_step_group": group}, **kw)
        s_part = dict(tfd.last_stats)
        err = float((part - full[:, lo:hi]).abs().max())
        res[method] = (err, s_full["n_accepted"], s_full["n_rejected"], s_part["n_accepted"], s_part["n_rejected"])
    # the persistent fused kernel (built-in right-hand side) with the same group
    fb = tfd.rhs.Lorenz()
    full = tfd.odeint(fb, torch.tensor(y0, device=dev), t, method="dopri5")
    s_full = dict(tfd.last_stats)
    part = tfd.odeint(fb, torch.tensor(y0[lo:hi], device=dev), t, method="dopri5", options={"shared_step_group": group})
    s_part = dict(tfd.last_stats)
    assert s_full["fused_rhs"] and s_part["fused_rhs"]
    res["fused_dopri5"] = (float((part - full[:, lo:hi]).abs().max()), s_full["n_accepted"], s_full["n_rejected"],
                           s_part["n_accepted"], s_part["n_rejected"])
    group.close()
    q.put((rank, res))
    dist.destroy_process_group()


def test_shared_step_group_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        for method, (err, fa, fr, pa, pr) in res.items():
            assert (fa, fr) == (pa, pr), (rank, method, fa, fr, pa, pr)
            assert err < 1e-9, (rank, method, err)
