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
    # one group, shard sizes that change from solve to solve (the persistent kernel's grid shrinks and grows): a partial left
    # in the mailbox by an earlier, larger solve must never be taken for a fresh one (slots no longer written are poisoned)
    tt = torch.arange(21, dtype=torch.float64) * 0.01
    ref, worst = {}, 0.0
    for cyc in range(6):
        for n in (40000, 700):
            yb = np.array([1., 1., 1.]) + 0.1 * np.random.default_rng(n).standard_normal((n, 3))
            if n not in ref:
                ref[n] = (tfd.odeint(fb, torch.tensor(yb, device=dev), tt, method="dopri5"), dict(tfd.last_stats))
            l_, h_ = shard_bounds(n, world, rank)
            part = tfd.odeint(fb, torch.tensor(yb[l_:h_], device=dev), tt, method="dopri5",
                              options={"shared_step_group": group})
            sp = dict(tfd.last_stats)
            assert sp["fused_rhs"]
            assert (sp["n_accepted"], sp["n_rejected"]) == (ref[n][1]["n_accepted"], ref[n][1]["n_rejected"]), (cyc, n)
            worst = max(worst, float((part - ref[n][0][:, l_:h_]).abs().max()))
    res["fused_regrid"] = (worst, 0, 0, 0, 0)
    # sharded odeint_adjoint: y / adj_y sharded like the batch, adj_t / adj_params replicated (all-reduced derivatives);
    # the returned parameter gradient is the gradient of the WHOLE batch's loss on every rank
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(3)
            self.W1 = nn.Parameter(0.5 * torch.randn(3, 8, generator=g, dtype=torch.float64))
            self.W2 = nn.Parameter(0.5 * torch.randn(8, 3, generator=g, dtype=torch.float64))

        def forward(self, t, y):
            return torch.tanh(y @ self.W1) @ self.W2 - 0.1 * y
    net = Net().to(dev)
    ya = torch.tensor(y0[:1001], device=dev)
    ta = torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64)
    kwa = dict(rtol=1e-8, atol=1e-10, method="dopri5")
    yw = ya.clone().requires_grad_(True)
    out = tfd.odeint_adjoint(net, yw, ta, **kwa)
    (out[-1] ** 2).sum().backward()
    g_full = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    gy_full = yw.grad.clone()
    for p_ in net.parameters():
        p_.grad = None
    lo2, hi2 = shard_bounds(ya.shape[0], world, rank)
    ys = ya[lo2:hi2].clone().requires_grad_(True)
    out = tfd.odeint_adjoint(net, ys, ta, options={"shared_step_group": group}, **kwa)
    (out[-1] ** 2).sum().backward()
    g_part = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    res["adjoint_param_grad"] = (float((g_part - g_full).abs().max() / g_full.abs().max()), 0, 0, 0, 0)
    res["adjoint_input_grad"] = (float((ys.grad - gy_full[lo2:hi2]).abs().max() / gy_full.abs().max()), 0, 0, 0, 0)
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
