"""GPU: built-in right-hand sides (tfdiffeq_b200/rhs.py) solved by the single persistent kernel
(b2ode_fused_solve) against (1) the generic path with the same module as an ordinary func, (2) the numpy
oracle, (3) the reference's golden vectors."""
import numpy as np
import pytest
import torch

import np_ref
from cases import CASES_BY_NAME
from golden_util import load_golden, max_rel_err
from problems import PROBLEMS

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def tfd():
    import tfdiffeq_b200
    return tfdiffeq_b200


def _both(f, y0, t, method, opts=None, **kw):
    a = tfd().odeint(f, y0, t, method=method, options=dict(opts or {}), **kw)
    sa = dict(tfd().last_stats)
    b = tfd().odeint(f, y0, t, method=method, options=dict(opts or {}, fused_rhs=False), **kw)
    sb = dict(tfd().last_stats)
    assert sa["fused_rhs"] and not sb["fused_rhs"]
    return a, sa, b, sb


@pytest.mark.parametrize("method,kw,opts,dt", [
    ("dopri5", {}, None, 0.01), ("dopri8", dict(rtol=1e-9, atol=1e-9), None, 0.01),
    ("bosh3", dict(rtol=1e-5, atol=1e-7), dict(textbook_tableau=True), 0.002),
    ("adaptive_heun", dict(rtol=1e-3, atol=1e-5), None, 0.001),
    ("dopri5", dict(rtol=1e-6, atol=1e-8), dict(first_step=0.01, safety=0.8, ifactor=5.0, dfactor=0.3), 0.01)])
@pytest.mark.parametrize("batch", [1, 127, 4099])
def test_fused_lorenz_matches_generic_path_fp64(method, kw, opts, dt, batch):
    rng = np.random.default_rng(batch)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((batch, 3)), device=DEV)
    t = torch.arange(41, dtype=torch.float64) * dt
    a, sa, b, sb = _both(tfd().rhs.Lorenz(), y0, t, method, opts, **kw)
    assert (sa["n_accepted"], sa["n_rejected"], sa["nfe"]) == (sb["n_accepted"], sb["n_rejected"], sb["nfe"])
    # identical arithmetic per stage; only the summation order of the error norm differs
    assert float((a - b).abs().max()) <= 1e-9 * max(1.0, float(b.abs().max()))


def test_fused_fp32_lotka_volterra_and_reverse_time():
    rng = np.random.default_rng(2)
    y0 = torch.tensor(1.0 + 0.3 * rng.random((513, 2)), dtype=torch.float32, device=DEV)
    f = tfd().rhs.LotkaVolterra()
    t = torch.linspace(0., 3., 31)
    a, sa, b, sb = _both(f, y0, t, "dopri5", rtol=1e-4, atol=1e-5)
    assert abs(sa["n_accepted"] - sb["n_accepted"]) <= 1 and abs(sa["n_rejected"] - sb["n_rejected"]) <= 1
    assert float((a - b).abs().max()) <= 1e-3
    y64 = y0.double()
    tr = torch.linspace(2., 0., 21, dtype=torch.float64)
    a, sa, b, sb = _both(f, y64, tr, "dopri5")
    assert (sa["n_accepted"], sa["n_rejected"]) == (sb["n_accepted"], sb["n_rejected"])
    assert float((a - b).abs().max()) <= 1e-9


@pytest.mark.parametrize("name", ["lorenz_b16_dopri5", "lorenz_b16_dopri8", "lv_dopri5_cfg1", "lv_dopri8", "lv_opts",
                                  "lorenz_b64_dopri5_f32", "lv_batched_f32", "lv_adaptive_heun"])
def test_fused_matches_reference_golden(name):
    c = CASES_BY_NAME[name]
    g = load_golden(name)
    f = tfd().rhs.Lorenz() if c["problem"] == "lorenz" else tfd().rhs.LotkaVolterra()
    dt = torch.float64 if c["dtype"] == "float64" else torch.float32
    y0 = torch.tensor(np.asarray(c["y0"]), dtype=dt, device=DEV)
    kw = dict(rtol=c["rtol"], atol=c["atol"], method=c["method"] or "dopri5")
    if c["options"] is not None:
        kw["options"] = c["options"]
    sol = tfd().odeint(f, y0, torch.from_numpy(np.ascontiguousarray(c["t"])), **kw)
    st = dict(tfd().last_stats)
    assert st["fused_rhs"]
    tol = 1e-6 if c["dtype"] == "float64" else 1e-3
    assert max_rel_err(sol.cpu().numpy()[g["idx"]], g["sol0"]) <= tol
    if c["dtype"] == "float64":
        assert (st["n_accepted"], st["n_rejected"], st["nfe"]) == (g["n_acc"], g["n_rej"], g["nfe"])


def test_fused_full_size_vs_oracle_short_horizon():
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((65536, 3))
    t = np.arange(11) * 0.01
    st = np_ref.Stats()
    ref = np_ref.odeint(PROBLEMS["lorenz"](), y0, t, method="dopri5", stats=st)
    got = tfd().odeint(tfd().rhs.Lorenz(), torch.tensor(y0, device=DEV), torch.tensor(t), method="dopri5")
    s = dict(tfd().last_stats)
    assert s["fused_rhs"]
    assert max_rel_err(got.cpu().numpy(), ref) <= 1e-6
    assert (s["n_accepted"], s["n_rejected"], s["nfe"]) == (st.n_acc, st.n_rej, st.nfe)


def test_fused_status_bits_and_fallbacks():
    f = tfd().rhs.Lorenz()
    y0 = torch.ones(8, 3, dtype=torch.float64, device=DEV)
    t = torch.tensor([0., 10.])
    with pytest.raises(AssertionError, match="max_num_steps exceeded"):
        tfd().odeint(f, y0, t, method="dopri5", options=dict(max_num_steps=3))
    bad = y0.clone()
    bad[2, 1] = float("nan")
    with pytest.raises(AssertionError, match="non-finite values in state"):
        tfd().odeint(f, bad, t, method="dopri5", options=dict(first_step=0.01))
    # a state whose last axis is not the system dimension, a tuple state, tsit5 -> generic path, same API
    out = tfd().odeint(lambda t, y: (f(t, y[0]),), (y0,), torch.tensor([0., 0.1]), method="dopri5")
    assert isinstance(out, tuple) and not tfd().last_stats["fused_rhs"]
    tfd().odeint(f, y0, torch.tensor([0., 0.001]), method="tsit5", rtol=1e-2, atol=1e-2)
    assert not tfd().last_stats["fused_rhs"]
    # a batch too large to stay co-resident silently takes the generic path
    big = torch.ones(300000, 3, dtype=torch.float64, device=DEV)
    tfd().odeint(f, big, torch.tensor([0., 0.01]), method="dopri5")
    assert not tfd().last_stats["fused_rhs"]


# --------------------------------------------------------------------------------------------------
# fixed-grid methods with a built-in right-hand side: one launch, no reductions
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["euler", "midpoint", "heun", "rk4"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_fixed_grid_is_bit_identical_to_generic_path(method, dtype):
    """Same IEEE operations in the same order, and no reduction anywhere: the two paths must agree exactly."""
    rng = np.random.default_rng(4)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((1000, 3)), dtype=dtype, device=DEV)
    t = torch.linspace(0., 0.5, 51)
    f = tfd().rhs.Lorenz()
    a = tfd().odeint(f, y0, t, method=method)
    assert tfd().last_stats["fused_rhs"]
    nfe_a = tfd().last_stats["nfe"]
    b = tfd().odeint(f, y0, t, method=method, options=dict(fused_rhs=False))
    assert not tfd().last_stats["fused_rhs"] and tfd().last_stats["nfe"] == nfe_a
    assert torch.equal(a, b)
    # reverse time and a finer internal grid with interpolated outputs
    tr = torch.tensor([0.1, 0.07, 0.03, 0.0])
    a = tfd().odeint(f, y0, tr, method=method, options=dict(step_size=0.013))
    b = tfd().odeint(f, y0, tr, method=method, options=dict(step_size=0.013, fused_rhs=False))
    assert bool(torch.isfinite(b).all()) and torch.equal(a, b)


def test_cubic_mlp_builtin_fixed_and_adaptive_vs_generic():
    g = torch.Generator().manual_seed(0)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-3)):
        m = tfd().rhs.CubicMLP(hidden=50, dtype=dtype, generator=g).to(DEV)
        rng = np.random.default_rng(9)
        y0 = torch.tensor(np.array([2., 0.]) + 0.1 * rng.standard_normal((777, 2)), dtype=dtype, device=DEV)
        t = torch.linspace(0., 2., 81)
        a = tfd().odeint(m, y0, t, method="rk4")
        assert tfd().last_stats["fused_rhs"]
        b = tfd().odeint(m, y0, t, method="rk4", options=dict(fused_rhs=False))
        # torch evaluates the two matrix products with cuBLAS (its own FMA order): agreement to rounding
        assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))
        kw = dict(rtol=1e-6, atol=1e-8) if dtype == torch.float64 else dict(rtol=1e-3, atol=1e-4)
        a = tfd().odeint(m, y0, t[:21], method="dopri5", **kw)
        sa = dict(tfd().last_stats)
        b = tfd().odeint(m, y0, t[:21], method="dopri5", options=dict(fused_rhs=False), **kw)
        sb = dict(tfd().last_stats)
        assert sa["fused_rhs"] and not sb["fused_rhs"]
        assert abs(sa["n_accepted"] - sb["n_accepted"]) <= 1 and abs(sa["n_rejected"] - sb["n_rejected"]) <= 1
        assert float((a - b).abs().max()) <= max(tol, 1e-6) * max(1.0, float(b.abs().max()))


def test_config3_full_size_fused_mlp_rk4():
    """BASELINE config 3 at full size through the built-in module: 131 072 x 2 fp32, rk4, 2 000 grid cells, ONE
    kernel launch; a random subset is checked against the oracle (fixed grids have no coupling between
    trajectories)."""
    g = torch.Generator().manual_seed(1)
    m = tfd().rhs.CubicMLP(hidden=50, dtype=torch.float32, generator=g).to(DEV)
    rng = np.random.default_rng(3)
    y0 = (np.array([2., 0.]) + 0.1 * rng.standard_normal((131072, 2))).astype(np.float32)
    t = np.linspace(0., 25., 2001).astype(np.float32)
    sol = tfd().odeint(m, torch.tensor(y0, device=DEV), torch.tensor(t), method="rk4")
    s = dict(tfd().last_stats)
    assert s["fused_rhs"] and s["nfe"] == 8000 and sol.shape == (2001, 131072, 2)
    W1, b1, W2, b2 = (p.detach().cpu().numpy() for p in (m.W1, m.b1, m.W2, m.b2))
    f_np = lambda tt, y: np.tanh((y ** 3) @ W1 + b1) @ W2 + b2               # noqa: E731
    idx = rng.choice(131072, size=32, replace=False)
    ref = np_ref.odeint(f_np, y0[idx], t, method="rk4")
    got = sol[:, torch.tensor(idx, device=DEV)].cpu().numpy()
    assert max_rel_err(got, ref) <= 1e-3


@pytest.mark.parametrize("batch", [300, 65536])
def test_host_output_streams_the_solution_behind_the_solve(batch):
    """options={'host_output': pinned}: the solution is delivered into the caller's page-locked buffer; with a built-in
    right-hand side the device-to-host copies are issued while the persistent kernel is still running, gated by the row
    watermark the kernel publishes (rows [0, mark) complete on every block).  Values must equal the ordinary result."""
    import tfdiffeq_b200 as tfd
    rng = np.random.default_rng(11)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((batch, 3)), device=DEV)
    t = torch.arange(400, dtype=torch.float64) * 0.01
    f = tfd.rhs.Lorenz()
    want = tfd.odeint(f, y0, t, method="dopri5")
    host = torch.empty((400, batch, 3), dtype=torch.float64).pin_memory()
    for rep in range(3):
        host.fill_(float("nan"))
        got = tfd.odeint(f, y0, t, method="dopri5", options={"host_output": host})
        assert got.data_ptr() == host.data_ptr() and tfd.last_stats["fused_rhs"]
        assert torch.equal(got, want.cpu())
    # generic path and fixed grid: delivered with one copy at the end
    g = PROBLEMS["lorenz"](backend="torch", device=DEV)
    host.fill_(float("nan"))
    got = tfd.odeint(g, y0, t, method="dopri5", options={"host_output": host})
    assert float((got - want.cpu()).abs().max()) < 1e-9
    got = tfd.odeint(g, y0, t, method="rk4", options={"host_output": host})
    assert torch.equal(got, tfd.odeint(g, y0, t, method="rk4").cpu())
    with pytest.raises(ValueError):
        tfd.odeint(f, y0, t, method="dopri5", options={"host_output": torch.empty((400, batch, 3), dtype=torch.float64)})
