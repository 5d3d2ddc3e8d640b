"""CPU: pin the numpy oracle (oracle/np_ref.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py), and against the reference tests' analytic solutions."""
import warnings

import numpy as np
import pytest

import np_ref
from cases import CASES
from golden_util import dt_trace_rtol, load_golden, max_rel_err, tolerances
from problems import PROBLEMS


def run_oracle(c, stats=None):
    prob = PROBLEMS[c["problem"]](backend="numpy", dtype=c["dtype"], **c["pkw"])
    dt = np.dtype(c["dtype"])
    y0 = c["y0"]
    y0 = tuple(np.asarray(v).astype(dt) for v in y0) if isinstance(y0, tuple) else np.asarray(y0).astype(dt)
    kw = dict(rtol=c["rtol"], atol=c["atol"])
    if c["method"] is not None:
        kw["method"] = c["method"]
    if c["options"] is not None:
        kw["options"] = c["options"]
    return np_ref.odeint(prob, y0, c["t"], stats=stats, **kw)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    g = load_golden(case["name"])
    st = np_ref.Stats()
    if case["expect_error"] == "AssertionError":
        assert g["error"].startswith("AssertionError")
        with pytest.raises(AssertionError, match="max_num_steps exceeded"):
            run_oracle(case, st)
        return
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sol = run_oracle(case, st)
    if case["expect_error"] == "UserWarning":
        assert "Unexpected arguments" in g["warned"]
        assert any("Unexpected arguments" in str(x.message) for x in w)
    assert g["error"] == ""
    sols = sol if isinstance(sol, tuple) else (sol,)
    assert len(sols) == int(g["ncomp"])
    tol = tolerances(case)["oracle"]
    for i, s in enumerate(sols):
        ref = g["sol%d" % i]
        assert s.dtype == ref.dtype
        assert s[g["idx"]].shape == ref.shape
        assert max_rel_err(s[g["idx"]], ref) <= tol, (case["name"], max_rel_err(s[g["idx"]], ref))
    # identical control flow: accepted / rejected attempts and function evaluations
    assert (st.n_acc, st.n_rej, st.nfe) == (g["n_acc"], g["n_rej"], g["nfe"])
    if len(g["dt_trace"]):
        n = min(len(g["dt_trace"]), len(st.dt_trace))
        np.testing.assert_allclose(np.array(st.dt_trace[:n]), g["dt_trace"][:n],
                                   rtol=dt_trace_rtol(case))


@pytest.mark.parametrize("method,kw", [("euler", {}), ("midpoint", {}), ("huen", {}), ("rk4", {}),
                                       ("bosh3", {}), ("adaptive_heun", {}), ("dopri5", {}),
                                       ("dopri8", dict(rtol=1e-12, atol=1e-14))])
def test_oracle_reference_test_tolerance_constant(method, kw):
    """reference tests/odeint_tests.py:25-109: max |(true - est)/true| < 1e-4 on `constant`."""
    f = PROBLEMS["constant"]()
    t = np.linspace(1., 8., 10).astype(np.float32)
    y = np_ref.odeint(f, f.y0(t[0]), t, method=method, **kw)
    true = f.exact(t)
    assert np.max(np.abs((true - y) / true)) < 1e-4


@pytest.mark.parametrize("method,kw", [("dopri5", {}), ("dopri8", dict(rtol=1e-12, atol=1e-14))])
def test_oracle_reference_test_tolerance_sine(method, kw):
    f = PROBLEMS["sine"]()
    t = np.linspace(1., 8., 10).astype(np.float32)
    y = np_ref.odeint(f, f.y0(t[0]), t, method=method, **kw)
    true = f.exact(t)
    assert np.max(np.abs((true - y) / true)) < 1e-4


def test_oracle_errors():
    f = PROBLEMS["lv"]()
    with pytest.raises(ValueError):
        np_ref.odeint(f, np.array([1., 1.]), np.array([0., 1.]), options=dict(first_step=0.1))
    with pytest.raises(KeyError):
        np_ref.odeint(f, np.array([1., 1.]), np.array([0., 1.]), method="nope")


def test_oracle_torch_backend_matches_numpy():
    """bench.py times the oracle on torch-CPU tensors (all host threads); same arithmetic, same counts."""
    import torch
    fn, ft = PROBLEMS["lorenz"](backend="numpy"), PROBLEMS["lorenz"](backend="torch")
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((32, 3))
    t = np.arange(21) * 0.01
    s1, s2 = np_ref.Stats(), np_ref.Stats()
    a = np_ref.odeint(fn, y0, t, stats=s1)
    b = np_ref.odeint(ft, torch.from_numpy(y0), t, stats=s2)
    assert (s1.n_acc, s1.n_rej, s1.nfe) == (s2.n_acc, s2.n_rej, s2.nfe)
    assert np.max(np.abs(a - b.numpy())) < 1e-10


def test_gradient_fixtures_are_self_consistent(golden_dir):
    """tests/golden/grad_*.npz hold the UNMODIFIED reference adjoint's gradients next to gradients back-propagated
    through this oracle's discrete solver (oracle/make_golden_grads.py).  Two independent routes to the same derivative
    must agree to the solver tolerance (the dopri8 dense output is 4th order, rk4 is a fixed grid: looser)."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(golden_dir, "grad_*.npz")))
    assert len(files) >= 7
    loose = {"spiral3_dopri8": 2e-4, "spiral3_rk4": 2e-2, "mlp_tanh_f32": 5e-3, "mlp_tanh_dopri5": 5e-5}
    for f in files:
        name = os.path.basename(f)[5:-4]
        g = np.load(f)
        tol = loose.get(name, 2e-6)
        for k in g.files:
            if k.startswith("bp_"):
                a, b = g[k[3:]], g[k]
                assert a.shape == b.shape
                assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), (name, k)
        assert np.all(np.isfinite(g["g_t"])) and g["g_t"].shape == g["t"].shape
