"""The shared list of parity cases.

``oracle/make_golden.py`` runs each case through the UNMODIFIED reference (over the tf shim) and stores
the result in ``tests/golden/<name>.npz``; ``tests/test_oracle_golden.py`` runs the numpy oracle on the
same inputs (CPU); ``tests/test_parity_gpu.py`` runs the CUDA engine on them (GPU).

A case = dict(name, problem, pkw, y0, t, dtype, method, rtol, atol, options, keep)
``y0`` / ``t`` are numpy arrays (y0 may be a tuple of arrays); ``keep`` is a stride applied to the output
time axis before it is stored (keeps the fixtures small; the last point is always kept).
"""
import numpy as np

from problems import PROBLEMS  # noqa: F401  (re-exported for users of this module)

F32 = np.float32


def _lin32(a, b, n):
    # the reference fixtures build `t` with tf.linspace(1., 8., n): float32 (tests/problems.py:78)
    return np.linspace(a, b, n, dtype=np.float64).astype(np.float32)


def _lorenz_y0(batch, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    return (np.array([1.0, 1.0, 1.0]) + 0.1 * rng.standard_normal((batch, 3))).astype(dtype)


def build_cases():
    C = []

    def add(name, problem, y0, t, method=None, rtol=1e-7, atol=1e-9, options=None, dtype="float64", pkw=None,
            keep=1, expect_error=None):
        C.append(dict(name=name, problem=problem, pkw=pkw or {}, y0=y0, t=np.asarray(t), dtype=dtype,
                      method=method, rtol=rtol, atol=atol, options=options, keep=keep,
                      expect_error=expect_error))

    t18 = _lin32(1., 8., 10)
    # --- reference tests/odeint_tests.py:25-109 (TestSolverError) on the `constant` problem -------------
    y0c = np.array(0.2 * float(t18[0]) + 3.0)
    for m in ["euler", "midpoint", "huen", "rk4", "bosh3", "adaptive_heun", "dopri5"]:
        add("constant_" + m, "constant", y0c, t18, method=m)
    # tsit5 is untested in the reference and its error estimate does not vanish with dt (the step size
    # collapses to ~1e-6 at default tolerances, millions of steps) -> only loose tolerances / short spans
    add("constant_tsit5", "constant", y0c, _lin32(1., 1.5, 6), method="tsit5", rtol=1e-2, atol=1e-2)
    add("constant_dopri8", "constant", y0c, t18, method="dopri8", rtol=1e-12, atol=1e-14)
    # reverse time (odeint_tests.py:112-171)
    for m in ["euler", "rk4", "dopri5", "dopri8", "adaptive_heun"]:
        add("constant_rev_" + m, "constant", np.array(0.2 * float(t18[-1]) + 3.0), t18[::-1].copy(), method=m)
    # zero-length integration (odeint_tests.py:174-210)
    for m in ["euler", "rk4", "dopri5", "tsit5"]:
        add("constant_len1_" + m, "constant", y0c, t18[0:1], method=m)
    # --- sine (bosh3 / adaptive_heun "never finish", odeint_tests.py:47-49) ------------------------------
    from problems import Sine
    y0s = np.array(Sine().exact(np.float64(t18[0])))
    add("sine_dopri5", "sine", y0s, t18, method="dopri5")
    add("sine_dopri8", "sine", y0s, t18, method="dopri8", rtol=1e-12, atol=1e-14)
    add("sine_tsit5", "sine", y0s, _lin32(1., 1.25, 6), method="tsit5", rtol=1e-2, atol=1e-2)
    add("sine_rev_dopri5", "sine", np.array(Sine().exact(np.float64(t18[-1]))), t18[::-1].copy(), method="dopri5")
    # --- linear: the reference's degenerate A == 0, and the intended skew-symmetric system ---------------
    for m in ["dopri5", "bosh3", "adaptive_heun"]:
        add("linear0_" + m, "linear", np.ones(10), t18, method=m, pkw=dict(degenerate=True))
    add("linear0_dopri8", "linear", np.ones(10), t18, method="dopri8", rtol=1e-12, atol=1e-14,
        pkw=dict(degenerate=True))
    add("linear_skew_dopri5", "linear", np.ones(10), t18, method="dopri5", pkw=dict(degenerate=False))
    add("linear_skew_dopri8", "linear", np.ones(10), t18, method="dopri8", rtol=1e-10, atol=1e-12,
        pkw=dict(degenerate=False))
    add("linear_skew_rk4", "linear", np.ones(10), _lin32(1., 8., 71), method="rk4", pkw=dict(degenerate=False))
    add("linear_skew_heun_f32", "linear", np.ones(10, dtype=F32), _lin32(1., 8., 141), method="heun",
        dtype="float32", pkw=dict(degenerate=False))
    # --- Lotka-Volterra (BASELINE config 1 and the README demo) ------------------------------------------
    tlv = _lin32(0., 10., 1000)
    add("lv_dopri5_default", "lv", np.array([1., 1.]), tlv, method=None, keep=37)
    add("lv_dopri5_cfg1", "lv", np.array([1., 1.]), tlv, method="dopri5", rtol=1e-6, atol=1e-9, keep=37)
    add("lv_tsit5", "lv", np.array([1., 1.]), _lin32(0., 0.5, 11), method="tsit5", rtol=1e-2, atol=1e-2)
    add("lv_dopri8", "lv", np.array([1., 1.]), _lin32(0., 10., 41), method="dopri8", rtol=1e-9, atol=1e-9)
    add("lv_adaptive_heun", "lv", np.array([1., 1.]), _lin32(0., 1., 5), method="adaptive_heun",
        rtol=1e-4, atol=1e-6)
    add("lv_opts", "lv", np.array([1., 1.]), _lin32(0., 10., 21), method="dopri5", rtol=1e-6, atol=1e-8,
        options=dict(first_step=0.01, safety=0.8, ifactor=5.0, dfactor=0.3))
    add("lv_unknown_opt", "lv", np.array([1., 1.]), _lin32(0., 1., 3), method="dopri5",
        options=dict(bogus=1), expect_error="UserWarning")
    add("lv_max_num_steps", "lv", np.array([1., 1.]), _lin32(0., 10., 3), method="dopri5",
        options=dict(max_num_steps=3), expect_error="AssertionError")
    add("lv_batched_f32", "lv", (np.array([[1., 1.], [1.2, 0.8], [0.7, 1.5], [2.0, 1.0]])).astype(F32),
        _lin32(0., 5., 26), method="dopri5", rtol=1e-3, atol=1e-4, dtype="float32")
    # --- y**3 spiral (examples/ode_demo.py) ---------------------------------------------------------------
    y0sp = np.array([[2., 0.]])
    add("spiral_rk4", "spiral", y0sp, _lin32(0., 25., 2001), method="rk4", keep=100)
    add("spiral_rk4_f32", "spiral", y0sp.astype(F32), _lin32(0., 25., 2001), method="rk4", dtype="float32", keep=100)
    add("spiral_euler", "spiral", y0sp, _lin32(0., 5., 501), method="euler", keep=50)
    add("spiral_midpoint", "spiral", y0sp, _lin32(0., 5., 201), method="midpoint", keep=20)
    add("spiral_dopri5", "spiral", y0sp, _lin32(0., 25., 1000), method="dopri5", keep=37)
    add("spiral_dopri8", "spiral", y0sp, _lin32(0., 25., 101), method="dopri8", rtol=1e-9, atol=1e-9, keep=5)
    add("spiral_dopri5_f32", "spiral", y0sp.astype(F32), _lin32(0., 25., 101), method="dopri5", rtol=1e-3,
        atol=1e-3, dtype="float32", keep=5)
    add("spiral_rev_dopri5", "spiral", np.array([[0.5, 0.1]]), _lin32(1., 0., 11), method="dopri5")
    add("spiral_mlp_rk4_f32", "spiral_mlp", (np.array([[2., 0.], [1.5, 0.5], [1.0, -1.0]])).astype(F32),
        _lin32(0., 5., 101), method="rk4", dtype="float32", keep=10)
    # --- tuple state of unequal shapes (segmented error norm; accept iff every component passes) ----------
    add("tuple_dopri5", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 1., 3), method="dopri5")
    add("tuple_dopri8", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 1., 3), method="dopri8", rtol=1e-9, atol=1e-11)
    add("tuple_adaptive_heun", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 0.2, 3), method="adaptive_heun", rtol=1e-4, atol=1e-6)
    add("tuple_tsit5", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 0.1, 3), method="tsit5", rtol=1e-2, atol=1e-2)
    add("tuple_rk4", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 1., 201), method="rk4", keep=20)
    # --- batched Lorenz: ONE shared step and a GLOBAL scalar tolerance across the batch (BASELINE cfg 2) --
    add("lorenz_b16_dopri5", "lorenz", _lorenz_y0(16), np.arange(101) * 0.01, method="dopri5", keep=10)
    add("lorenz_b64_dopri5_f32", "lorenz", _lorenz_y0(64, dtype=F32), np.arange(51) * 0.01, method="dopri5",
        rtol=1e-4, atol=1e-5, dtype="float32", keep=10)
    add("lorenz_b16_dopri8", "lorenz", _lorenz_y0(16), np.arange(21) * 0.05, method="dopri8", rtol=1e-9, atol=1e-9)
    add("lorenz_b16_tsit5", "lorenz", _lorenz_y0(16), np.arange(11) * 0.005, method="tsit5", rtol=1e-2, atol=1e-2)
    # --- DETEST-style stiff-ish tridiagonal, dim 16, reject stress (BASELINE cfg 5 in miniature) ----------
    rng = np.random.default_rng(1)
    y0tri = np.zeros((8, 16)); y0tri[:, 0] = 1.0; y0tri += 0.01 * rng.standard_normal((8, 16))
    add("tridiag_dopri8", "tridiag", y0tri, np.linspace(0., 20., 11), method="dopri8", rtol=1e-9, atol=1e-9)
    add("tridiag_dopri5", "tridiag", y0tri, np.linspace(0., 20., 11), method="dopri5", rtol=1e-6, atol=1e-6)
    # --- multistep solvers (fixed_adams.py, adams.py) -----------------------------------------
    lv0 = np.array([1., 1.])
    add("lv_fixed_adams", "lv", lv0, np.linspace(0., 2., 201), method="fixed_adams", rtol=1e-6, atol=1e-8, keep=20)
    # (at the default max_order = 12 the 11-step Bashforth formula is unstable on this problem even at dt = 0.005 -- the
    #  reference ends in NaN -- so the explicit cases cap the order or use the sine problem)
    add("lv_explicit_adams", "lv", lv0, np.linspace(0., 2., 401), method="explicit_adams", options=dict(max_order=6), keep=40)
    add("lv_fixed_adams_order5_iter1", "lv", lv0, np.linspace(0., 2., 201), method="fixed_adams", rtol=1e-6, atol=1e-8,
        options=dict(max_order=5, max_iters=1), keep=20)
    add("lv_fixed_adams_f32", "lv", lv0.astype(F32), np.linspace(0., 2., 201), method="fixed_adams", rtol=1e-4, atol=1e-6,
        dtype="float32", keep=20)
    add("lv_rev_fixed_adams", "lv", lv0, np.linspace(2., 0., 101), method="fixed_adams", rtol=1e-6, atol=1e-8, keep=10)
    add("sine_explicit_adams", "sine", np.array([1.0, 0.5]), np.linspace(1., 3., 401), method="explicit_adams", keep=40)
    add("tuple_fixed_adams", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 1., 401), method="fixed_adams", rtol=1e-6, atol=1e-8, keep=40)
    add("lorenz_b16_explicit_adams_step", "lorenz", _lorenz_y0(16), np.arange(11) * 0.05, method="explicit_adams",
        options=dict(max_order=6))
    add("lv_adams", "lv", lv0, np.linspace(0., 2., 201), method="adams", rtol=1e-6, atol=1e-8, keep=20)
    add("lv_adams_order4", "lv", lv0, np.linspace(0., 2., 201), method="adams", rtol=1e-6, atol=1e-8,
        options=dict(max_order=4), keep=20)
    add("lv_rev_adams", "lv", lv0, np.linspace(2., 0., 101), method="adams", rtol=1e-6, atol=1e-8, keep=10)
    add("lv_adams_default_tol", "lv", lv0, np.linspace(0., 1., 11), method="adams")
    add("tuple_adams", "tuple_decay", (np.linspace(1., 2., 2), np.linspace(0.5, 1.5, 5)),
        np.linspace(0., 0.5, 6), method="adams", rtol=1e-5, atol=1e-7)
    add("lorenz_b16_adams", "lorenz", _lorenz_y0(16), np.arange(11) * 0.05, method="adams", rtol=1e-6, atol=1e-8)
    add("lv_adams_unknown_opt", "lv", lv0, np.linspace(0., 0.2, 3), method="adams", rtol=1e-5, atol=1e-7,
        options=dict(bogus=1), expect_error="UserWarning")
    # --- the reference's DETEST benchmark (tests/DETEST/run.py: [0, 20], dopri5 and adams, tol = rtol = atol) ----------
    from detest_problems import make as _detest
    t20 = np.array([0.0, 5.0, 20.0])
    for nm, method, tol in [("A3", "dopri5", 1e-6), ("B1", "dopri5", 1e-6), ("B4", "dopri5", 1e-9), ("C3", "dopri5", 1e-6),
                            ("C4", "dopri8", 1e-9), ("D3", "dopri5", 1e-6), ("D5", "dopri5", 1e-3), ("E2", "dopri5", 1e-6),
                            ("E5", "dopri5", 1e-6), ("B1", "adams", 1e-6), ("C3", "adams", 1e-6), ("E2", "adams", 1e-3),
                            ("A4", "adams", 1e-6)]:
        add("detest_%s_%s" % (nm, method), "detest", np.asarray(_detest(nm)[1]), t20, method=method, rtol=tol, atol=tol,
            pkw=dict(name=nm))
    return C


CASES = build_cases()
CASES_BY_NAME = {c["name"]: c for c in CASES}
