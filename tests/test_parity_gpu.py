"""GPU parity tests: the CUDA engine (through the public odeint API, i.e. through the C ABI) against
(1) the golden vectors produced by the unmodified reference, (2) the numpy oracle on the same seeded inputs,
(3) size-independent properties at BASELINE.json's full sizes.

Tolerances are the ones BASELINE.json.north_star states: max-abs 1e-6 (fp64) / 1e-3 (fp32) relative to
max(1, max|y|), plus identical accepted / rejected / NFE counts.
"""
import warnings

import numpy as np
import pytest
import torch

import np_ref
from cases import CASES
from golden_util import load_golden, max_rel_err, tolerances
from problems import PROBLEMS

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
TD = {"float32": torch.float32, "float64": torch.float64}


def tfd():
    import tfdiffeq_b200
    return tfdiffeq_b200


def run_engine(c):
    prob = PROBLEMS[c["problem"]](backend="torch", dtype=c["dtype"], device=DEV, **c["pkw"])
    y0 = c["y0"]
    if isinstance(y0, tuple):
        y0 = tuple(torch.tensor(np.asarray(v), dtype=TD[c["dtype"]], device=DEV) for v in y0)
    else:
        y0 = torch.tensor(np.asarray(y0), dtype=TD[c["dtype"]], device=DEV)
    t = torch.from_numpy(np.ascontiguousarray(c["t"]))
    kw = dict(rtol=c["rtol"], atol=c["atol"])
    if c["method"] is not None:
        kw["method"] = c["method"]
    if c["options"] is not None:
        kw["options"] = c["options"]
    sol = tfd().odeint(prob, y0, t, **kw)
    return sol, dict(tfd().last_stats)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_engine_matches_reference_golden(case):
    g = load_golden(case["name"])
    if case["expect_error"] == "AssertionError":
        with pytest.raises(AssertionError, match="max_num_steps exceeded"):
            run_engine(case)
        return
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sol, stats = run_engine(case)
    if case["expect_error"] == "UserWarning":
        assert any(g["warned"] == str(x.message) for x in w), [str(x.message) for x in w]
    sols = sol if isinstance(sol, tuple) else (sol,)
    assert len(sols) == int(g["ncomp"])
    tol = tolerances(case)["engine"]
    for i, s in enumerate(sols):
        ref = g["sol%d" % i]
        got = s.cpu().numpy()
        assert got.dtype == ref.dtype
        assert got.shape[1:] == ref.shape[1:] and got.shape[0] == len(case["t"])
        assert max_rel_err(got[g["idx"]], ref) <= tol, (case["name"], max_rel_err(got[g["idx"]], ref))
    if g["n_acc"] + g["n_rej"] > 0:
        counts = (stats["n_accepted"], stats["n_rejected"], stats["nfe"])
        want = (g["n_acc"], g["n_rej"], g["nfe"])
        if case["rtol"] >= 1e-9 and case["dtype"] == "float64":
            assert counts == want
        else:
            # at rtol <= 1e-10 (and in fp32) the error estimate is rounding noise: a borderline accept may flip
            assert abs(counts[0] - want[0]) <= 2 and abs(counts[1] - want[1]) <= 2
    else:
        assert stats["nfe"] == g["nfe"]


# --------------------------------------------------------------------------------------------------
# engine vs oracle on seeded inputs at sizes the oracle finishes in seconds
# --------------------------------------------------------------------------------------------------
def _oracle_vs_engine(problem, y0, t, dtype, pkw=None, **kw):
    pkw = pkw or {}
    fn = PROBLEMS[problem](backend="numpy", dtype=dtype, **pkw)
    ft = PROBLEMS[problem](backend="torch", dtype=dtype, device=DEV, **pkw)
    st = np_ref.Stats()
    ref = np_ref.odeint(fn, y0, t, stats=st, **kw)
    got = tfd().odeint(ft, torch.tensor(y0, device=DEV), torch.tensor(t), **kw)
    stats = dict(tfd().last_stats)
    return ref, got.cpu().numpy(), st, stats


@pytest.mark.parametrize("method,rtol,atol", [("dopri5", 1e-7, 1e-9), ("dopri8", 1e-9, 1e-9), ("bosh3", 1e-4, 1e-6),
                                              ("adaptive_heun", 1e-3, 1e-5), ("tsit5", 1e-2, 1e-2)])
def test_lorenz_4096_fp64_vs_oracle(method, rtol, atol):
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((4096, 3))
    t = np.arange(41) * 0.01 if method not in ("adaptive_heun", "tsit5", "bosh3") else np.arange(11) * 0.002
    opts = dict(textbook_tableau=True) if method == "bosh3" else None
    kw = dict(method=method, rtol=rtol, atol=atol)
    fn = PROBLEMS["lorenz"](backend="numpy")
    ft = PROBLEMS["lorenz"](backend="torch", device=DEV)
    st = np_ref.Stats()
    ref = np_ref.odeint(fn, y0, t, stats=st, method="bosh3_textbook" if opts else method, rtol=rtol, atol=atol)
    got = tfd().odeint(ft, torch.tensor(y0, device=DEV), torch.tensor(t), options=opts, **kw).cpu().numpy()
    stats = tfd().last_stats
    assert max_rel_err(got, ref) <= 1e-6
    assert (stats["n_accepted"], stats["n_rejected"], stats["nfe"]) == (st.n_acc, st.n_rej, st.nfe)


def test_tridiag_256x128_dopri8_reject_stress_vs_oracle():
    """BASELINE config 5 in miniature: dim 128, rtol 1e-9, dopri8."""
    rng = np.random.default_rng(3)
    y0 = np.zeros((256, 128))
    y0[:, 0] = 1.0
    y0 += 0.01 * rng.standard_normal(y0.shape)
    ref, got, st, stats = _oracle_vs_engine("tridiag", y0, np.linspace(0., 2., 5), "float64", pkw=dict(dim=128),
                                            method="dopri8", rtol=1e-9, atol=1e-9)
    assert max_rel_err(got, ref) <= 1e-6
    assert abs(stats["n_accepted"] - st.n_acc) <= 1 and abs(stats["n_rejected"] - st.n_rej) <= 1


def test_spiral_mlp_rk4_fp32_vs_oracle():
    """BASELINE config 3 shape (2 -> 50 -> 2 MLP, rk4, fp32) at a batch the oracle finishes quickly."""
    rng = np.random.default_rng(5)
    y0 = (np.array([2., 0.]) + 0.1 * rng.standard_normal((8192, 2))).astype(np.float32)
    ref, got, st, stats = _oracle_vs_engine("spiral_mlp", y0, np.linspace(0., 2.5, 201).astype(np.float32),
                                            "float32", method="rk4")
    assert max_rel_err(got, ref) <= 1e-3
    assert stats["nfe"] == st.nfe == 800


@pytest.mark.parametrize("method", ["euler", "midpoint", "heun", "rk4"])
def test_fixed_grid_bit_exact_vs_oracle_fp64(method):
    """Fixed-grid steppers have no reductions: with an elementwise func the engine must equal the oracle
    bit for bit (same IEEE operations in the same order)."""
    rng = np.random.default_rng(7)
    y0 = rng.standard_normal((1000, 3))

    class Cubic(object):
        def __call__(self, t, y):
            return -(y * y * y) * 0.5 + y * 0.25
    t = np.linspace(0., 1., 33)
    ref = np_ref.odeint(Cubic(), y0, t, method=method)
    got = tfd().odeint(Cubic(), torch.tensor(y0, device=DEV), torch.tensor(t), method=method).cpu().numpy()
    assert np.array_equal(got, ref)


def test_fixed_grid_step_size_option_and_interior_outputs():
    """`step_size` builds a finer internal grid; outputs strictly inside a cell are linearly interpolated
    (solvers.py:106-115)."""
    rng = np.random.default_rng(8)
    y0 = rng.standard_normal((64, 5))
    f = lambda t, y: -y                                            # noqa: E731
    t = np.array([0., 0.33, 0.5, 1.0])
    ref = np_ref.odeint(f, y0, t, method="rk4", options=dict(step_size=0.125))
    got = tfd().odeint(f, torch.tensor(y0, device=DEV), torch.tensor(t), method="rk4",
                       options=dict(step_size=0.125)).cpu().numpy()
    assert np.array_equal(got, ref)
    with pytest.raises(ValueError):
        tfd().odeint(f, torch.tensor(y0, device=DEV), torch.tensor(t), method="rk4",
                     options=dict(step_size=0.1, grid_constructor=lambda f, y, t: t))


# --------------------------------------------------------------------------------------------------
# kernel-level bit-exactness through the C ABI: one Dopri5 attempt on fixed inputs
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 127, 4096 + 3])
def test_single_attempt_bit_exact(dtype, n):
    """With first_step given there is no reduction before the first attempt's outputs: the stage inputs the
    func sees, and the dense output written for an accepted first step, must be bit-identical to the oracle
    (explicit mul/add, no FMA contraction, reference operation order).  Odd n exercises the scalar tail."""
    rng = np.random.default_rng(n)
    npdt = np.dtype(dtype)
    y0 = rng.standard_normal(n).astype(npdt)
    seen_e, seen_o = [], []

    def f_engine(t, y):
        seen_e.append(y.detach().cpu().numpy().copy())
        return torch.sin(y) * 0.5

    def f_oracle(t, y):
        seen_o.append(np.array(y, copy=True))
        # torch CPU sin so both sides evaluate the same libm-free kernel family? no: use identical values instead
        return (torch.sin(torch.from_numpy(np.ascontiguousarray(y)).to(DEV)) * 0.5).cpu().numpy()

    t = np.array([0., 0.05, 0.1])
    kw = dict(method="dopri5", rtol=1e-3, atol=1e-3, options=dict(first_step=0.1))
    ref = np_ref.odeint(f_oracle, y0, t, **kw)
    got = tfd().odeint(f_engine, torch.tensor(y0, device=DEV), torch.tensor(t), **kw).cpu().numpy()
    m = min(len(seen_e), len(seen_o), 7)       # f0 + the six stage inputs of the first attempt
    assert m == 7
    for a, b in zip(seen_e[:m], seen_o[:m]):
        assert np.array_equal(a.reshape(-1), b.reshape(-1))
    assert np.array_equal(got, ref)


# --------------------------------------------------------------------------------------------------
# edge cases
# --------------------------------------------------------------------------------------------------
def test_func_returning_its_input_or_a_reused_buffer():
    y0 = torch.linspace(0.5, 1.5, 37, dtype=torch.float64, device=DEV)
    t = torch.linspace(0., 1., 5)
    a = tfd().odeint(lambda t, y: y, y0, t, method="dopri5")                      # returns its input (aliases)
    buf = torch.empty_like(y0)

    def reuse(t, y):
        buf.copy_(y)
        return buf                                                                  # same storage every call
    b = tfd().odeint(reuse, y0, t, method="dopri5")
    exact = (y0.cpu().numpy()[None, :] * np.exp(np.linspace(0., 1., 5))[:, None])
    assert np.max(np.abs(a.cpu().numpy() - exact)) < 1e-6
    assert torch.equal(a, b)


def test_misaligned_func_output_takes_scalar_path():
    y0 = torch.linspace(0.5, 1.5, 64, dtype=torch.float64, device=DEV)
    t = torch.linspace(0., 1., 5)
    big = torch.empty(65, dtype=torch.float64, device=DEV)

    def mis(t, y):
        out = torch.empty(65, dtype=torch.float64, device=DEV)[1:]                  # 8-byte, not 16-byte aligned
        out.copy_(-y)
        return out
    a = tfd().odeint(mis, y0, t, method="dopri5")
    b = tfd().odeint(lambda t, y: -y, y0, t, method="dopri5")
    assert torch.equal(a, b)
    del big


def test_nonfinite_state_and_underflow_raise_like_the_reference():
    y0 = torch.ones(8, dtype=torch.float64, device=DEV)
    t = torch.tensor([0., 1.])
    bad = y0.clone()
    bad[3] = float("inf")
    # with a given first step the first assertion that can fail is the finite check (dopri5.py:100) ...
    with pytest.raises(AssertionError, match="non-finite values in state"):
        tfd().odeint(lambda t, y: -y, bad, t, method="dopri5", options=dict(first_step=0.1))
    # ... without it, the initial-step heuristic already produces dt = NaN and the reference's FIRST assertion,
    # `t0 + dt > t0` (dopri5.py:98), is the one that fires
    with pytest.raises(AssertionError, match="underflow in dt"):
        tfd().odeint(lambda t, y: -y, bad, t, method="dopri5")
    with pytest.raises(AssertionError, match="underflow in dt"):
        # a derivative that is NaN makes every error ratio NaN -> dt becomes NaN -> `t0 + dt > t0` fails
        tfd().odeint(lambda t, y: y * float("nan"), y0, t, method="dopri5")


def test_tuple_state_with_per_component_tolerances():
    fn, ft = PROBLEMS["tuple_decay"](), PROBLEMS["tuple_decay"]()
    y0 = (np.linspace(1., 2., 7), np.linspace(0.5, 1.5, 33))
    t = np.linspace(0., 1., 4)
    kw = dict(method="dopri5", rtol=[1e-5, 1e-8], atol=[1e-7, 1e-10])
    st = np_ref.Stats()
    ref = np_ref.odeint(fn, y0, t, stats=st, **kw)
    got = tfd().odeint(ft, tuple(torch.tensor(v, device=DEV) for v in y0), torch.tensor(t), **kw)
    for r, g_ in zip(ref, got):
        assert max_rel_err(g_.cpu().numpy(), r) <= 1e-9
    s = tfd().last_stats
    assert (s["n_accepted"], s["n_rejected"]) == (st.n_acc, st.n_rej)


def test_output_dtype_and_shape_follow_y0():
    y0 = torch.ones(2, 3, 4, dtype=torch.float32, device=DEV)
    out = tfd().odeint(lambda t, y: -y, y0, torch.linspace(0., 1., 6, dtype=torch.float64), method="rk4")
    assert out.shape == (6, 2, 3, 4) and out.dtype == torch.float32
    out = tfd().odeint(lambda t, y: -y, y0.double(), torch.linspace(0., 1., 6), method="dopri5")
    assert out.shape == (6, 2, 3, 4) and out.dtype == torch.float64


# --------------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# --------------------------------------------------------------------------------------------------
def test_full_size_lorenz_short_horizon_vs_oracle():
    """BASELINE config 2 inputs (65 536 x 3 fp64, seeded) over a horizon the oracle finishes in seconds."""
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((65536, 3))
    t = np.arange(11) * 0.01
    ref, got, st, stats = _oracle_vs_engine("lorenz", y0, t, "float64", method="dopri5")
    assert max_rel_err(got, ref) <= 1e-6
    assert (stats["n_accepted"], stats["n_rejected"], stats["nfe"]) == (st.n_acc, st.n_rej, st.nfe)


def test_full_size_batch_duplication_invariance():
    """The error norm is a mean and the tolerance a max over the whole batch (misc.py:257-263): solving the
    batch stacked with a copy of itself must take the identical step sequence and give the same values."""
    rng = np.random.default_rng(1)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((32768, 3)), device=DEV)
    t = torch.arange(101, dtype=torch.float64) * 0.01
    f = PROBLEMS["lorenz"](backend="torch", device=DEV)
    a = tfd().odeint(f, y0, t, method="dopri5")
    sa = dict(tfd().last_stats)
    b = tfd().odeint(f, torch.cat([y0, y0]), t, method="dopri5")
    sb = dict(tfd().last_stats)
    assert (sa["n_accepted"], sa["n_rejected"]) == (sb["n_accepted"], sb["n_rejected"])
    assert torch.equal(b[:, :32768], b[:, 32768:])
    assert float((a - b[:, :32768]).abs().max()) < 1e-9


def test_full_size_linear_scaling_is_exact_for_power_of_two():
    """dim 128 x batch 65 536 fp64 (the headline kernel size).  For a linear right-hand side and atol = 0 the
    whole algorithm is homogeneous: scaling y0 by 4 scales errors and tolerances by exactly 4, so the step
    sequence is identical and every output is exactly 4x -- bit for bit."""
    torch.manual_seed(0)
    y0 = torch.randn(65536, 128, dtype=torch.float64, device=DEV)
    A = -0.5 * torch.eye(128, dtype=torch.float64, device=DEV) + 0.05 * torch.randn(128, 128, dtype=torch.float64,
                                                                                    device=DEV)
    f = lambda t, y: y @ A                                         # noqa: E731
    t = torch.tensor([0., 0.5, 1.0])
    a = tfd().odeint(f, y0, t, method="dopri5", rtol=1e-6, atol=0.0)
    sa = dict(tfd().last_stats)
    b = tfd().odeint(f, y0 * 4.0, t, method="dopri5", rtol=1e-6, atol=0.0)
    sb = dict(tfd().last_stats)
    assert (sa["n_accepted"], sa["n_rejected"]) == (sb["n_accepted"], sb["n_rejected"])
    assert torch.equal(a * 4.0, b)
    # and it is an actual solution: compare with the matrix exponential
    exact = y0 @ torch.linalg.matrix_exp(A * 1.0)
    assert float((a[-1] - exact).abs().max() / exact.abs().max()) < 1e-4


@pytest.mark.parametrize("method", ["euler", "midpoint", "heun", "rk4"])
def test_fixed_grid_func_reusing_its_output_buffer(method):
    y0 = torch.linspace(0.5, 1.5, 37, dtype=torch.float64, device=DEV)
    t = torch.linspace(0., 1., 9)
    buf = torch.empty_like(y0)

    def reuse(t, y):
        torch.mul(y, -0.7, out=buf)
        return buf
    a = tfd().odeint(lambda t, y: y * -0.7, y0, t, method=method)
    b = tfd().odeint(reuse, y0, t, method=method)
    assert torch.equal(a, b)


@pytest.mark.parametrize("method,kw", [("dopri5", {}), ("dopri8", dict(rtol=1e-9, atol=1e-9)), ("adaptive_heun", dict(rtol=1e-3, atol=1e-5))])
def test_cuda_graph_replay_is_bit_identical_to_eager(method, kw):
    """options={'cuda_graph': True} captures one attempt (func included) and replays it; every address and
    every kernel argument is static, so the results must be identical to the eager path bit for bit."""
    rng = np.random.default_rng(11)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((2048, 3)), device=DEV)
    t = torch.arange(41, dtype=torch.float64) * (0.01 if method != "adaptive_heun" else 0.001)
    f = PROBLEMS["lorenz"](backend="torch", device=DEV)
    a = tfd().odeint(f, y0, t, method=method, **kw)
    sa = dict(tfd().last_stats)
    b = tfd().odeint(f, y0, t, method=method, options=dict(cuda_graph=True), **kw)
    sb = dict(tfd().last_stats)
    assert sb["cuda_graph"] and not sa["cuda_graph"]
    assert (sa["n_accepted"], sa["n_rejected"], sa["nfe"]) == (sb["n_accepted"], sb["n_rejected"], sb["nfe"])
    assert torch.equal(a, b)


def test_cuda_graph_with_tuple_state_and_reverse_time():
    f = PROBLEMS["tuple_decay"]()
    y0 = (torch.linspace(1., 2., 7, dtype=torch.float64, device=DEV), torch.linspace(0.5, 1.5, 33, dtype=torch.float64, device=DEV))
    t = torch.linspace(1., 0., 5)
    a = tfd().odeint(f, y0, t, method="dopri5")
    b = tfd().odeint(f, y0, t, method="dopri5", options=dict(cuda_graph=True))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


# --------------------------------------------------------------------------------------------------
# option surface
# --------------------------------------------------------------------------------------------------
def test_option_surface_and_input_forms():
    f = lambda t, y: -y                                            # noqa: E731
    y0 = torch.linspace(0.5, 1.5, 10, dtype=torch.float64, device=DEV)
    t_cpu = torch.linspace(0., 1., 5)
    base = tfd().odeint(f, y0, t_cpu, method="dopri5")
    # t on the GPU, in float64, or as a python list: same result
    assert torch.equal(base, tfd().odeint(f, y0, t_cpu.to(DEV), method="dopri5"))
    assert torch.equal(base, tfd().odeint(f, y0, t_cpu.tolist(), method="dopri5"))
    # Dopri5's `tableau=` option (dopri5.py:53,67) and per-component tolerances as lists
    from tfdiffeq_b200 import tableaus
    assert torch.equal(base, tfd().odeint(f, y0, t_cpu, method="dopri5", options=dict(tableau=tableaus.DOPRI5)))
    assert torch.equal(base, tfd().odeint(f, y0, t_cpu, method="dopri5", rtol=[1e-7], atol=[1e-9]))
    # the caller's options dict is not modified
    opts = dict(first_step=0.05, cuda_graph=False)
    tfd().odeint(f, y0, t_cpu, method="dopri5", options=opts)
    assert opts == dict(first_step=0.05, cuda_graph=False)
    # grid_constructor (solvers.py:41-56): a finer uniform grid, outputs by linear interpolation
    gc = lambda func, y, t: torch.linspace(float(t[0]), float(t[-1]), 41, dtype=t.dtype)   # noqa: E731
    a = tfd().odeint(f, y0, t_cpu, method="rk4", options=dict(grid_constructor=gc))
    exact = y0.cpu().numpy()[None, :] * np.exp(-np.linspace(0., 1., 5))[:, None]
    assert np.max(np.abs(a.cpu().numpy() - exact)) < 1e-7
    # integer states are rejected (the engine integrates float32 / float64)
    with pytest.raises(TypeError):
        tfd().odeint(f, torch.ones(3, dtype=torch.int64, device=DEV), t_cpu, method="dopri5")
    with pytest.raises(AssertionError):
        tfd().odeint(f, y0, torch.tensor([0., 2., 1.]), method="dopri5")
    # func returning a tensor of the wrong size is an error, not silent corruption
    with pytest.raises((ValueError, RuntimeError)):
        tfd().odeint(lambda t, y: y[:3], y0, t_cpu, method="dopri5")


def test_nfe_counter_on_the_module_matches_the_reference_pattern():
    """The reference's models count function evaluations on the module (dense_odenet.py:38,78); on the eager
    path every evaluation is a real python call, so the count equals last_stats['nfe']."""
    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.nfe = 0

        def forward(self, t, y):
            self.nfe += 1
            return -y
    m = F()
    tfd().odeint(m, torch.ones(4, dtype=torch.float64, device=DEV), torch.tensor([0., 1.]), method="dopri5")
    assert m.nfe == tfd().last_stats["nfe"]
