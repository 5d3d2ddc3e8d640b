"""CPU: the reference's DETEST benchmark (tests/DETEST/run.py) against the oracle -- closed forms for A1-A4, and the
benchmark's own score (error against dopri5 at a tight tolerance over [0, 20]) falling with the tolerance, for the
adaptive Runge-Kutta path and the variable-coefficient Adams solver."""
import math

import numpy as np
import pytest

import np_ref
from detest_problems import NAMES, make

T = np.array([0.0, 20.0])


def _solve(name, method, tol):
    f, y0, _ = make(name, np)
    st = np_ref.Stats()
    out = np_ref.odeint(f, y0, T, rtol=tol, atol=tol, method=method, stats=st)
    return np.asarray(out[-1]), st


@pytest.mark.parametrize("name", ["A1", "A2", "A3", "A4"])
@pytest.mark.parametrize("method", ["dopri5", "dopri8", "adams"])
def test_closed_form_solutions(name, method):
    _, _, exact = make(name, np)
    got, _ = _solve(name, method, 1e-9)
    # adams carries its predictor forward (adams.py:211): its global error sits far above the tolerance
    assert abs(float(got) - exact(20.0)) <= (2e-4 if method == "adams" else 2e-5)


@pytest.mark.parametrize("name", NAMES)
def test_dopri5_score_falls_with_the_tolerance(name):
    ref, _ = _solve(name, "dopri5", 1e-11)
    errs, nfes = [], []
    for tol in (1e-3, 1e-6, 1e-9):
        got, st = _solve(name, "dopri5", tol)
        errs.append(float(np.sqrt(np.mean((got - ref) ** 2))))       # run.py:52
        nfes.append(st.nfe)
    scale = max(1.0, float(np.max(np.abs(ref))))
    assert errs[2] <= 1e-5 * scale and errs[1] <= 1e-2 * scale
    assert errs[2] <= errs[0] or errs[0] < 1e-9
    assert nfes[0] <= nfes[1] <= nfes[2]


@pytest.mark.parametrize("name", ["A3", "B1", "B5", "C1", "C3", "D1", "E1", "E2", "E4"])
def test_adams_score_at_1e6(name):
    ref, _ = _solve(name, "dopri5", 1e-11)
    got, st = _solve(name, "adams", 1e-6)
    assert float(np.sqrt(np.mean((got - ref) ** 2))) <= 5e-2 * max(1.0, float(np.max(np.abs(ref))))
    assert st.nfe == 2 + 2 * st.n_acc + st.n_rej                     # 2 at start-up, 2 per accepted step, 1 per reject
