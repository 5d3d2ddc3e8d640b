"""The DETEST problem set (Hull, Enright, Fellen & Sedgwick 1972) as the reference's own benchmark uses it
(tests/DETEST/detest.py: classes A-E; tests/DETEST/run.py integrates each over [0, 20] with dopri5 and adams at
tol = 1e-3 / 1e-6 / 1e-9 and scores it against dopri5 at 1e-12).  Restated backend-agnostically: `make(name, xp)`
returns (f, y0, exact-or-None) where xp is numpy or torch and `f(t, y)` works on that backend's arrays.
C5 (the five-body problem) is left out: the reference's initial data for it contains a typo (detest.py:247,
`165699966404` for 1.65699966404), so it does not define a meaningful trajectory."""
import math

import numpy as np


def _stack(xp, parts):
    return xp.stack(parts)


def _band(n, lower, diag, upper):
    a = np.zeros((n, n))
    for i in range(n):
        a[i, i] = diag[i] if hasattr(diag, "__len__") else diag
        if i + 1 < n:
            a[i + 1, i] = lower[i] if hasattr(lower, "__len__") else lower
            a[i, i + 1] = upper[i] if hasattr(upper, "__len__") else upper
    return a


def _as(xp, a, like=None):
    if xp is np:
        return np.asarray(a, dtype=np.float64)
    import torch
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=like)


def make(name, xp=np, device=None):
    cls, idx = name[0], int(name[1])
    sqrt, sin, cos, exp = xp.sqrt, xp.sin, xp.cos, xp.exp
    arr = lambda v: _as(xp, v, device)                                       # noqa: E731
    if cls == "A":                                                            # detest.py:9-40, scalar equations
        f = [lambda t, y: -y, lambda t, y: -y ** 3 / 2, lambda t, y: y * cos(t), lambda t, y: y / 4 * (1 - y / 20),
             lambda t, y: (y - t) / (y + t)][idx - 1]
        y0 = arr(4.0 if idx == 5 else 1.0)
        exact = [lambda t: math.exp(-t), lambda t: 1 / math.sqrt(t + 1), lambda t: math.exp(math.sin(t)),
                 lambda t: 20 / (1 + 19 * math.exp(-t / 4)), None][idx - 1]
        return f, y0, exact
    if cls == "B":                                                            # detest.py:46-113, small systems
        if idx == 1:
            return (lambda t, y: _stack(xp, [2 * (y[0] - y[0] * y[1]), -(y[1] - y[0] * y[1])])), arr([1., 3.]), None
        if idx == 2:
            A = arr([[-1., 1., 0.], [1., -2., 1.], [0., 1., -1.]])
            return (lambda t, y: A @ y), arr([2., 0., 1.]), None
        if idx == 3:
            return (lambda t, y: _stack(xp, [-y[0], y[0] - y[1] * y[1], y[1] * y[1]])), arr([1., 0., 0.]), None
        if idx == 4:
            def f(t, y):
                a = sqrt(y[0] * y[0] + y[1] * y[1])
                return _stack(xp, [-y[1] - y[0] * y[2] / a, y[0] - y[1] * y[2] / a, y[0] / a])
            return f, arr([3., 0., 0.]), None
        return (lambda t, y: _stack(xp, [y[1] * y[2], -y[0] * y[2], -0.51 * y[0] * y[1]])), arr([0., 1., 1.]), None
    if cls == "C":                                                            # detest.py:119-202, banded linear systems
        if idx == 1:
            A = _band(10, 1.0, [-1.0] * 9 + [0.0], 0.0)
        elif idx == 2:
            k = np.linspace(1., 9., 9)
            A = _band(10, k, list(-k) + [0.0], 0.0)
        else:
            n = 10 if idx == 3 else 51
            A = _band(n, 1.0, -2.0, 1.0)
        A = arr(A)
        y0 = np.zeros(A.shape[0])
        y0[0] = 1.0
        return (lambda t, y: A @ y), arr(y0), None
    if cls == "D":                                                            # detest.py:263-286, Kepler orbits
        eps = [0.1, 0.3, 0.5, 0.7, 0.9][idx - 1]

        def f(t, y):
            r = (y[0] ** 2 + y[1] ** 2) ** (3 / 2)
            return _stack(xp, [y[2], y[3], -y[0] / r, -y[1] / r])
        return f, arr([1 - eps, 0., 0., math.sqrt((1 + eps) / (1 - eps))]), None
    if cls == "E":                                                            # detest.py:289-351, second-order equations
        if idx == 1:
            return (lambda t, y: _stack(xp, [y[1], -(y[1] / (t + 1) + (1 - 0.25 / (t + 1) ** 2) * y[0])])), \
                arr([.671396707141803, .0954005144474744]), None
        if idx == 2:
            return (lambda t, y: _stack(xp, [y[1], (1 - y[0] ** 2) * y[1] - y[0]])), arr([2., 0.]), None
        if idx == 3:
            return (lambda t, y: _stack(xp, [y[1], y[0] ** 3 / 6 - y[0] + 2 * sin(2.78535 * t)])), arr([0., 0.]), None
        if idx == 4:
            return (lambda t, y: _stack(xp, [y[1], .32 - .4 * y[1] ** 2])), arr([30., 0.]), None
        return (lambda t, y: _stack(xp, [y[1], sqrt(1 + y[1] ** 2) / (25 - t)])), arr([0., 0.]), None
    raise KeyError(name)


NAMES = [c + str(i) for c in "ABCDE" for i in range(1, 6) if c + str(i) != "C5"]
