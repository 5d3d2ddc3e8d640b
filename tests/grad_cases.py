"""Gradient parity cases shared by ``oracle/make_golden_grads.py`` (the unmodified reference adjoint on torch-CPU) and
``tests/test_adjoint_gpu.py`` (``tfdiffeq_b200.odeint_adjoint`` on the GPU).

Every case is ``dict(kind, dtype, y0 (tuple of arrays), t, w (loss weights, one array of shape (T, *y0_i.shape) per
component: loss = sum_i sum_j <w_i[j], y_i(t_j)>), rtol, atol, method, options, seed)``.  The right-hand sides restate the
reference's own gradient fixtures (tests/gradient_tests.py:106-123: ``y**3 @ A``) plus an ODENet-style MLP
(tfdiffeq/models/dense_odenet.py:85-92), a time-dependent field and a two-component tuple state.  Tuple components have
equal shapes because the reference's adjoint stacks ``func``'s outputs (tfdiffeq/adjoint.py:80).
"""
import numpy as np
import torch


def _rng(seed):
    return np.random.default_rng(seed)


def _w(seed, T, shapes):
    g = _rng(seed + 1000)
    return tuple(g.standard_normal((T,) + tuple(s)) for s in shapes)


def _case(kind, dtype, y0, t, rtol, atol, method, seed, options=None):
    y0 = tuple(np.asarray(v, dtype=np.float64) for v in y0)
    t = np.asarray(t, dtype=np.float64)
    return dict(kind=kind, dtype=dtype, y0=y0, t=t, w=_w(seed, len(t), [v.shape for v in y0]), rtol=rtol, atol=atol,
                method=method, options=options, seed=seed)


GRAD_CASES = {
    # tests/gradient_tests.py:106-123
    "spiral3_dopri5": _case("spiral3", "float64", [[[2.0, 0.0], [1.0, 0.5]]], np.linspace(0., 1.5, 4), 1e-8, 1e-10, "dopri5", 1),
    "spiral3_dopri8": _case("spiral3", "float64", [[[2.0, 0.0], [1.0, 0.5]]], np.linspace(0., 1.5, 4), 1e-8, 1e-10, "dopri8", 2),
    "spiral3_rk4": _case("spiral3", "float64", [[[2.0, 0.0], [1.0, 0.5]]], np.linspace(0., 1.5, 31), 1e-8, 1e-10, "rk4", 3),
    "mlp_tanh_dopri5": _case("mlp", "float64", [_rng(4).standard_normal((5, 4))], [0.0, 0.4, 1.0], 1e-7, 1e-9, "dopri5", 4),
    "mlp_tanh_f32": _case("mlp", "float32", [_rng(5).standard_normal((5, 4))], [0.0, 0.4, 1.0], 1e-4, 1e-5, "dopri5", 5),
    "timedep_dopri5": _case("timedep", "float64", [[0.5, -1.0, 2.0]], [0.0, 0.4, 1.1, 1.7], 1e-8, 1e-10, "dopri5", 6),
    "tuple2_dopri5": _case("tuple2", "float64", [[1.0, 2.0, -0.5], [0.5, 0.1, -0.2]], np.linspace(0., 1., 3), 1e-8, 1e-10, "dopri5", 7),
}


def build_params(case, tdt, device=None):
    """name -> leaf tensor with requires_grad (deterministic from the case's seed)."""
    g = _rng(case["seed"] + 2000)
    k = case["kind"]
    if k == "spiral3":
        vals = {"A": np.array([[-0.1, 2.0], [-2.0, -0.1]])}
    elif k == "mlp":
        vals = {"W1": 0.5 * g.standard_normal((4, 8)), "b1": 0.1 * g.standard_normal(8),
                "W2": 0.5 * g.standard_normal((8, 4)), "b2": 0.1 * g.standard_normal(4)}
    elif k == "timedep":
        vals = {"w": np.array([0.7, -0.3, 0.2])}
    elif k == "tuple2":
        vals = {"a": np.array(0.3), "c": np.array([0.2, -0.1, 0.4])}
    else:
        raise KeyError(k)
    return {n: torch.tensor(v, dtype=tdt, device=device, requires_grad=True) for n, v in vals.items()}


def rhs_torch(case, p, t, y):
    """func(t, y) in torch ops; ``y`` is a tensor for single-component cases' inner call or a tuple (the reference's
    TupleFunc hands a 1-tuple's element through)."""
    k = case["kind"]
    if k == "spiral3":
        return (y ** 3) @ p["A"]
    if k == "mlp":
        return torch.tanh(y @ p["W1"] + p["b1"]) @ p["W2"] + p["b2"]
    if k == "timedep":
        return torch.tanh(y * p["w"]) * torch.cos(t).to(y.dtype) - 0.1 * y
    if k == "tuple2":
        u, v = y
        return (-p["a"] * u + v * p["c"], -2.0 * p["a"] * v + 0.1 * torch.sin(u))
    raise KeyError(k)
