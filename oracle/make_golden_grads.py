"""TEST INFRASTRUCTURE ONLY -- golden GRADIENTS from the UNMODIFIED reference adjoint.

Run in the build container (needs /root/reference):   python oracle/make_golden_grads.py

The reference's ``odeint_adjoint`` (tfdiffeq/adjoint.py:35-224) is loaded as it is and run on torch-CPU tensors over
``oracle/tf_shim.py`` plus the few extra ``tf.*`` symbols only the adjoint needs, defined here:

* ``tf.custom_gradient``   -> calls the decorated function, returns its value and keeps the ``grad`` closure so that this
                              script can call it with the loss's output gradients (what TF's tape would do);
* ``tf.GradientTape``      -> grad mode is enabled only inside the ``with`` block (``watch`` = ``requires_grad_``),
                              ``tape.gradient(..., output_gradients=, unconnected_gradients=ZERO)`` = ``torch.autograd.grad``
                              with ``None`` replaced by zeros;
* ``tf.keras.Model``       -> a minimal base class (``__call__`` -> ``call``, ``dtype``);
* ``tf.matmul``, ``tf.expand_dims``, ``tf.UnconnectedGradients``.

For every case the script stores the inputs, the loss weights ``w`` (the loss is ``sum_j <w_j, y(t_j)>``, so the output
gradient handed to the adjoint is exactly ``w``), the solution, and the reference's gradients w.r.t. ``y0``, ``t`` and
every parameter, plus -- as an independent cross-check -- the gradient obtained by back-propagating through the oracle's
discrete solver (``oracle/np_ref.py`` on torch-CPU tensors with autograd: discretise-then-optimise).  The fixtures
(``tests/golden/grad_*.npz``) travel to the GPU box; the reference does not.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_loader  # noqa: E402
import np_ref  # noqa: E402
from grad_cases import GRAD_CASES, build_params, rhs_torch  # noqa: E402


def extend_shim(tf):
    """The adjoint-only symbols (see the module docstring)."""
    import types

    def custom_gradient(f):
        def wrapper(*args):
            ans, grad = f(*args)
            wrapper.last_grad = grad
            return ans
        wrapper.last_grad = None
        return wrapper

    class GradientTape(object):
        def __enter__(self):
            self._prev = torch.is_grad_enabled()
            torch.set_grad_enabled(True)
            return self

        def __exit__(self, *exc):
            torch.set_grad_enabled(self._prev)
            return False

        def watch(self, x):
            for v in (x if isinstance(x, (list, tuple)) else (x,)):
                if not v.requires_grad:
                    v.requires_grad_(True)

        def gradient(self, target, sources, output_gradients=None, unconnected_gradients=None):
            sources = tuple(sources)
            if not target.requires_grad:
                return tuple(torch.zeros_like(s) for s in sources)
            gs = torch.autograd.grad(target, sources, grad_outputs=output_gradients, allow_unused=True)
            return tuple(torch.zeros_like(s) if g is None else g.detach() for g, s in zip(gs, sources))

    class Model(object):
        def __init__(self, dtype=None, **kw):
            self.dtype = dtype

        def __call__(self, *a, **kw):
            return self.call(*a, **kw)

    keras = types.ModuleType("tensorflow.keras")
    keras.Model = Model
    tf.keras = keras
    tf.custom_gradient = custom_gradient
    tf.GradientTape = GradientTape
    tf.matmul = torch.matmul
    tf.expand_dims = lambda x, axis=0: x.unsqueeze(axis)

    _make_variable = tf.Variable

    class Variable(object):
        """usable both as `tf.Variable(value)` (adams.py:34) and in `isinstance(x, (tf.Tensor, tf.Variable))`"""
        def __new__(cls, *a, **kw):
            return _make_variable(*a, **kw)
    tf.Variable = Variable

    class UnconnectedGradients(object):
        ZERO = "zero"
        NONE = "none"
    tf.UnconnectedGradients = UnconnectedGradients
    sys.modules["tensorflow.keras"] = keras


def load_reference_adjoint():
    pkg = ref_loader.load()
    import tensorflow as tf
    extend_shim(tf)
    path = os.path.join(ref_loader.REF_ROOT, "tfdiffeq", "adjoint.py")
    spec = importlib.util.spec_from_file_location("tfdiffeq.adjoint", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["tfdiffeq.adjoint"] = mod
    spec.loader.exec_module(mod)
    mod.move_to_device = lambda x, device: x
    return pkg, tf, mod


def reference_gradients(case, tf, adj):
    tdt = {"float32": torch.float32, "float64": torch.float64}[case["dtype"]]
    params = build_params(case, tdt)                                 # dict name -> leaf tensor (requires_grad)
    names = sorted(params)

    class Func(tf.keras.Model):
        def __init__(self):
            super(Func, self).__init__(dtype=tdt)

        def call(self, t, y):
            return rhs_torch(case, params, t, y)
    y0 = tuple(torch.tensor(np.asarray(v), dtype=tdt) for v in case["y0"])
    t = torch.tensor(np.asarray(case["t"], dtype=np.float64))
    w = tuple(torch.tensor(np.asarray(v), dtype=tdt) for v in case["w"])
    kw = dict(rtol=case["rtol"], atol=case["atol"], method=case["method"])
    if case.get("options"):
        kw["options"] = case["options"]
    single = len(y0) == 1
    with torch.no_grad():
        ys = adj.odeint_adjoint(Func(), y0[0] if single else y0, t, **kw)
        if not isinstance(ys, (tuple, list)):
            ys = (ys,)
        grad = adj.OdeintAdjointMethod.last_grad
        (*g_y0_t,), g_params = grad(*w, variables=[params[n] for n in names])
    g_y0, g_t = g_y0_t[:-1], g_y0_t[-1]
    return dict(sol=[v.detach().numpy() for v in ys], g_y0=[v.detach().numpy() for v in g_y0],
                g_t=g_t.detach().numpy().reshape(-1), g_params={n: g.detach().numpy() for n, g in zip(names, g_params)},
                params={n: params[n].detach().numpy() for n in names})


def oracle_backprop_gradients(case):
    """Discretise-then-optimise: autograd through the oracle's own discrete solver on torch-CPU tensors."""
    tdt = {"float32": torch.float32, "float64": torch.float64}[case["dtype"]]
    params = build_params(case, tdt)
    names = sorted(params)
    y0 = tuple(torch.tensor(np.asarray(v), dtype=tdt, requires_grad=True) for v in case["y0"])
    w = tuple(torch.tensor(np.asarray(v), dtype=tdt) for v in case["w"])

    def f(t, y):
        tt = torch.tensor(float(t), dtype=tdt)
        return rhs_torch(case, params, tt, y)
    sol = np_ref.odeint(f, y0[0] if len(y0) == 1 else y0, np.asarray(case["t"], dtype=np.float64), rtol=case["rtol"],
                        atol=case["atol"], method=case["method"], options=case.get("options"))
    if len(y0) == 1:
        sol = (sol,)
    loss = sum((s * w_).sum() for s, w_ in zip(sol, w))
    gs = torch.autograd.grad(loss, list(y0) + [params[n] for n in names], allow_unused=True)
    gs = [torch.zeros_like(x) if g is None else g for g, x in zip(gs, list(y0) + [params[n] for n in names])]
    return dict(g_y0=[g.numpy() for g in gs[:len(y0)]], g_params={n: g.numpy() for n, g in zip(names, gs[len(y0):])})


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    pkg, tf, adj = load_reference_adjoint()
    for name, case in GRAD_CASES.items():
        ref = reference_gradients(case, tf, adj)
        bp = oracle_backprop_gradients(case)
        blob = {"n_state": np.int64(len(case["y0"])), "t": np.asarray(case["t"], dtype=np.float64), "g_t": ref["g_t"]}
        for i in range(len(case["y0"])):
            if i < len(ref["sol"]):          # adjoint.py:221-222 hands back only the FIRST component of a tuple state
                blob["sol%d" % i] = ref["sol"][i]
            blob["g_y0_%d" % i] = ref["g_y0"][i]
            blob["bp_g_y0_%d" % i] = bp["g_y0"][i]
        for n in ref["g_params"]:
            blob["param_" + n] = ref["params"][n]
            blob["g_param_" + n] = ref["g_params"][n]
            blob["bp_g_param_" + n] = bp["g_params"][n]
        np.savez(os.path.join(out_dir, "grad_" + name + ".npz"), **blob)
        worst = 0.0
        for n in ref["g_params"]:
            a, b = ref["g_params"][n], bp["g_params"][n]
            worst = max(worst, float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))))
        for a, b in zip(ref["g_y0"], bp["g_y0"]):
            worst = max(worst, float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))))
        print("%-18s reference adjoint vs oracle back-prop: max rel diff %.2e   |g_t| %s" % (
            name, worst, np.array2string(ref["g_t"], precision=4)))


if __name__ == "__main__":
    main()
