"""TEST INFRASTRUCTURE ONLY -- load the UNMODIFIED reference hot-path modules over ``tf_shim``.

Works only where ``/root/reference`` exists (the build container).  Used to
(1) pin ``oracle/np_ref.py`` against the real reference source and
(2) generate ``tests/golden/*.npz`` (see ``oracle/make_golden.py``).

Recipe: a synthetic package object ``tfdiffeq`` whose
``__path__`` points at the reference, sub-modules loaded by path in dependency
order, ``misc.move_to_device`` patched to identity *before* the others bind it
(it parses TF device strings), ``tfdiffeq.compat`` stubbed.
``tfdiffeq/__init__.py`` itself is bypassed because it imports matplotlib.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("B2ODE_REFERENCE", "/root/reference")

_ORDER = ["misc", "rk_common", "interp", "solvers", "dopri5", "dopri8", "bosh3", "adaptive_huen",
          "tsit5", "fixed_grid", "fixed_adams", "adams", "odeint"]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "tfdiffeq"))


def load():
    """Return the loaded reference ``tfdiffeq`` package (cached)."""
    if "tfdiffeq" in sys.modules and getattr(sys.modules["tfdiffeq"], "_b2ode_shimmed", False):
        return sys.modules["tfdiffeq"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import tf_shim
    tf_shim.install()

    pkg_dir = os.path.join(REF_ROOT, "tfdiffeq")
    pkg = types.ModuleType("tfdiffeq")
    pkg.__path__ = [pkg_dir]
    pkg._b2ode_shimmed = True
    sys.modules["tfdiffeq"] = pkg

    compat = types.ModuleType("tfdiffeq.compat")

    def assign(tensor, val):
        import torch
        tensor.copy_(val if isinstance(val, torch.Tensor) else torch.as_tensor(val, dtype=tensor.dtype))
        return tensor
    compat.assign = assign
    sys.modules["tfdiffeq.compat"] = compat
    pkg.compat = compat

    for name in _ORDER:
        spec = importlib.util.spec_from_file_location("tfdiffeq." + name, os.path.join(pkg_dir, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["tfdiffeq." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        if name == "misc":
            mod.move_to_device = lambda x, device: x
    pkg.odeint = sys.modules["tfdiffeq.odeint"].odeint
    pkg.SOLVERS = sys.modules["tfdiffeq.odeint"].SOLVERS
    return pkg


class Counters(object):
    """Accepted / rejected / NFE counters obtained by wrapping, not editing, the reference."""

    def __init__(self):
        self.n_acc = 0
        self.n_rej = 0
        self.nfe = 0
        self.dt_trace = []


def run_reference(func, y0, t, counters=None, **kw):
    """Call the reference ``odeint`` (torch-CPU tensors in, torch-CPU tensors out).

    If ``counters`` is given, NFE is counted by wrapping ``func`` and accepts /
    rejects by wrapping each adaptive solver's step method and comparing ``t1``.
    """
    pkg = load()
    if counters is None:
        return pkg.odeint(func, y0, t, **kw)

    def counted(tt, yy):
        counters.nfe += 1
        return func(tt, yy)

    patched = []
    step_names = {"dopri5": "_adaptive_dopri5_step", "dopri8": "_adaptive_dopri8_step",
                  "bosh3": "_adaptive_bosh3_step", "adaptive_huen": "_adaptive_heun_step",
                  "tsit5": "_adaptive_tsit5_step"}
    cls_names = {"dopri5": "Dopri5Solver", "dopri8": "Dopri8Solver", "bosh3": "Bosh3Solver",
                 "adaptive_huen": "AdaptiveHeunSolver", "tsit5": "Tsit5Solver"}
    for modname, meth in step_names.items():
        cls = getattr(sys.modules["tfdiffeq." + modname], cls_names[modname])
        orig = getattr(cls, meth)

        def wrapped(self, rk_state, _orig=orig):
            new = _orig(self, rk_state)
            counters.dt_trace.append(float(rk_state.dt))
            if float(new.t1) > float(rk_state.t1):
                counters.n_acc += 1
            else:
                counters.n_rej += 1
            return new
        setattr(cls, meth, wrapped)
        patched.append((cls, meth, orig))
    # variable-coefficient Adams: a rejected step hands back the same y_n object (adams.py:172), an accepted one the predictor
    vcls = sys.modules["tfdiffeq.adams"].VariableCoefficientAdamsBashforth
    vorig = vcls._adaptive_adams_step

    def vwrapped(self, st, final_t, _orig=vorig):
        dt = float(min(float(st.next_t), float(final_t)) - float(st.prev_t[0]))
        new = _orig(self, st, final_t)
        counters.dt_trace.append(dt)
        if new.y_n is not st.y_n:
            counters.n_acc += 1
        else:
            counters.n_rej += 1
        return new
    vcls._adaptive_adams_step = vwrapped
    patched.append((vcls, "_adaptive_adams_step", vorig))
    try:
        return pkg.odeint(counted, y0, t, **kw)
    finally:
        for cls, meth, orig in patched:
            setattr(cls, meth, orig)
