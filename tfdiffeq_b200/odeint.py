"""Public entry point with the reference's signature and dispatch (tfdiffeq/odeint.py:11-81)."""
from .misc import _check_inputs
from .multistep import AdamsBashforth, AdamsBashforthMoulton, VariableCoefficientAdamsBashforth
from .solvers import (AdaptiveHeunSolver, Bosh3Solver, Dopri5Solver, Dopri8Solver, Euler, Heun, Midpoint, RK4,
                      Tsit5Solver)

# method name -> solver class; the names (including the historical 'huen' spelling) are the reference's
# (tfdiffeq/odeint.py:11-25)
_ADAPTIVE_RK = dict(dopri5=Dopri5Solver, dopri8=Dopri8Solver, bosh3=Bosh3Solver, tsit5=Tsit5Solver,
                    adaptive_heun=AdaptiveHeunSolver)
_FIXED_GRID = dict(euler=Euler, midpoint=Midpoint, rk4=RK4, heun=Heun, huen=Heun)
_MULTISTEP = dict(adams=VariableCoefficientAdamsBashforth, fixed_adams=AdamsBashforthMoulton, explicit_adams=AdamsBashforth)
SOLVERS = dict(_ADAPTIVE_RK, **_FIXED_GRID, **_MULTISTEP)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    """Integrate ``dy/dt = func(t, y), y(t[0]) = y0`` and return ``y`` at every time in ``t``.

    Same contract as the reference (tfdiffeq/odeint.py:28-81): ``y0`` is a tensor of any shape or a tuple of
    tensors, ``t`` a strictly monotone 1-D tensor (decreasing ``t`` integrates backwards, misc.py:318-321),
    the result has shape ``(len(t), *y0.shape)`` (a tuple of such for tuple states) in ``y0``'s dtype.
    ``func(t, y)`` is any callable on torch CUDA tensors, typically an ``nn.Module``; ``t`` reaches it as a
    0-dim device tensor.  Raises ``ValueError`` if ``options`` is given without ``method``, ``KeyError`` for
    an unknown ``method``, ``TypeError`` for non-numeric inputs, ``AssertionError`` for non-monotone ``t``,
    step-size underflow, non-finite states or ``max_num_steps``; unknown option keys only warn.
    """
    tensor_input, func, y0, t = _check_inputs(func, y0, t)
    if options is not None and method is None:
        raise ValueError('cannot supply `options` without specifying `method`')      # odeint.py:72-73
    solver_cls = SOLVERS['dopri5' if method is None else method]                     # unknown name: KeyError (:77)
    odeint.last_solver = solver = solver_cls(func, y0, rtol=rtol, atol=atol, **(options or {}))
    solution = solver.integrate(t)
    return solution[0] if tensor_input else solution


odeint.last_solver = None
