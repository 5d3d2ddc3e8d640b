"""tfdiffeq_b200 -- the Runge-Kutta hot path of tfdiffeq's ``odeint`` on B200 (sm_100a) kernels.

Drop-in for ``tfdiffeq.odeint`` / ``tfdiffeq.odeint_adjoint`` (tfdiffeq/__init__.py:2-3) with torch CUDA
tensors in place of TF tensors and ``func`` a PyTorch callable.  Importing this package loads
``libb2ode.so``; there is no CPU fallback.
"""
from .odeint import SOLVERS, odeint            # noqa: F401
from .adjoint import odeint_adjoint            # noqa: F401
from .solvers import last_stats                # noqa: F401
from . import rhs                               # noqa: F401

__all__ = ['odeint', 'odeint_adjoint', 'SOLVERS', 'last_stats', 'rhs']
__version__ = '0.1.0'
