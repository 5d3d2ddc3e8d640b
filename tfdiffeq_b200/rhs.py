"""Built-in right-hand sides.

Each class is an ordinary ``nn.Module`` -- ``forward(t, y)`` is written in plain torch ops and works anywhere
(on the generic path, on CPU, under autograd).  When an instance is handed to ``odeint`` with an adaptive
Runge-Kutta method and a single ``(..., dim)`` CUDA state, the solver recognises it and runs the WHOLE solve in
one persistent kernel (``b2ode_fused_solve``): every trajectory lives in one thread's registers, HBM traffic is
the solution slab only.  The kernel evaluates exactly the same IEEE operations in the same order as
``forward`` does, so both paths agree to the last bit per stage; ``options={'fused_rhs': False}`` forces the
generic path.
"""
import torch
import torch.nn as nn

from . import _lib


class BuiltinRHS(nn.Module):
    kind = None      # B2ODE_RHS_* code
    dim = None       # size of the last state axis

    def rhs_params(self):
        raise NotImplementedError

    def rhs_data(self, dtype, device):
        """Device buffer of staged weights for the kernel (None for parameter-free systems)."""
        return None


class Lorenz(BuiltinRHS):
    """examples/lorenz_attractor.py:20-37, vectorised over leading batch axes: state (..., 3)."""
    kind, dim = _lib.RHS_LORENZ, 3

    def __init__(self, sigma=10.0, beta=8.0 / 3.0, rho=28.0):
        super(Lorenz, self).__init__()
        self.sigma, self.beta, self.rho = float(sigma), float(beta), float(rho)

    def rhs_params(self):
        return [self.sigma, self.beta, self.rho]

    def forward(self, t, y):
        x, yy, z = y[..., 0], y[..., 1], y[..., 2]
        return torch.stack([self.sigma * (yy - x), x * (self.rho - z) - yy, x * yy - self.beta * z], -1)


class LotkaVolterra(BuiltinRHS):
    """README.md:67-81: x' = a x - b x z, z' = -c z + d x z; state (..., 2)."""
    kind, dim = _lib.RHS_LOTKA_VOLTERRA, 2

    def __init__(self, a=1.5, b=1.0, c=3.0, d=1.0):
        super(LotkaVolterra, self).__init__()
        self.a, self.b, self.c, self.d = float(a), float(b), float(c), float(d)

    def rhs_params(self):
        return [self.a, self.b, self.c, self.d]

    def forward(self, t, y):
        x, z = y[..., 0], y[..., 1]
        return torch.stack([self.a * x - self.b * x * z, -self.c * z + self.d * x * z], -1)


class CubicMLP(BuiltinRHS):
    """examples/ode_demo.py:115-129 (BASELINE config 3): ``W2 . tanh(W1 . y**3 + b1) + b2`` with a 2 -> H -> 2
    network (H <= 128); ``cube=False`` drops the ``y**3``.  Weights are ordinary ``nn.Parameter`` s, so the module
    trains like any other (gradients come from the generic / adjoint path; the fused kernels are forward only).
    """
    kind, dim = _lib.RHS_CUBIC_MLP, 2

    def __init__(self, hidden=50, cube=True, std=0.1, dtype=torch.float32, generator=None):
        super(CubicMLP, self).__init__()
        if not 1 <= hidden <= 128:
            raise ValueError("hidden width must be in [1, 128]")
        self.hidden, self.cube = int(hidden), bool(cube)
        self.W1 = nn.Parameter(torch.randn(2, hidden, dtype=dtype, generator=generator) * std)
        self.b1 = nn.Parameter(torch.zeros(hidden, dtype=dtype))
        self.W2 = nn.Parameter(torch.randn(hidden, 2, dtype=dtype, generator=generator) * std)
        self.b2 = nn.Parameter(torch.zeros(2, dtype=dtype))

    def rhs_params(self):
        return [float(self.hidden), 1.0 if self.cube else 0.0]

    def rhs_data(self, dtype, device):
        with torch.no_grad():
            return torch.cat([self.W1.reshape(-1), self.b1.reshape(-1), self.W2.reshape(-1), self.b2.reshape(-1)]).to(
                device=device, dtype=dtype).contiguous()

    def forward(self, t, y):
        u = y ** 3 if self.cube else y
        return torch.tanh(u @ self.W1 + self.b1) @ self.W2 + self.b2
