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


_ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2, "softplus": 3}


_TF32_CACHE = {}


def _tf32_weight(weight):
    """The B operand rounded to TF32 (round-to-nearest, ties away: what cvt.rna.tf32.f32 does), cached per weight
    version so the rounding runs once per optimiser step, not once per func evaluation."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape))
    hit = _TF32_CACHE.get(id(weight))
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        i = weight.detach().contiguous().view(torch.int32)
        r = ((i + 0x1000) & -0x2000).view(torch.float32)
    _TF32_CACHE[id(weight)] = (key, r)
    return r


def dense_layer(x, weight, bias, act="none", stage=None):
    """``act(x @ weight.T + bias)`` on the tcgen05 tensor cores (fp32 storage, TF32 math): ``b2ode_dense_layer``.

    ``stage = (k_tensors, coefs, state_ptr, ystage_or_None)`` makes the Runge-Kutta stage combine the A-operand
    producer: A = x + sum_j (dt * coefs[j]) * k_tensors[j], dt read from the device state."""
    import ctypes as C
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    karr = carr = state = ys = None
    nk = 0
    if stage is not None:
        ks, coefs, state, ys = stage
        nk = len(ks)
        karr = (C.c_void_p * nk)(*[k.data_ptr() for k in ks])
        carr = (C.c_double * nk)(*coefs)
    _lib.check(_lib.lib.b2ode_dense_layer(
        C.c_void_p(x.data_ptr()), karr, carr, nk, C.c_void_p(state) if state else None,
        C.c_void_p(ys.data_ptr()) if ys is not None else None, C.c_void_p(_tf32_weight(weight).data_ptr()),
        C.c_void_p(bias.data_ptr()) if bias is not None else None, C.c_void_p(out.data_ptr()), M, K, N, _ACT[act],
        C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out


_PACK_CACHE = {}


def _mlp3_packed(fc1, fc2, fc3):
    """The three weights as ``b2ode_mlp3``'s shared-memory image, rebuilt only when a weight changes."""
    import ctypes as C
    ws = (fc1.weight, fc2.weight, fc3.weight)
    key = tuple((w.data_ptr(), w._version, tuple(w.shape)) for w in ws)
    hit = _PACK_CACHE.get(id(fc1))
    if hit is not None and hit[0] == key:
        return hit[1]
    H, D = fc1.weight.shape
    nbytes = _lib.lib.b2ode_mlp3_packed_bytes(D, H)
    if nbytes < 0:
        raise ValueError("mlp3 needs dim and hidden to be multiples of 16 in [16, 256]")
    packed = torch.empty(nbytes, dtype=torch.uint8, device=fc1.weight.device)
    cw = [w.detach().contiguous() for w in ws]
    _lib.check(_lib.lib.b2ode_mlp3_pack(
        C.c_void_p(cw[0].data_ptr()), C.c_void_p(cw[1].data_ptr()), C.c_void_p(cw[2].data_ptr()), D, H,
        C.c_void_p(packed.data_ptr()), C.c_void_p(torch.cuda.current_stream(packed.device).cuda_stream)))
    _PACK_CACHE[id(fc1)] = (key, packed)
    return packed


def mlp3(x, fc1, fc2, fc3, act="relu", stage=None):
    """``fc3(act(fc2(act(fc1(x)))))`` in one launch (``b2ode_mlp3``): hidden activations never reach HBM.
    ``stage`` as in :func:`dense_layer`."""
    import ctypes as C
    M, D = x.shape
    H = fc1.weight.shape[0]
    out = torch.empty((M, D), dtype=torch.float32, device=x.device)
    karr = carr = state = ys = None
    nk = 0
    if stage is not None:
        ks, coefs, state, ys = stage
        nk = len(ks)
        karr = (C.c_void_p * nk)(*[k.data_ptr() for k in ks])
        carr = (C.c_double * nk)(*coefs)

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(_lib.lib.b2ode_mlp3(
        ptr(x), karr, carr, nk, C.c_void_p(state) if state else None, ptr(ys),
        ptr(_mlp3_packed(fc1, fc2, fc3)), ptr(fc1.bias), ptr(fc2.bias), ptr(fc3.bias), ptr(out), M, D, H, _ACT[act],
        C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out


class DenseMLP(nn.Module):
    """The reference's ``ODEFunc`` (tfdiffeq/models/dense_odenet.py:11-92, time-independent form): fc1 -> act ->
    fc2 -> act -> fc3 on a ``(batch, dim)`` state, counting ``nfe`` like the reference does (:78).

    Under ``torch.no_grad()`` on a CUDA fp32 state -- which is how ``odeint`` evaluates ``func`` -- the three
    layers run on the tcgen05 tensor cores (TF32 math, TensorFlow's default for fp32 matmuls on Ampere+), and the
    adaptive solvers feed the first layer straight from the stage combine (the stage input never round-trips
    HBM for the GEMM).  With autograd enabled (training, ``odeint_adjoint``'s VJPs) it is plain torch.
    ``tensor_cores=False`` forces plain torch everywhere."""

    def __init__(self, dim, hidden, non_linearity="relu", tensor_cores=True, dtype=torch.float32, chain=True):
        super(DenseMLP, self).__init__()
        if non_linearity not in ("relu", "tanh", "softplus"):
            raise ValueError("non_linearity must be relu, tanh or softplus")
        self.dim, self.hidden, self.non_linearity, self.tensor_cores = int(dim), int(hidden), non_linearity, tensor_cores
        self.fc1 = nn.Linear(dim, hidden, dtype=dtype)
        self.fc2 = nn.Linear(hidden, hidden, dtype=dtype)
        self.fc3 = nn.Linear(hidden, dim, dtype=dtype)
        self.nfe = 0
        self.chain = chain

    def uses_tensor_cores(self, x):
        return (self.tensor_cores and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()
                and self.fc1.weight.dtype == torch.float32 and self.dim % 16 == 0 and self.hidden % 16 == 0
                and x.shape[-1] == self.dim)

    def chained(self):
        """One-launch form (``b2ode_mlp3``) when both widths fit the 128 KB activation tile."""
        return self.chain and self.dim <= 256 and self.hidden <= 256

    def _tail(self, h1):
        h2 = dense_layer(h1, self.fc2.weight, self.fc2.bias, self.non_linearity)
        return dense_layer(h2, self.fc3.weight, self.fc3.bias, "none")

    def forward_from_stage(self, y0, ks, coefs, state_ptr, ystage):
        """First layer fed by the stage combine of y0 and the k's (all ``(batch, dim)`` fp32 CUDA tensors)."""
        self.nfe += 1
        if self.chained():
            return mlp3(y0, self.fc1, self.fc2, self.fc3, self.non_linearity, stage=(ks, coefs, state_ptr, ystage))
        h1 = dense_layer(y0, self.fc1.weight, self.fc1.bias, self.non_linearity, stage=(ks, coefs, state_ptr, ystage))
        return self._tail(h1)

    def forward(self, t, x):
        self.nfe += 1
        if self.uses_tensor_cores(x):
            x2 = x.reshape(-1, self.dim)
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            if self.chained():
                return mlp3(x2, self.fc1, self.fc2, self.fc3, self.non_linearity).reshape(x.shape)
            h1 = dense_layer(x2, self.fc1.weight, self.fc1.bias, self.non_linearity)
            return self._tail(h1).reshape(x.shape)
        act = {"relu": torch.relu, "tanh": torch.tanh, "softplus": torch.nn.functional.softplus}[self.non_linearity]
        return self.fc3(act(self.fc2(act(self.fc1(x)))))
