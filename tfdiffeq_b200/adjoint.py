"""``odeint_adjoint``: O(1)-memory gradients by solving the augmented system backwards in time.

Caller of the hot path: the reference implements it as a ``tf.custom_gradient`` around
``odeint`` (tfdiffeq/adjoint.py:35-180); here it is a ``torch.autograd.Function`` with the same algorithm.
The backward pass runs the same sm_100a kernels on the augmented tuple state
``(y, adj_y, adj_t, adj_params)`` -- components of unequal shapes, per-component error norms -- and takes
the vector-Jacobian products of ``func`` from PyTorch autograd (the reference uses a ``GradientTape``,
adjoint.py:77-96).
"""
import torch
import torch.nn as nn

from . import rhs as _rhs
from .odeint import odeint


def _flatten(seq):
    flat = [p.reshape(-1) for p in seq]
    return torch.cat(flat) if len(flat) > 0 else torch.tensor([])


# statistics of the most recent odeint_adjoint call: the forward solve and every backward (augmented) solve
last_stats = {"forward": None, "backward": []}


def _group_of(options):
    return (options or {}).get("shared_step_group") if isinstance(options, dict) else None


def _all_reduce_sum(x, group):
    """Sum over the ranks of a shared-step group (NCCL on the current stream).  Batch-summed quantities of the
    adjoint -- a^T df/dtheta, a^T df/dt, dL/dt_i -- are sums over ALL trajectories of the system, i.e. over all shards."""
    import torch.distributed as dist
    if x.numel():
        dist.all_reduce(x, group=group.group)
    return x


class _OdeintAdjoint(torch.autograd.Function):
    """tfdiffeq/adjoint.py:35-180"""

    @staticmethod
    def forward(ctx, func, n_tensors, options, t, flat_params, *y0):
        ctx.func, ctx.options, ctx.n_tensors = func, options, n_tensors
        _rhs._FORCE_ACCURATE[0] += 1
        try:
            with torch.no_grad():
                ans = odeint(func, tuple(y0), t, rtol=options["rtol"], atol=options["atol"], method=options["method"],
                             options=options["options"])                                      # adjoint.py:54
        finally:
            _rhs._FORCE_ACCURATE[0] -= 1
        from . import solvers as _solvers
        last_stats["forward"] = dict(_solvers.last_stats)
        last_stats["backward"] = []
        ctx.save_for_backward(t, flat_params, *ans)
        return ans

    @staticmethod
    def backward(ctx, *grad_output):
        t, flat_params, *ans = ctx.saved_tensors
        func, opts, n_tensors = ctx.func, ctx.options, ctx.n_tensors
        f_params = tuple(p for p in func.parameters() if p.requires_grad)
        dev, dtype = ans[0].device, ans[0].dtype
        grad_output = tuple(g if g is not None else torch.zeros_like(a) for g, a in zip(grad_output, ans))
        # Shards of one system on several GPUs (options={'shared_step_group': g}): y and adj_y are sharded like the
        # batch; adj_t and adj_params are sums over the whole batch, so their derivatives are all-reduced on every
        # evaluation and the two components are *replicated* (bit-identical on every rank, counted once in the norm)
        adj_options = opts["adjoint_options"]
        group = _group_of(adj_options)
        if group is not None and group.world > 1:
            adj_options = dict(adj_options, replicated_components=(2 * n_tensors, 2 * n_tensors + 1))
        else:
            group = None

        def augmented_dynamics(tt, y_aug):
            # adjoint.py:71-107: (f, -a^T df/dy, -a^T df/dt, -a^T df/dtheta)
            y, adj_y = y_aug[:n_tensors], y_aug[n_tensors:2 * n_tensors]
            with torch.enable_grad():
                tt_ = tt.detach().requires_grad_(True)
                y_ = tuple(v.detach().requires_grad_(True) for v in y)
                func_eval = func(tt_, y_)
                # outputs that depend on none of (t, y, theta) -- a constant field, a component func passes through
                # detached -- have no graph: their VJP is zero (the reference asks for UnconnectedGradients.ZERO,
                # adjoint.py:88-96) and autograd.grad must not see them
                live = [(f, -a) for f, a in zip(func_eval, adj_y) if f.requires_grad]
                wrt = (tt_,) + y_ + f_params
                if live:
                    vjps = torch.autograd.grad([f for f, _ in live], wrt, [a for _, a in live], allow_unused=True)
                else:
                    vjps = (None,) * len(wrt)
            vjp_t, vjp_y, vjp_params = vjps[0], vjps[1:1 + n_tensors], vjps[1 + n_tensors:]
            vjp_t = torch.zeros_like(tt_) if vjp_t is None else vjp_t
            vjp_y = tuple(torch.zeros_like(v) if g is None else g for g, v in zip(vjp_y, y_))
            if len(f_params) == 0:
                vjp_p = torch.zeros((), dtype=dtype, device=dev)                              # adjoint.py:103-105
            else:
                vjp_p = _flatten([torch.zeros_like(p) if g is None else g for g, p in zip(vjp_params, f_params)])
                vjp_p = vjp_p.to(dtype)
            vjp_t = vjp_t.to(dtype)
            if group is not None:
                vjp_t = _all_reduce_sum(vjp_t.contiguous(), group)
                if len(f_params):
                    vjp_p = _all_reduce_sum(vjp_p.contiguous(), group)
            return (*(f.detach() for f in func_eval), *vjp_y, vjp_t, vjp_p)

        T = ans[0].shape[0]
        _rhs._FORCE_ACCURATE[0] += 1
        try:
            with torch.no_grad():
                adj_y = tuple(g[-1] for g in grad_output)                                     # adjoint.py:110-113
                adj_params = torch.zeros_like(flat_params, dtype=dtype) if flat_params.numel() else \
                    torch.zeros((), dtype=dtype, device=dev)
                adj_time = torch.zeros((), dtype=dtype, device=dev)
                time_vjps = []
                for i in range(T - 1, 0, -1):                                                 # adjoint.py:118
                    ans_i = tuple(a[i] for a in ans)
                    grad_i = tuple(g[i] for g in grad_output)
                    func_i = func(t[i].to(dtype), ans_i)
                    # effect of moving the current measurement time (adjoint.py:133-139)
                    dLd_cur_t = sum(torch.dot(f.reshape(-1), g.reshape(-1)).reshape(1) for f, g in zip(func_i, grad_i))
                    if group is not None:
                        dLd_cur_t = _all_reduce_sum(dLd_cur_t.contiguous(), group)
                    adj_time = adj_time - dLd_cur_t.reshape(())
                    time_vjps.append(dLd_cur_t)
                    aug_y0 = (*ans_i, *adj_y, adj_time, adj_params)                           # adjoint.py:146
                    aug_ans = odeint(augmented_dynamics, aug_y0, torch.stack([t[i], t[i - 1]]),
                                     rtol=opts["adjoint_rtol"], atol=opts["adjoint_atol"], method=opts["adjoint_method"],
                                     options=adj_options)                                     # adjoint.py:148-153
                    from . import solvers as _solvers
                    last_stats["backward"].append(dict(_solvers.last_stats))
                    adj_y = tuple(a[1] for a in aug_ans[n_tensors:2 * n_tensors])
                    adj_time = aug_ans[2 * n_tensors][1]
                    adj_params = aug_ans[2 * n_tensors + 1][1]
                    adj_y = tuple(a + g[i - 1] for a, g in zip(adj_y, grad_output))           # adjoint.py:164
                    del aug_y0, aug_ans
                time_vjps.append(adj_time.reshape(1))
                time_vjps = torch.cat(time_vjps[::-1]).to(t.dtype)                            # adjoint.py:169
                grad_params = adj_params if flat_params.numel() else None
        finally:
            _rhs._FORCE_ACCURATE[0] -= 1
        return (None, None, None, time_vjps, grad_params, *adj_y)


class _TupleFunc(nn.Module):
    """adjoint.py:205-212"""

    def __init__(self, base_func):
        super(_TupleFunc, self).__init__()
        self.base_func = base_func

    def forward(self, t, y):
        return (self.base_func(t, y[0]),)


class _FlatParamsGrad(torch.autograd.Function):
    """Routes the flat parameter gradient produced by the adjoint back onto the individual parameters."""

    @staticmethod
    def forward(ctx, *params):
        ctx.shapes = [p.shape for p in params]
        return _flatten(params) if params else torch.tensor([])

    @staticmethod
    def backward(ctx, g):
        out, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            out.append(g[off:off + n].reshape(shp))
            off += n
        return tuple(out)


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None, adjoint_method=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_options=None):
    """tfdiffeq/adjoint.py:183-224.  ``func`` must be an ``nn.Module`` (the reference demands a
    ``tf.keras.Model``, :187) so that its parameters can be found.

    The reference accepts ``adjoint_rtol`` / ``adjoint_atol`` but silently discards them and integrates the
    adjoint with ``rtol`` / ``atol`` (adjoint.py:18-19, :63-64).  That behaviour is reproduced.

    A state of n tensors integrates an augmented state of 2n + 2 components backwards; the engine carries at most
    ``B2ODE_MAXSEG`` = 12 components, i.e. n <= 5 (the reference has no such limit).  Tensor-core funcs
    (``rhs.DenseMLP`` / ``rhs.Conv2dODEFunc``) run both passes in their fp32-accurate mode so that the forward solve,
    the backward reconstruction of y and the autograd VJPs see the same dynamics.  With
    ``options={'shared_step_group': g}`` (batch shards on several GPUs) the returned parameter and time gradients are
    already summed over all shards and identical on every rank.
    """
    if not isinstance(func, nn.Module):
        raise ValueError('func is required to be an instance of nn.Module')
    if adjoint_method is None:
        adjoint_method = method
    if adjoint_options is None:
        adjoint_options = options
    tensor_input = False
    if isinstance(y0, torch.Tensor):
        tensor_input = True
        y0 = (y0,)
        func = _TupleFunc(func)
    params = tuple(p for p in func.parameters() if p.requires_grad)
    flat_params = _FlatParamsGrad.apply(*params) if params else torch.zeros(0, device=y0[0].device, dtype=y0[0].dtype)
    opts = dict(rtol=rtol, atol=atol, method=method, options=options, adjoint_method=adjoint_method,
                adjoint_rtol=rtol, adjoint_atol=atol, adjoint_options=adjoint_options)
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    t = t.to(y0[0].device)
    ys = _OdeintAdjoint.apply(func, len(y0), opts, t, flat_params, *y0)
    if tensor_input:
        ys = ys[0]
    return ys
