"""TEST INFRASTRUCTURE ONLY -- a stand-in ``tensorflow`` module backed by torch-CPU.

TensorFlow is not installable in the build container (no network), so the
reference (`/root/reference/tfdiffeq/*.py`, pure Python on TF-Eager) cannot be
imported natively.  This module provides the ~45 ``tf.*`` symbols that the
reference's hot-path files use, with the TF semantics
that matter for parity kept faithful:

* ``convert_to_tensor(python float)`` -> **float32** (TF default), python int -> int32;
* ``reduce_max/min(list_of_tensors)`` -> stack, then a **global** reduction;
* ``python_scalar * tensor`` keeps the tensor dtype (torch already does this);
* ``math.add_n`` sums left to right;
* there is deliberately **no** ``tf.ceil`` (removed in TF2 -- `solvers.py:64` is
  broken there).

It is used ONLY by ``oracle/ref_loader.py`` (golden-vector generation and oracle
pinning, in the build container where /root/reference exists).  Nothing in the
product package imports it.
"""
import contextlib
import types

import numpy as np
import torch

_m = types.ModuleType("tensorflow")

Tensor = torch.Tensor


class _TFIntTensor(torch.Tensor):
    """``tf.range`` result: TF's true division promotes int32 operands to float64 (torch would give float32);
    adams.py:41 relies on it (``c = 1 / tf.range(1, k + 2)``)."""

    def __rtruediv__(self, other):
        return other / self.as_subclass(torch.Tensor).to(torch.float64)

    def __truediv__(self, other):
        return self.as_subclass(torch.Tensor).to(torch.float64) / other


def Variable(initial_value, trainable=True, dtype=None, name=None):  # noqa: N802
    """adams.py:34 only needs a mutable tensor whose slices can be assigned (``compat.assign(g[j], v)``)."""
    out = initial_value.clone() if isinstance(initial_value, torch.Tensor) else torch.as_tensor(initial_value)
    return out if dtype is None else out.to(dtype)

float16, float32, float64 = torch.float16, torch.float32, torch.float64
int32, int64 = torch.int32, torch.int64
bool = torch.bool  # noqa: A001  (mirrors tf.bool)


def _py_default_dtype(x):
    import builtins
    if isinstance(x, builtins.bool):
        return torch.bool
    if isinstance(x, int):
        return torch.int32
    if isinstance(x, float):
        return torch.float32
    return None


def convert_to_tensor(value, dtype=None, name=None):
    if isinstance(value, torch.Tensor):
        return value if dtype is None else value.to(dtype)
    if isinstance(value, np.ndarray):
        out = torch.from_numpy(value)
        return out if dtype is None else out.to(dtype)
    if isinstance(value, (list, tuple)):
        if len(value) > 0 and all(isinstance(v, torch.Tensor) for v in value):
            out = torch.stack([v for v in value])
            return out if dtype is None else out.to(dtype)
        if len(value) > 0 and any(isinstance(v, torch.Tensor) for v in value):
            tens = [v for v in value if isinstance(v, torch.Tensor)][0]
            out = torch.stack([v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=tens.dtype)
                               for v in value])
            return out if dtype is None else out.to(dtype)
        arr = np.asarray(value)
        if dtype is None:
            if arr.dtype == np.float64:
                arr = arr.astype(np.float32)       # TF: python floats -> float32
            elif arr.dtype == np.int64:
                arr = arr.astype(np.int32)
            return torch.from_numpy(arr)
        return torch.from_numpy(arr).to(dtype)
    d = dtype if dtype is not None else _py_default_dtype(value)
    if isinstance(value, (np.floating, np.integer)):
        d = dtype
        return torch.tensor(value.item(), dtype=d if d is not None else torch.from_numpy(np.asarray(value)).dtype)
    return torch.tensor(value, dtype=d)


def cast(x, dtype):
    if not isinstance(x, torch.Tensor):
        x = convert_to_tensor(x)
    return x.to(dtype)


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)):
        return convert_to_tensor(x)
    if isinstance(x, torch.Size):
        return torch.tensor(list(x), dtype=torch.int32)
    return convert_to_tensor(x)


def abs(x):  # noqa: A001
    return torch.abs(x)


def sqrt(x):
    return torch.sqrt(x)


def reduce_max(x, axis=None):
    return torch.max(_as_tensor(x))


def reduce_min(x, axis=None):
    return torch.min(_as_tensor(x))


def reduce_sum(x, axis=None):
    return torch.sum(_as_tensor(x))


def reduce_mean(x, axis=None):
    return torch.mean(_as_tensor(x))


def reduce_prod(x, axis=None):
    x = _as_tensor(x)
    if x.numel() == 0:
        return torch.tensor(1, dtype=torch.int32)
    return torch.prod(x)


def reduce_all(x, axis=None):
    return torch.all(_as_tensor(x))    # empty -> True, like TF


def reduce_any(x, axis=None):
    return torch.any(_as_tensor(x))


def norm(x):
    return torch.linalg.vector_norm(x.reshape(-1))


def stack(values, axis=0):
    return torch.stack(list(values), dim=axis)


def concat(values, axis=0):
    return torch.cat(list(values), dim=axis)


def reshape(x, shape):
    return x.reshape(list(shape))


def split(x, sizes):
    return list(torch.split(x, [int(s) for s in sizes]))


def maximum(a, b):
    a, b = _pair(a, b)
    return torch.maximum(a, b)


def minimum(a, b):
    a, b = _pair(a, b)
    return torch.minimum(a, b)


def _pair(a, b):
    if not isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        a = torch.tensor(a, dtype=b.dtype)
    if not isinstance(b, torch.Tensor) and isinstance(a, torch.Tensor):
        b = torch.tensor(b, dtype=a.dtype)
    return a, b


def multiply(a, b):
    return a * b


def zeros(shape, dtype=torch.float32):
    return torch.zeros(list(shape), dtype=dtype)


def zeros_like(x, dtype=None):
    return torch.zeros_like(x, dtype=dtype)


def identity(x):
    return x.clone()


def equal(a, b):
    return torch.eq(a, b) if isinstance(a, torch.Tensor) else torch.eq(b, a)


def range(*args, **kw):  # noqa: A001
    out = torch.arange(*args)
    return out.to(torch.int32).as_subclass(_TFIntTensor) if not out.dtype.is_floating_point else out


def is_tensor(x):
    return isinstance(x, torch.Tensor)


@contextlib.contextmanager
def device(name):
    yield


def executing_eagerly():
    return True


_math = types.ModuleType("tensorflow.math")
_math.add_n = lambda xs: _add_n(xs)
_math.is_nan = torch.isnan
_math.is_inf = torch.isinf
_math.abs = torch.abs


def _add_n(xs):
    xs = list(xs)
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


_debugging = types.ModuleType("tensorflow.debugging")
_debugging.is_numeric_tensor = lambda x: isinstance(x, torch.Tensor) and (
    x.dtype.is_floating_point or x.dtype.is_complex or x.dtype in (torch.int8, torch.int16, torch.int32,
                                                                  torch.int64, torch.uint8))

_version = types.ModuleType("tensorflow.version")
_version.VERSION = "2.0.0-b200-shim"


def install():
    """Register this module as ``tensorflow`` in sys.modules (idempotent)."""
    import sys
    g = globals()
    for name in ("Tensor", "Variable", "float16", "float32", "float64", "int32", "int64", "bool",
                 "convert_to_tensor", "cast", "abs", "sqrt", "reduce_max", "reduce_min", "reduce_sum",
                 "reduce_mean", "reduce_prod", "reduce_all", "reduce_any", "norm", "stack", "concat",
                 "reshape", "split", "maximum", "minimum", "multiply", "zeros", "zeros_like", "identity",
                 "equal", "range", "is_tensor", "device", "executing_eagerly"):
        setattr(_m, name, g[name])
    _m.math = _math
    _m.debugging = _debugging
    _m.version = _version
    sys.modules["tensorflow"] = _m
    sys.modules["tensorflow.math"] = _math
    return _m
