"""Timeline of the persistent fused kernel (k_fused_adaptive) on BASELINE config 2: cycles per phase of an attempt.

    make -C tfdiffeq_b200/csrc trace && B2ODE_LIB=tfdiffeq_b200/libb2ode_trace.so python scripts/fused_trace.py

Stamps are clock64() of the control warp and of compute warp 0 in the first and the last block (see the table printed
below).  Round-1 kernel for comparison (profiles/r02_fused_trace.md): 15.7k cycles per attempt."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfdiffeq_b200 as tfd  # noqa: E402
from tfdiffeq_b200 import _lib  # noqa: E402

B = int(os.environ.get("B", 65536))
rng = np.random.default_rng(0)
y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((B, 3)), device="cuda")
t = torch.arange(1000, dtype=torch.float64) * 0.01
f = tfd.rhs.Lorenz()
for _ in range(2):
    tfd.odeint(f, y0, t, method="dopri5")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
tfd.odeint(f, y0, t, method="dopri5")
b.record()
torch.cuda.synchronize()
st = dict(tfd.last_stats)
att = st["n_accepted"] + st["n_rejected"]
print("solve %.3f ms, %d attempts, %.2f us/attempt" % (a.elapsed_time(b), att, a.elapsed_time(b) * 1e3 / att))
raw = C.CDLL(_lib.LIB_PATH)
if not hasattr(raw, "b2ode_debug_fused_trace"):
    sys.exit("library was not built with -DB2ODE_FUSED_TRACE")
NA, NP = 64, 16
buf = (C.c_ulonglong * (2 * NA * NP))()
assert raw.b2ode_debug_fused_trace(buf) == 0
tr = np.frombuffer(buf, dtype=np.uint64).reshape(2, NA, NP).astype(np.int64)
# control warp (lane 0): 0 loop top, 1 all compute warps' partials in, 2 block total, 3 leader: all blocks gathered /
# others: own partial published, 4 group total known, 5 decision published
# compute warp 0 (lane 0): 8 partial handed over, 9 speculative dense output written, 10 decision received
for blk in range(2):
    x = tr[blk]
    tot = x[1:, 0] - x[:-1, 0]
    print("block %s: cycles per attempt median %d (%.2f us at 1.965 GHz)" % ("first (leader)" if blk == 0 else "last", np.median(tot),
                                                                           np.median(tot) / 1965.0))
    rows = [("compute: decision -> stages+err -> hand-over", x[1:, 8] - x[:-1, 10]),
            ("compute: speculative dense output", x[:, 9] - x[:, 8]),
            ("compute: hand-over -> decision received", x[:, 10] - x[:, 8]),
            ("control: warp-0 hand-over -> all partials in", x[:, 1] - x[:, 8]),
            ("control: block reduce", x[:, 2] - x[:, 1]),
            ("control: gather (leader) / publish (others)", x[:, 3] - x[:, 2]),
            ("control: -> group total known", x[:, 4] - x[:, 3]),
            ("control: controller + publish", x[:, 5] - x[:, 4]),
            ("decision published -> received by warp 0", x[:, 10] - x[:, 5])]
    for n, d in rows:
        print("   %-46s median %6d   p90 %6d" % (n, np.median(d), np.percentile(d, 90)))
