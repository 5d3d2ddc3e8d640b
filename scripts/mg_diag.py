"""torchrun --nproc-per-node N scripts/mg_diag.py : per-rank timeline of repeated sharded fused solves (cfg2)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfdiffeq_b200 as tfd  # noqa: E402
from tfdiffeq_b200 import _lib  # noqa: E402
from tfdiffeq_b200.comm import SharedStepGroup  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
group = SharedStepGroup()
y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * np.random.default_rng(rank).standard_normal((65536, 3)), device=dev)
t = torch.arange(1000, dtype=torch.float64) * 0.01
f = tfd.rhs.Lorenz()
opts = {"shared_step_group": group}
for _ in range(3):
    tfd.odeint(f, y0, t, method="dopri5", options=opts)
_lib.check(_lib.lib.b2ode_timing_enable(1 << _lib.FAM_FUSED))
dist.barrier()
torch.cuda.synchronize()
rows = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
FLUSH = os.environ.get("FLUSH", "0") == "1"
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    if FLUSH:
        flush.fill_(i)
    a.record()
    tfd.odeint(f, y0, t, method="dopri5", options=opts)
    b.record()
    torch.cuda.synchronize()
    rows.append((a.elapsed_time(b), 1e3 * (time.perf_counter() - w0)))
ms, cnt = C.c_double(), C.c_int()
_lib.check(_lib.lib.b2ode_timing_read(_lib.FAM_FUSED, C.byref(ms), C.byref(cnt)))
print("rank %d: kernel avg %.3f ms over %d launches; per solve (event ms, wall ms): %s" % (
    rank, ms.value / max(cnt.value, 1), cnt.value, " ".join("(%.2f, %.2f)" % r for r in rows)), flush=True)
group.close()
dist.destroy_process_group()
