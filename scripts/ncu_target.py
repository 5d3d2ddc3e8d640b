"""Short, deterministic kernel sequence for ncu (never a bench number).

  --workload lorenz    : BASELINE config 2 shape (65 536 x 3 fp64 dopri5), first --npts output points, external func
  --workload fused     : the same system with the library's own right-hand side: the persistent kernel, all 1 000 points
  --workload headline  : north-star kernel size (65 536 x 128 fp64 dopri5, linear func), 3 output points
  --workload mlp       : ODENet func rhs.DenseMLP(64, 256) on 131 072 rows (fp32 / TF32 tcgen05), one dopri5 solve
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import tfdiffeq_b200 as tfd  # noqa: E402
from problems import PROBLEMS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="lorenz")
ap.add_argument("--npts", type=int, default=40)
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.workload == "lorenz":
    rng = np.random.default_rng(0)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((65536, 3)), device=dev)
    f = PROBLEMS["lorenz"](backend="torch", device=dev)
    t = torch.arange(a.npts, dtype=torch.float64) * 0.01
    kw = dict(method="dopri5")
elif a.workload == "fused":
    rng = np.random.default_rng(0)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((65536, 3)), device=dev)
    f = tfd.rhs.Lorenz()
    t = torch.arange(1000, dtype=torch.float64) * 0.01
    kw = dict(method="dopri5")
elif a.workload == "mlp":
    torch.manual_seed(0)
    f = tfd.rhs.DenseMLP(64, 256, "relu").to(dev)
    y0 = torch.randn(131072, 64, device=dev)
    t = torch.tensor([0., 1.])
    kw = dict(method="dopri5", rtol=1e-3, atol=1e-3)
else:
    torch.manual_seed(0)
    y0 = torch.randn(65536, 128, dtype=torch.float64, device=dev)
    A = -0.5 * torch.eye(128, dtype=torch.float64, device=dev) + 0.05 * torch.randn(128, 128, dtype=torch.float64, device=dev)
    f = lambda t, y: y @ A   # noqa: E731
    t = torch.linspace(0., 0.5, 3, dtype=torch.float64)
    kw = dict(method="dopri5", rtol=1e-6, atol=1e-9)
for _ in range(a.repeat):
    tfd.odeint(f, y0, t, **kw)
    torch.cuda.synchronize()
print(a.workload, tfd.last_stats)
