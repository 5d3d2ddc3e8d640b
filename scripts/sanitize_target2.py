"""compute-sanitizer target for the kernels added after the first pass: tcgen05 dense layer, chained MLP (single CTA
and CTA pair), multistep lincomb / reduce.  Small sizes with ragged edges."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tfdiffeq_b200 as tfd
from problems import PROBLEMS
dev = torch.device("cuda:0")
torch.manual_seed(0)
for pair in ("0", "1"):
    os.environ["B2ODE_MLP3_PAIR"] = pair
    for (M, D, H) in ((300, 32, 64), (129, 48, 80), (700, 64, 256)):
        m = tfd.rhs.DenseMLP(D, H, "tanh").to(dev)
        y0 = torch.randn(M, D, device=dev)
        tfd.odeint(m, y0, torch.tensor([0., 0.3, 1.0]), rtol=1e-3, atol=1e-3, method="dopri5")
os.environ["B2ODE_MLP3_PAIR"] = "0"
m = tfd.rhs.DenseMLP(32, 512, "relu").to(dev)                     # per-layer kernels (width > 256)
tfd.odeint(m, torch.randn(200, 32, device=dev), torch.tensor([0., 1.0]), rtol=1e-3, atol=1e-3, method="dopri5")
rng = np.random.default_rng(0)
for dtype in (torch.float64, torch.float32):
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((333, 3)), dtype=dtype, device=dev)
    f = PROBLEMS["lorenz"](backend="torch", device=dev)
    t = torch.arange(31, dtype=torch.float64) * 0.005
    tfd.odeint(f, y0, t, method="explicit_adams", options=dict(max_order=5))
    tfd.odeint(f, y0, t, method="fixed_adams", rtol=1e-4, atol=1e-6)
    tfd.odeint(f, y0, t[:9], method="adams", rtol=1e-4, atol=1e-6)
    y = (torch.linspace(1., 2., 7, dtype=dtype, device=dev), torch.linspace(.5, 1.5, 33, dtype=dtype, device=dev))
    tfd.odeint(lambda t, yz: (-yz[0], -2.0 * yz[1]), y, torch.linspace(0., 1., 5), method="adams", rtol=1e-4, atol=1e-6)
torch.cuda.synchronize()
print("sanitize target 2 done", tfd.last_stats)
