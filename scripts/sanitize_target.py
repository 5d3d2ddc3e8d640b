"""Small end-to-end sequence for compute-sanitizer (memcheck / racecheck): every kernel family, odd sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tfdiffeq_b200 as tfd
from problems import PROBLEMS
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for dtype in (torch.float64, torch.float32):
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((333, 3)), dtype=dtype, device=dev)
    t = torch.arange(11, dtype=torch.float64) * 0.01
    f = PROBLEMS["lorenz"](backend="torch", device=dev)
    for m, kw in (("dopri5", {}), ("dopri8", dict(rtol=1e-6, atol=1e-6)), ("adaptive_heun", dict(rtol=1e-2, atol=1e-3)),
                  ("tsit5", dict(rtol=1e-2, atol=1e-2)), ("rk4", {}), ("euler", {}), ("midpoint", {}), ("heun", {})):
        tfd.odeint(f, y0, t, method=m, **kw)
        tfd.odeint(tfd.rhs.Lorenz(), y0, t, method=m, **kw)
    tfd.odeint(f, y0, t, method="dopri5", options=dict(cuda_graph=True))
    # tuple state with odd segment lengths (scalar tails, several segments per launch)
    y = (torch.linspace(1., 2., 7, dtype=dtype, device=dev), torch.linspace(.5, 1.5, 33, dtype=dtype, device=dev),
         torch.ones((), dtype=dtype, device=dev))
    tfd.odeint(lambda t, yz: (-yz[0], -2.0 * yz[1], -yz[2]), y, torch.linspace(0., 1., 4), method="dopri5")
    m = tfd.rhs.CubicMLP(hidden=50, dtype=dtype).to(dev)
    tfd.odeint(m, y0[:, :2].contiguous(), t, method="rk4")
    tfd.odeint(m, y0[:, :2].contiguous(), t, method="dopri5", rtol=1e-4, atol=1e-5)
torch.cuda.synchronize()
print("sanitize target done", tfd.last_stats)
