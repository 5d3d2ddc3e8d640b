"""Time one rhs.DenseMLP evaluation (b2ode_mlp3) for the library named by B2ODE_LIB; stage combine with NK k's."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfdiffeq_b200 as tfd
from tfdiffeq_b200 import _lib
dev = torch.device("cuda:0")
B, D, H = int(os.environ.get("B", 131072)), int(os.environ.get("D", 64)), int(os.environ.get("H", 256))
torch.manual_seed(0)
m = tfd.rhs.DenseMLP(D, H, "relu").to(dev)
x = torch.randn(B, D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = []
for NK in (0, 2, 5):
    ks = [torch.randn(B, D, device=dev) for _ in range(NK)]
    st = _lib.State(); st.dt = 0.01
    state = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
    stage = (ks, [0.1] * NK, state.data_ptr(), None) if NK else None
    for _ in range(3):
        tfd.rhs.mlp3(x, m.fc1, m.fc2, m.fc3, "relu", stage=stage)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tfd.rhs.mlp3(x, m.fc1, m.fc2, m.fc3, "relu", stage=stage); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res.append("nk=%d %.1f us" % (NK, ts[len(ts) // 2]))
print(os.path.basename(os.environ.get("B2ODE_LIB", "default")), " | ".join(res))
