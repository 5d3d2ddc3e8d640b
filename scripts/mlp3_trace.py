"""Timeline of k_mlp3_tf32's block 0 (needs a library built with NVCC_EXTRA=-DB2ODE_TRACE): per tile, the
%globaltimer stamp of each hand-off between the warp roles, printed relative to the tile's first event."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfdiffeq_b200 as tfd
from tfdiffeq_b200 import _lib
dev = torch.device("cuda:0")
B, D, H, NK = int(os.environ.get("B", 131072)), int(os.environ.get("D", 64)), int(os.environ.get("H", 256)), int(os.environ.get("NK", 5))
m = tfd.rhs.DenseMLP(D, H, "relu").to(dev)
x = torch.randn(B, D, device=dev)
ks = [torch.randn(B, D, device=dev) for _ in range(NK)]
st = _lib.State(); st.dt = 0.01
state = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
stage = (ks, [0.1] * NK, state.data_ptr(), None) if NK else None
for _ in range(3):
    tfd.rhs.mlp3(x, m.fc1, m.fc2, m.fc3, "relu", stage=stage)
torch.cuda.synchronize()
raw = C.CDLL(_lib.lib._name)
buf = (C.c_ulonglong * 336)()
raw.b2ode_debug_mlp3_trace(buf)
names = ["in regs", "actfree", "a1 arr", "g1 start", "g1 issued", "g2 start", "g2 issued", "g3 start", "g3 issued",
         "t1 seen", "a2 arr", "t2 seen", "a3 arr", "t3 seen", "out done", "-"]
t00 = min(v for v in buf[:256] if v)
for t in range(8):
    row = list(buf[t * 16:(t + 1) * 16])
    print("tile %d: " % t + "  ".join("%s %.2f" % (names[i], (row[i] - t00) / 1e3) for i in range(15) if row[i]))

clk = list(buf[256:336])
print("activation epilogue, warp q=0 (cycles): per K block  [ld+wait, math+stores, fences]")
for i in range(0, 64, 4):
    a, b, c, d = clk[i:i + 4]
    nxt = clk[i + 4]
    print("  kb %2d: %5d %5d %5d   (to next iteration start %5d)" % (i // 4, b - a, c - b, d - c, nxt - d))
print("output epilogue: [ld+wait, transpose to smem, global stores]")
for i in range(64, 72, 4):
    a, b, c, d = clk[i:i + 4]
    print("  c0 %2d: %5d %5d %5d" % ((i - 64) // 4 * 32, b - a, c - b, d - c))
