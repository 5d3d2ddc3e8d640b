"""ODENet-style func (fc-relu-fc-relu-fc) inside dopri5: torch fp32 matmul vs torch TF32 matmul vs the tcgen05
dense layer vs tcgen05 + stage combine fused into layer 1.  One solve = t in [0, 1], rtol = atol = 1e-3."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfdiffeq_b200 as tfd
dev = torch.device("cuda:0")
B, D, H = int(os.environ.get("B", 131072)), int(os.environ.get("D", 64)), int(os.environ.get("H", 256))
torch.manual_seed(0)
m = tfd.rhs.DenseMLP(D, H, "relu").to(dev)
y0 = torch.randn(B, D, device=dev)
t = torch.tensor([0., 1.])
kw = dict(rtol=1e-3, atol=1e-3, method="dopri5")


def timeit(name, setup, opts):
    setup()
    for _ in range(2):
        tfd.odeint(m, y0, t, options=dict(opts), **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = tfd.odeint(m, y0, t, options=dict(opts), **kw)
    b.record(); torch.cuda.synchronize()
    st = tfd.last_stats
    att = st["n_accepted"] + st["n_rejected"]
    ms = a.elapsed_time(b) / 5
    flop = 2.0 * B * (D * H + H * H + H * D) * st["nfe"]
    print("%-46s %8.3f ms/solve  %2d attempts  %7.1f us/attempt  %6.1f TFLOP/s(func)" % (name, ms, att, 1e3 * ms / att, flop / ms / 1e9))
    return out


def torch32():
    m.tensor_cores = False; torch.backends.cuda.matmul.allow_tf32 = False
def torchtf32():
    m.tensor_cores = False; torch.backends.cuda.matmul.allow_tf32 = True
def ours():
    m.tensor_cores = True; torch.backends.cuda.matmul.allow_tf32 = False
print("B=%d dim=%d hidden=%d" % (B, D, H))
r0 = timeit("torch fp32 matmul func", torch32, {})
r1 = timeit("torch TF32 matmul func (allow_tf32)", torchtf32, {})
def perlayer():
    m.tensor_cores = True; torch.backends.cuda.matmul.allow_tf32 = False; m.chain = False
def chained():
    m.tensor_cores = True; torch.backends.cuda.matmul.allow_tf32 = False; m.chain = True
r2 = timeit("tcgen05 per-layer, separate stage kernel", perlayer, {"fused_rhs": False})
r2b = timeit("tcgen05 per-layer + fused stage combine", perlayer, {})
ours = chained
r3 = timeit("tcgen05 chained mlp3 + fused stage combine", ours, {})
r4 = timeit("   ... + CUDA-graph replay", ours, {"cuda_graph": True})
print("max|tc - fp32| = %.3e   max|torchTF32 - fp32| = %.3e   scale %.3f" % (float((r3 - r0).abs().max()), float((r1 - r0).abs().max()), float(r0.abs().max())))

# func alone
x = torch.randn(B, D, device=dev)
for name, ch in (("per-layer", False), ("chained mlp3", True)):
    m.tensor_cores = True; m.chain = ch
    with torch.no_grad():
        for _ in range(3): m(0.0, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): m(0.0, x)
        b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print("func eval %-14s %8.1f us   %6.1f TFLOP/s" % (name, us, 2.0 * B * (D * H + H * H + H * D) / us / 1e6))
