"""Determinism stress: the same evaluation repeated many times must give bit-identical results (a race in the
tcgen05 / mbarrier plumbing shows up as a rare mismatch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfdiffeq_b200 as tfd
from tfdiffeq_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
REPS = int(os.environ.get("REPS", 300))
bad = 0
for (M, D, H, act) in ((2000, 64, 128, "tanh"), (131072, 64, 256, "relu"), (5000, 32, 64, "relu"), (37965, 64, 128, "tanh"), (4096, 256, 256, "relu")):
    m = tfd.rhs.DenseMLP(D, H, act).to(dev)
    x = torch.randn(M, D, device=dev)
    ks = [torch.randn(M, D, device=dev) for _ in range(3)]
    st = _lib.State(); st.dt = 0.02
    state = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)
    for name, fn in (("chained", lambda: tfd.rhs.mlp3(x, m.fc1, m.fc2, m.fc3, act, stage=(ks, [0.3, -0.2, 0.1], state.data_ptr(), None))),
                     ("per-layer", lambda: tfd.rhs.dense_layer(tfd.rhs.dense_layer(tfd.rhs.dense_layer(x, m.fc1.weight, m.fc1.bias, act, stage=(ks, [0.3, -0.2, 0.1], state.data_ptr(), None)), m.fc2.weight, m.fc2.bias, act), m.fc3.weight, m.fc3.bias, "none"))):
        ref = fn().clone()
        mism = 0
        for i in range(REPS):
            out = fn()
            if not torch.equal(out, ref):
                mism += 1
        torch.cuda.synchronize()
        print("%-10s M=%6d D=%3d H=%3d %s: %d / %d repeats differ" % (name, M, D, H, act, mism, REPS))
        bad += mism
print("TOTAL mismatches", bad)

# the same through odeint (chained and per-layer): counts and results must repeat exactly
m = tfd.rhs.DenseMLP(64, 128, "tanh").to(dev)
y0 = torch.randn(2000, 64, device=dev)
t = torch.tensor([0., 1.0])
for chain in (True, False):
    m.chain = chain
    ref, refst = None, None
    diffs = 0
    for i in range(40):
        out = tfd.odeint(m, y0, t, rtol=1e-4, atol=1e-4, method="dopri5")
        st = (tfd.last_stats["n_accepted"], tfd.last_stats["n_rejected"], tfd.last_stats["nfe"])
        if ref is None:
            ref, refst = out.clone(), st
        elif not torch.equal(out, ref) or st != refst:
            diffs += 1
            print("  rep %d differs: stats %s vs %s, max|d| %.3e" % (i, st, refst, float((out - ref).abs().max())))
    print("odeint chain=%s: %d / 40 repeats differ  stats %s" % (chain, diffs, refst))
