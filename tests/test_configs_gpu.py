"""GPU: BASELINE.json's configs 3, 4 and 5 at (or near) their full sizes, through size-independent properties
and oracle comparisons on what the oracle can finish in seconds."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import np_ref
from golden_util import max_rel_err
from problems import PROBLEMS

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def tfd():
    import tfdiffeq_b200
    return tfdiffeq_b200


def test_config3_full_size_rk4_mlp_fp32_subset_equals_oracle():
    """Config 3: spiral MLP 2 -> 50 -> 2, batch 131 072 fp32, rk4 on 2 000 grid cells.  A fixed grid has no
    step-size coupling, so every trajectory evolves independently: a random subset of the full-batch result
    must match the oracle run on that subset alone."""
    rng = np.random.default_rng(3)
    y0 = (np.array([2., 0.]) + 0.1 * rng.standard_normal((131072, 2))).astype(np.float32)
    t = np.linspace(0., 25., 2001).astype(np.float32)
    f_gpu = PROBLEMS["spiral_mlp"](backend="torch", dtype="float32", device=DEV)
    f_cpu = PROBLEMS["spiral_mlp"](backend="numpy", dtype="float32")
    sol = tfd().odeint(f_gpu, torch.tensor(y0, device=DEV), torch.tensor(t), method="rk4")
    assert sol.shape == (2001, 131072, 2) and sol.dtype == torch.float32
    assert tfd().last_stats["nfe"] == 8000
    idx = rng.choice(131072, size=48, replace=False)
    ref = np_ref.odeint(f_cpu, y0[idx], t, method="rk4")
    got = sol[:, torch.tensor(idx, device=DEV)].cpu().numpy()
    assert max_rel_err(got, ref) <= 1e-3
    assert bool(torch.isfinite(sol).all())


def test_config5_dopri8_dim128_fp64_vs_oracle():
    """Config 5: dopri8, dim 128, rtol = atol = 1e-9, stiff-ish tridiagonal (DETEST C-class), full per-GPU batch
    of the 4-GPU split (4 096 x 128), short horizon so the oracle finishes in seconds."""
    rng = np.random.default_rng(5)
    y0 = np.zeros((4096, 128))
    y0[:, 0] = 1.0
    y0 += 0.01 * rng.standard_normal(y0.shape)
    t = np.linspace(0., 0.6, 4)
    fn = PROBLEMS["tridiag"](backend="numpy", dim=128)
    ft = PROBLEMS["tridiag"](backend="torch", device=DEV, dim=128)
    st = np_ref.Stats()
    ref = np_ref.odeint(fn, y0, t, method="dopri8", rtol=1e-9, atol=1e-9, stats=st)
    got = tfd().odeint(ft, torch.tensor(y0, device=DEV), torch.tensor(t), method="dopri8", rtol=1e-9, atol=1e-9)
    s = tfd().last_stats
    assert max_rel_err(got.cpu().numpy(), ref) <= 1e-6
    assert abs(s["n_accepted"] - st.n_acc) <= 1 and abs(s["n_rejected"] - st.n_rej) <= 1


class ConvODEFunc(nn.Module):
    """The shape of the reference's Conv2dODEFunc (tfdiffeq/models/conv_odenet.py:80-86,139-143):
    conv1x1 -> relu -> conv3x3 (same) -> relu -> conv1x1, channels kept."""

    def __init__(self, ch, dtype):
        super().__init__()
        self.c1 = nn.Conv2d(ch, ch, 1, dtype=dtype)
        self.c2 = nn.Conv2d(ch, ch, 3, padding=1, dtype=dtype)
        self.c3 = nn.Conv2d(ch, ch, 1, dtype=dtype)

    def forward(self, t, x):
        return self.c3(torch.relu(self.c2(torch.relu(self.c1(x)))))


def test_config4_conv_odefunc_forward_vs_oracle_and_adjoint_gradient():
    """Config 4 in miniature: ConvODEFunc on (B, C, H, W) feature maps, dopri5 rtol = atol = 1e-3 with
    max_num_steps = 1000 (dense_odenet.py:126-127), forward against the oracle (torch-CPU arrays, same module
    weights), backward through odeint_adjoint against a finite difference."""
    torch.manual_seed(0)
    m = ConvODEFunc(8, torch.float64).to(DEV)
    m_cpu = ConvODEFunc(8, torch.float64)
    m_cpu.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    x0 = torch.randn(16, 8, 12, 12, dtype=torch.float64)
    t = torch.tensor([0., 1.])
    kw = dict(rtol=1e-3, atol=1e-3, method="dopri5", options=dict(max_num_steps=1000))
    with torch.no_grad():
        st = np_ref.Stats()
        ref = np_ref.odeint(lambda tt, y: m_cpu(torch.as_tensor(float(tt)), y), x0, t.numpy(), rtol=1e-3, atol=1e-3,
                            method="dopri5", options=dict(max_num_steps=1000), stats=st)
        got = tfd().odeint(m, x0.to(DEV), t, **kw)
    s = dict(tfd().last_stats)
    assert (s["n_accepted"], s["n_rejected"]) == (st.n_acc, st.n_rej)
    assert max_rel_err(got.cpu().numpy(), ref.numpy()) <= 1e-6
    # adjoint backward (reference ODEBlock(adjoint=True) raises, dense_odenet.py:116: the func is used directly)
    x = x0.to(DEV).requires_grad_(True)
    tight = dict(rtol=1e-10, atol=1e-10, method="dopri5")
    out = tfd().odeint_adjoint(m, x, t, **tight)
    loss = (out[-1] ** 2).mean()
    loss.backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())
    # directional derivative along a random direction of the 3x3 conv weights (one entry alone is below the
    # noise floor of a finite difference of two 1e-10-accurate solves)
    w = m.c2.weight
    v = torch.randn_like(w)
    v /= v.norm()
    g_dir = float((w.grad * v).sum())

    def fwd():
        with torch.no_grad():
            return float((tfd().odeint(m, x0.to(DEV), t, **tight)[-1] ** 2).mean())
    w0 = w.data.clone()
    eps = 1e-4
    w.data = w0 + eps * v
    lp = fwd()
    w.data = w0 - eps * v
    lm = fwd()
    w.data = w0
    fd = (lp - lm) / (2 * eps)
    # ReLU kinks make the finite difference itself O(eps)-inaccurate; smooth funcs are checked to 5e-6 in test_adjoint_gpu.py
    assert abs(fd - g_dir) <= 5e-3 * max(abs(fd), 1e-3), (fd, g_dir)


# ----------------------------------------------------------------------------------------------------------------
# BASELINE config 2 in full: 65 536 x 3 fp64 Lorenz, dopri5 defaults, all 1 000 output points (what bench.py times)
# ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg2_oracle():
    """One full oracle solve (about half a minute of numpy), shared by the three engine paths below."""
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((65536, 3))
    t = np.arange(1000) * 0.01
    st = np_ref.Stats()
    ref = np_ref.odeint(PROBLEMS["lorenz"](backend="numpy"), y0, t, rtol=1e-7, atol=1e-9, method="dopri5", stats=st)
    assert (st.n_acc, st.n_rej, st.nfe) == (390, 54, 2666)            # SURVEY 8(c) known-answer vector
    return y0, t, ref, st


@pytest.mark.parametrize("path", ["fused_rhs", "external_func_eager", "external_func_cuda_graph"])
def test_config2_full_batch_full_horizon_vs_oracle(cfg2_oracle, path):
    """The benchmarked solve itself: same accepted / rejected / NFE counts as the oracle over the WHOLE horizon, values
    within north_star's 1e-6 fp64 bar for t <= 5 (Lorenz is chaotic, parity horizon t <= 10; the tail
    t in (5, 10) is checked at 1e-4, which a 1-ulp difference in pow() amplified by e^{0.9 t} stays far below)."""
    import tfdiffeq_b200 as tfd
    y0, t, ref, st = cfg2_oracle
    if path == "fused_rhs":
        f, opts = tfd.rhs.Lorenz(), {}
    else:
        f, opts = PROBLEMS["lorenz"](backend="torch", device=DEV), ({"cuda_graph": True} if path.endswith("graph") else {})
    got = tfd.odeint(f, torch.tensor(y0, device=DEV), torch.tensor(t), rtol=1e-7, atol=1e-9, method="dopri5", options=opts)
    s = dict(tfd.last_stats)
    assert s["fused_rhs"] == (path == "fused_rhs")
    assert (s["n_accepted"], s["n_rejected"], s["nfe"]) == (st.n_acc, st.n_rej, st.nfe) == (390, 54, 2666)
    got = got.cpu().numpy()
    scale = np.maximum(np.abs(ref), 1.0)
    rel = np.abs(got - ref) / scale
    head = float(rel[:501].max())          # t <= 5.0
    tail = float(rel.max())                # t <= 9.99
    print("cfg2 %s: max rel err t<=5: %.3e, whole horizon: %.3e" % (path, head, tail))
    assert head <= 1e-6, head
    assert tail <= 1e-4, tail
