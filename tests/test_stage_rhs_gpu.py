"""GPU: stage kernels with a built-in right-hand side fused in (b2ode_rk_stage_rhs / b2ode_rhs_eval, SURVEY 8f-2 for
batches the persistent kernel cannot hold): same step sequence and values as the module called as an ordinary func, as
the persistent kernel, and as the oracle; the Kepler right-hand side of BASELINE config 5."""
import numpy as np
import pytest
import torch

import np_ref
from golden_util import max_rel_err
from problems import PROBLEMS

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def tfd():
    import tfdiffeq_b200
    return tfdiffeq_b200


@pytest.mark.parametrize("method,kw", [("dopri5", {}), ("dopri8", dict(rtol=1e-9, atol=1e-9)), ("bosh3", dict(rtol=1e-5, atol=1e-7)),
                                       ("tsit5", dict(rtol=1e-2, atol=1e-2)), ("adaptive_heun", dict(rtol=1e-4, atol=1e-6))])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_stage_rhs_lorenz_equals_func_path(method, kw, dtype):
    rng = np.random.default_rng(2)
    y0 = torch.tensor(np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((777, 3)), device=DEV, dtype=dtype)
    t = torch.arange(31, dtype=torch.float64) * 0.01
    f = tfd().rhs.Lorenz()
    if dtype == torch.float32 and method != "tsit5":
        kw = dict(rtol=1e-4, atol=1e-6)
    a = tfd().odeint(f, y0, t, method=method, options={"fused_rhs": "stages"}, **kw)
    sa = dict(tfd().last_stats)
    assert sa["stage_rhs"] and not sa["fused_rhs"]
    b = tfd().odeint(f, y0, t, method=method, options={"fused_rhs": False}, **kw)
    sb = dict(tfd().last_stats)
    assert not sb["stage_rhs"]
    assert (sa["n_accepted"], sa["n_rejected"], sa["nfe"]) == (sb["n_accepted"], sb["n_rejected"], sb["nfe"])
    # the kernel evaluates the same IEEE operations in the same order as the module's torch forward
    assert torch.equal(a, b)
    # CUDA-graph replay of the attempt (all launches are the library's own)
    c = tfd().odeint(f, y0, t, method=method, options={"fused_rhs": "stages", "cuda_graph": True}, **kw)
    assert torch.equal(a, c)


def test_kepler_rhs_three_paths_and_oracle():
    """BASELINE config 5's system in miniature: 32 orbits per row (dim 128), dopri8 at 1e-9."""
    k_np = PROBLEMS["kepler"](backend="numpy")
    y0 = k_np.y0(64, seed=3)
    t = np.linspace(0., 2., 5)
    st = np_ref.Stats()
    ref = np_ref.odeint(k_np, y0, t, rtol=1e-9, atol=1e-9, method="dopri8", stats=st)
    f = tfd().rhs.Kepler()
    y = torch.tensor(y0, device=DEV)
    tt = torch.tensor(t)
    kw = dict(rtol=1e-9, atol=1e-9, method="dopri8")
    res = {}
    for name, opt in (("persistent", True), ("stages", "stages"), ("func", False)):
        res[name] = tfd().odeint(f, y, tt, options={"fused_rhs": opt}, **kw)
        s = dict(tfd().last_stats)
        assert (s["n_accepted"], s["n_rejected"]) == (st.n_acc, st.n_rej), (name, s)
        assert max_rel_err(res[name].cpu().numpy(), ref) <= 1e-6, name
    assert float((res["stages"] - res["func"]).abs().max()) <= 1e-9
    assert float((res["persistent"] - res["func"]).abs().max()) <= 1e-9
    # reverse time and the fixed grid
    tr = torch.tensor(t[::-1].copy())
    a = tfd().odeint(f, y, tr, options={"fused_rhs": "stages"}, **kw)
    b = tfd().odeint(f, y, tr, options={"fused_rhs": False}, **kw)
    assert float((a - b).abs().max()) <= 1e-9
    g1 = tfd().odeint(f, y, torch.tensor(np.linspace(0., 1., 41)), method="rk4")
    g2 = tfd().odeint(f, y, torch.tensor(np.linspace(0., 1., 41)), method="rk4", options={"fused_rhs": False})
    assert float((g1 - g2).abs().max()) <= 1e-12


def test_oversized_batch_takes_the_stage_rhs_path_and_matches_oracle():
    """More trajectories than the persistent kernel can keep co-resident: a warning, then one launch per stage."""
    rng = np.random.default_rng(4)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((200000, 3))
    t = np.arange(6) * 0.01
    st = np_ref.Stats()
    ref = np_ref.odeint(PROBLEMS["lorenz"](backend="numpy"), y0, t, method="dopri5", stats=st)
    with pytest.warns(RuntimeWarning):
        got = tfd().odeint(tfd().rhs.Lorenz(), torch.tensor(y0, device=DEV), torch.tensor(t), method="dopri5")
    s = dict(tfd().last_stats)
    assert s["stage_rhs"] and (s["n_accepted"], s["n_rejected"], s["nfe"]) == (st.n_acc, st.n_rej, st.nfe)
    assert max_rel_err(got.cpu().numpy(), ref) <= 1e-9
