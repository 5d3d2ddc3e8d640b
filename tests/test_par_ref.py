"""CPU: the process-parallel oracle (oracle/par_ref.py -- bench.py's all-cores CPU arm) takes the same step sequence
and gives the same values as the single-process oracle it wraps."""
import numpy as np
import pytest

import np_ref
import par_ref
from problems import PROBLEMS


@pytest.mark.parametrize("nproc", [1, 3])
def test_parallel_oracle_equals_single_process_oracle(nproc):
    rng = np.random.default_rng(0)
    y0 = np.array([1., 1., 1.]) + 0.1 * rng.standard_normal((1001, 3))        # uneven shards
    t = np.arange(41) * 0.01
    st = np_ref.Stats()
    ref = np_ref.odeint(PROBLEMS["lorenz"](backend="numpy"), y0, t, rtol=1e-7, atol=1e-9, method="dopri5", stats=st)
    r = par_ref.solve("lorenz", y0, t, nproc=nproc, reps=1, want_solution=True, rtol=1e-7, atol=1e-9, method="dopri5")
    assert (r["n_acc"], r["n_rej"], r["nfe"]) == (st.n_acc, st.n_rej, st.nfe)
    assert r["nproc"] == nproc and len(r["seconds"]) == 1 and r["seconds"][0] > 0
    # only the summation order of the mean differs (sum of per-shard sums)
    assert np.max(np.abs(r["solution"] - ref)) <= 1e-10
