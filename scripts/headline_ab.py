"""A/B of libb2ode variants on the north-star kernel size (65 536 x 128 fp64 dopri5, linear func): per-kernel-family
launch times from CUDA events recorded by the library around its own launches.
Usage:  [B2ODE_LIB=path/to/lib.so] [B2ODE_FINALIZE_BULK=1] python scripts/headline_ab.py   (one line per run)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

ctx = bench.Ctx(0, 1, 0)
peak, src = bench.hbm_peak()
best = None
for rep in range(3):
    r = bench.roofline_northstar(ctx, peak, src)
    if best is None or r["per_kernel"]["finalize"]["avg_ms"] < best["per_kernel"]["finalize"]["avg_ms"]:
        best = r
tag = "%s bulk=%s" % (os.environ.get("B2ODE_LIB", "default"), os.environ.get("B2ODE_FINALIZE_BULK", "0"))
print(tag, json.dumps({k: (round(v["avg_ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in best["per_kernel"].items()}),
      "acc/rej %s/%s" % (best["n_accepted"], best["n_rejected"]))
