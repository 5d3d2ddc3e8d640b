"""A/B of libb2ode tuning variants on the north-star kernel size (65 536 x 128 fp64 dopri5, linear func).
Usage: B2ODE_LIB=path/to/lib.so python scripts/headline_ab.py   (prints one line per kernel family)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
dev = torch.device("cuda:0")
peak, _ = bench.peaks()
best = None
for rep in range(3):
    r = bench.headline_kernel_roofline(dev, peak)
    if best is None or r["per_kernel"]["finalize"]["avg_ms"] < best["per_kernel"]["finalize"]["avg_ms"]:
        best = r
print(os.environ.get("B2ODE_LIB", "default"), json.dumps({k: (round(v["avg_ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in best["per_kernel"].items()}))
