"""Turn ncu outputs in gpurun_out/ into the small tracked summaries under profiles/ (read here, no GPU needed).

  python scripts/summarize_ncu.py launches gpurun_out/launches_lorenz.csv profiles/r01_launches_lorenz.md
  python scripts/summarize_ncu.py full gpurun_out/prof_stage_headline.ncu-rep profiles/r01_stage_headline.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    ours = sum(sum(v) for k, v in agg.items() if k.startswith("void k_") or k.startswith("k_") or " k_" in k)
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): %s\n\n" % src)
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("total %.1f us over %d launches; libb2ode kernels %.1f%% of GPU time, func's torch kernels %.1f%%\n\n" % (
            tot / 1e3, sum(len(v) for v in agg.values()), 100 * ours / tot, 100 * (1 - ours / tot)))
        f.write("| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.2f | %.1f%% |\n" % (k[:110].replace("|", "/"), len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: %s\n\n" % src)
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            f.write("## `%s`  grid %s x block %s\n\n" % (d.get("Kernel Name"), d.get("launch__grid_size"), d.get("launch__block_size")))
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in d:
                    f.write("| %s | %s | %s |\n" % (m, d[m], u[m]))
            try:
                t_us = float(d["gpu__time_duration.sum"])
                if u["gpu__time_duration.sum"] == "ns":
                    t_us /= 1e3
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                tb = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
                f.write("| **dram traffic (read+write)** | %.1f | MB |\n| **traffic / duration** | %.0f | GB/s |\n" % (
                    tb / 1e6, tb / (t_us * 1e-6) / 1e9))
            except (KeyError, ValueError):
                pass
            f.write("\n")


def traffic(src, key, match):
    """Record dram bytes per launch of the first kernel whose name contains `match` in profiles/ncu_traffic.json under
    `key` (read by bench.py's roofline.traffic, with the report it came from)."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        if match in d.get("Kernel Name", ""):
            tb = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] + \
                float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
            db = json.load(open(path)) if os.path.exists(path) else {}
            db[key] = {"dram_bytes": tb, "kernel": d["Kernel Name"][:120], "source": "ncu --set full --clock-control none: " + os.path.basename(src),
                       "duration_us_under_ncu": float(d["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u["gpu__time_duration.sum"], 1.0)}
            json.dump(db, open(path, "w"), indent=1, sort_keys=True)
            print(key, db[key])
            return
    raise SystemExit("no kernel matching %r in %s" % (match, src))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
