#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configuration.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric   : accepted RK steps/s x state elements (fp64)
workload : configs[1] = Lorenz attractor, batch 65 536 x dim 3, fp64, dopri5 adaptive (rtol 1e-7, atol 1e-9),
           1 000 output points t = arange(1000) * 0.01, synthetic seeded initial states (cfg 2).
step     : ONE full odeint() solve of that workload (all accepted + rejected attempts, the dense output of
           all 1 000 points).  value = accepted_steps * state_elements / seconds, aggregated over ranks.
paths    : --path fused_rhs (default): func = tfdiffeq_b200.rhs.Lorenz, the library's own right-hand side, so the
           whole solve runs in one persistent kernel; --path external_func_cuda_graph / external_func_eager: func
           is an arbitrary external PyTorch callable (the general path).  The non-primary paths are measured too
           and reported under `other_paths`.
N > 1    : weak scaling -- every rank integrates its own 65 536-trajectory shard; the shards form ONE ODE
           system with a shared step size (reference semantics), the per-attempt error-norm exchange runs
           inside the finalize kernel over NVLink peer memory.

Extra objects on the JSON line: `roofline` (fused error/finalize kernel at this workload),
`roofline_headline` (the same kernel family at the north-star size 65 536 x 128 fp64, all buffers > L2),
`cpu_baseline` (the oracle port on the host cores), `e2e` (host buffers in, host buffers out through the
public odeint API), `clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "accepted_rk_steps_x_state_elements_per_s"
UNIT = "element-steps/s"
B, DIM, NPTS = 65536, 3, 1000
RTOL, ATOL = 1e-7, 1e-9
BYTES_PER_ELEM_STEP_FP64 = 352          # Dopri5 accepted step, 44 elements x 8 B
FINALIZE_ELEMS = 8                      # y0, y1, k1, k3..k7 read once by the fused finalize kernel


def lorenz_y0(batch, seed):
    rng = np.random.default_rng(seed)
    return np.array([1.0, 1.0, 1.0]) + 0.1 * rng.standard_normal((batch, 3))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe in B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on torch-CPU tensors, all host threads
# ------------------------------------------------------------------------------------------------------
def cpu_sample(npts, backend="numpy", batch=B):
    """A bounded sample of the same workload: the first `npts` output points of the full batch through the
    oracle (oracle/np_ref.py: op-for-op eager, the reference's execution model).  backend "torch" runs on
    torch-CPU tensors with every host thread (what TF-Eager would do); backend "numpy" is one thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_ref
    from problems import PROBLEMS
    y0 = lorenz_y0(batch, 0)
    if backend == "torch":
        torch.set_num_threads(os.cpu_count() or 1)
        f = PROBLEMS["lorenz"](backend="torch")
        y0 = torch.from_numpy(y0)
        threads = torch.get_num_threads()
    else:
        f = PROBLEMS["lorenz"](backend="numpy")
        threads = 1
    t = np.arange(npts) * 0.01
    st = np_ref.Stats()
    t0 = time.perf_counter()
    np_ref.odeint(f, y0, t, rtol=RTOL, atol=ATOL, method="dopri5", stats=st)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, n_acc=st.n_acc, n_elem=batch * DIM, threads=threads, backend=backend,
                value=st.n_acc * batch * DIM / dt,
                sample="first %d of %d output points of the full %dx%d batch, oracle port on %s (%d thread%s)" % (
                    npts, NPTS, batch, DIM, "torch-CPU eager" if backend == "torch" else "numpy", threads,
                    "" if threads == 1 else "s"))


def best_cpu_backend():
    """Eager multi-threaded dispatch of 1.5 MiB tensors can lose to one thread (thread wake-up per op): time
    both briefly and keep the faster one as THE cpu baseline, so the baseline is not artificially slow."""
    a = cpu_sample(12, "torch")
    b = cpu_sample(12, "numpy")
    return ("torch", a, b) if a["value"] >= b["value"] else ("numpy", a, b)


def run_reference(args, rank, world):
    if rank != 0:
        return
    backend, probe_t, probe_n = best_cpu_backend()
    vals, secs = [], []
    npts = 60 if backend == "numpy" else 30
    for i in range(args.warmup + args.steps):
        s = cpu_sample(npts, backend)
        if i >= args.warmup:
            vals.append(s["value"])
            secs.append(s["seconds"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "lorenz_b65536x3_f64_dopri5_1000pts", "sample": s["sample"],
                       "other_backend_probe": {"torch_all_threads": probe_t["value"], "numpy_1_thread": probe_n["value"]}},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": s["threads"], "kind": "port", "sample": s["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def tensor_core_func_block(dev, peaks):
    """the reference's ODENet func (dense_odenet.py:85-92, 64 -> 256 -> 256 -> 64, relu) on 131 072 rows
    inside dopri5 (rtol = atol = 1e-3, t in [0, 1]; fp32 state, TF32 tensor-core math).  One evaluation = one launch of
    k_mlp3_tf32; `roofline` is that kernel against the measured dense tensor throughput."""
    import tfdiffeq_b200 as tfd
    Bm, Dm, Hm = 131072, 64, 256
    torch.manual_seed(0)
    m = tfd.rhs.DenseMLP(Dm, Hm, "relu").to(dev)
    y0 = torch.randn(Bm, Dm, device=dev)
    t = torch.tensor([0., 1.])
    kw = dict(rtol=1e-3, atol=1e-3, method="dopri5")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def solve_ms(reps):
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tfd.odeint(m, y0, t, **kw)
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]
    for _ in range(2):
        tfd.odeint(m, y0, t, **kw)
    ms = solve_ms(5)
    st = dict(tfd.last_stats)
    m.tensor_cores = False
    torch.backends.cuda.matmul.allow_tf32 = False
    solve_ms(1)                                                   # cuBLAS handle / heuristics warm-up
    ms_fp32 = solve_ms(3)
    torch.backends.cuda.matmul.allow_tf32 = True
    solve_ms(1)
    ms_cublas_tf32 = solve_ms(3)
    torch.backends.cuda.matmul.allow_tf32 = False
    m.tensor_cores = True
    # the kernel alone, timed with events around each launch (inputs 32 MiB, L2 flushed)
    x = torch.randn(Bm, Dm, device=dev)
    ev = []
    with torch.no_grad():
        for _ in range(3):
            m(0.0, x)
        for _ in range(10):
            flush.fill_(2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m(0.0, x)
            b.record()
            torch.cuda.synchronize(dev)
            ev.append(a.elapsed_time(b))
    k_ms = sorted(ev)[len(ev) // 2]
    flops = 2.0 * Bm * (Dm * Hm + Hm * Hm + Hm * Dm)
    tf = flops / (k_ms * 1e-3) / 1e12
    # MEASURED_PEAKS.json holds dense bf16; TF32 runs at half the bf16 rate on this part
    bf16 = peaks.get("bf16_tflops")                              # burst figure: this kernel is timed alone
    peak_tf32 = (float(bf16) / 2.0) if bf16 else 1125.0
    return {"workload": "odenet_mlp_64x256x256x64_relu_b131072_f32_dopri5", "ms_per_solve": ms, "nfe": st.get("nfe"),
            "element_steps_per_s": st.get("n_accepted", 0) * Bm * Dm / (ms * 1e-3),
            "ms_per_solve_torch_fp32_matmul": ms_fp32, "ms_per_solve_cublas_tf32_layers": ms_cublas_tf32,
            "roofline": {"bound": "tensor", "kernel": "k_mlp3_tf32 (fc1-relu-fc2-relu-fc3 chained, one launch per evaluation)",
                         "achieved": tf, "peak": peak_tf32,
                         "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (TF32 = bf16 / 2 on sm_100)" if bf16 else "fallback: 2250 / 2 TFLOP/s nominal",
                         "unit": "TFLOP/s", "frac": tf / peak_tf32, "avg_launch_ms": k_ms,
                         "traffic": 35.2e6, "algorithmic_bytes_per_launch": int(2 * Bm * Dm * 4),
                         "note": "traffic = dram bytes of one ncu --set full capture (profiles/r01_mlp3_chained.md); the kernel is "
                                 "bound by each SM re-streaming the 384 KB of weights per 128-row tile from L2, not by the MMAs"}}


def headline_kernel_roofline(dev, peak):
    """The fused finalize kernel at the north-star size: 65 536 x 128 fp64 (64 MiB per buffer, 8 read streams
    = 512 MiB per launch, far beyond the 126 MB L2), timed with CUDA events around each launch."""
    import tfdiffeq_b200 as tfd
    from tfdiffeq_b200 import _lib
    import ctypes as C
    torch.manual_seed(0)
    n_b, n_d = 65536, 128
    y0 = torch.randn(n_b, n_d, dtype=torch.float64, device=dev)
    A = (-0.5 * torch.eye(n_d, dtype=torch.float64, device=dev)
         + 0.05 * torch.randn(n_d, n_d, dtype=torch.float64, device=dev))
    f = lambda t, y: y @ A                                          # noqa: E731
    t = torch.linspace(0., 2., 11, dtype=torch.float64)
    tfd.odeint(f, y0, t, method="dopri5", rtol=1e-6, atol=1e-9)    # warm-up
    _lib.check(_lib.lib.b2ode_timing_enable((1 << _lib.FAM_FINALIZE) | (1 << _lib.FAM_STAGE) | (1 << _lib.FAM_STAGE0)))
    tfd.odeint(f, y0, t, method="dopri5", rtol=1e-6, atol=1e-9)
    torch.cuda.synchronize(dev)
    out = {}
    n = n_b * n_d
    st = dict(tfd.last_stats)
    att = max(st["n_accepted"] + st["n_rejected"], 1)
    # stage rows 1..5 of Dopri5 read y0 + {2,3,4,5,5} k's and write one array: (4+5+6+7+7)/5 = 5.8 N per launch;
    # stage 0 moves 3 N (after a reject / first attempt) or 5 N (deferred commit after an accept)
    commit_frac = max(st["n_accepted"] - 1, 0) / float(att)
    for fam, name, elems in ((_lib.FAM_FINALIZE, "finalize", FINALIZE_ELEMS), (_lib.FAM_STAGE, "stages_1_to_5", 29.0 / 5),
                             (_lib.FAM_STAGE0, "stage_0", 3.0 + 2.0 * commit_frac)):
        ms, cnt = C.c_double(), C.c_int()
        _lib.check(_lib.lib.b2ode_timing_read(fam, C.byref(ms), C.byref(cnt)))
        if cnt.value:
            avg = ms.value / cnt.value
            gbs = elems * n * 8 / (avg * 1e-3) / 1e9
            out[name] = {"launches": cnt.value, "avg_ms": avg, "algorithmic_bytes": int(elems * n * 8),
                         "achieved": gbs, "frac": gbs / peak}
    _lib.check(_lib.lib.b2ode_timing_enable(0))
    fin = out.get("finalize", {})
    return {"bound": "hbm", "kernel": "k_rk_finalize<double,6>", "workload": "linear_b65536x128_f64_dopri5",
            "achieved": fin.get("achieved"), "peak": peak, "unit": "GB/s", "frac": fin.get("frac"),
            # ncu --set full, same launch (profiles/r01_finalize_65536x128_f64.md): 537.0 MB read + 3.3 MB written
            "traffic": 540.3e6, "per_kernel": out, "n_accepted": st.get("n_accepted"), "n_rejected": st.get("n_rejected")}


def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import tfdiffeq_b200 as tfd
    from tfdiffeq_b200 import _lib
    from problems import PROBLEMS
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        import torch.distributed as dist
        from tfdiffeq_b200.comm import SharedStepGroup
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        group = SharedStepGroup()
    peak, peak_src = peaks()
    # weak scaling: every rank owns a full 65 536-trajectory shard (different seed per rank)
    y0_host = torch.from_numpy(lorenz_y0(B, rank)).pin_memory()
    t_host = torch.arange(NPTS, dtype=torch.float64) * 0.01
    out_host = torch.empty((NPTS, B, DIM), dtype=torch.float64).pin_memory()
    y0_dev = y0_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    base_opts = {"shared_step_group": group} if group is not None else {}

    # the three ways the public API can run this workload
    paths = {
        # func = the library's own Lorenz module -> whole solve in one persistent kernel (b2ode_fused_solve)
        "fused_rhs": (tfd.rhs.Lorenz(), dict(base_opts)),
        # func = an arbitrary external PyTorch callable; one attempt captured in a CUDA graph and replayed
        "external_func_cuda_graph": (PROBLEMS["lorenz"](backend="torch", device=dev), dict(base_opts, cuda_graph=True)),
        # same, launched eagerly from python
        "external_func_eager": (PROBLEMS["lorenz"](backend="torch", device=dev), dict(base_opts)),
    }
    primary = args.path

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def measure(name, steps, warmup, e2e):
        f, opts = paths[name]
        kw = dict(rtol=RTOL, atol=ATOL, method="dopri5", options=opts)

        def solve():
            if e2e:
                y = y0_host.to(dev, non_blocking=True)            # H2D of the inputs inside the timed region
                sol = tfd.odeint(f, y, t_host, **kw)
                out_host.copy_(sol, non_blocking=True)            # D2H of the whole solution inside the timed region
                torch.cuda.synchronize(dev)
            else:
                tfd.odeint(f, y0_dev, t_host, **kw)
        for _ in range(warmup):
            solve()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        acc = rej = 0
        l0 = int(_lib.lib.b2ode_launch_count())
        barrier()
        w0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)                                 # L2 flush between timed iterations (untimed)
            ev[i][0].record()
            solve()
            ev[i][1].record()
            acc += tfd.last_stats["n_accepted"]
            rej += tfd.last_stats["n_rejected"]
        barrier()
        wall = time.perf_counter() - w0
        launches = int(_lib.lib.b2ode_launch_count()) - l0
        if name == "external_func_cuda_graph":
            # kernels inside the replayed graph are launched by the graph, not counted by the library's host-side
            # counter: each replayed attempt runs the same 8 library kernels (6 stages, finalize, dense output)
            launches += (acc + rej - steps) * 8
        ms = sum(a.elapsed_time(b) for a, b in ev)
        work = float(acc) * B * DIM
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ww = torch.tensor([work, float(launches)], dtype=torch.float64, device=dev)
            dist.all_reduce(ww, op=dist.ReduceOp.SUM)
            ms, work, launches = float(tt[0]), float(ww[0]), int(ww[1])
        return dict(value=work / (ms * 1e-3), ms_per_step=ms / steps, n_acc=acc / float(steps), n_rej=rej / float(steps),
                    launches=launches, wall=wall)

    W = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_res = measure(primary, args.steps, W, e2e=False)
    main_e2e = measure(primary, args.steps, 1, e2e=True)
    others = {}
    for name in paths:
        if name != primary:
            r = measure(name, max(1, min(args.steps, 2)), 1, e2e=False)
            others[name] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"]}
    clocks = sampler.stop() if rank == 0 else None     # sampled every 20 ms across all the timed regions above

    # ---- per-kernel timing passes (CUDA events recorded by the library around its own launches) ----------------
    n = B * DIM
    kern = {}
    if rank == 0 or world > 1:
        _lib.check(_lib.lib.b2ode_timing_enable((1 << _lib.FAM_FUSED) | (1 << _lib.FAM_FINALIZE)))
        for name in ("fused_rhs", "external_func_eager"):
            f, opts = paths[name]
            flush.fill_(3)
            tfd.odeint(f, y0_dev, t_host, rtol=RTOL, atol=ATOL, method="dopri5", options=opts)
            torch.cuda.synchronize(dev)
            kern[name] = dict(tfd.last_stats)
        fus_ms, fus_cnt, fin_ms, fin_cnt = C.c_double(), C.c_int(), C.c_double(), C.c_int()
        _lib.check(_lib.lib.b2ode_timing_read(_lib.FAM_FUSED, C.byref(fus_ms), C.byref(fus_cnt)))
        _lib.check(_lib.lib.b2ode_timing_read(_lib.FAM_FINALIZE, C.byref(fin_ms), C.byref(fin_cnt)))
        _lib.check(_lib.lib.b2ode_timing_enable(0))

    if rank == 0:
        headline = None
        cpu = None
        tc_func = None
        if world == 1:
            try:
                headline = headline_kernel_roofline(dev, peak)
            except Exception as e:                                 # noqa: BLE001  (never lose the main line)
                headline = {"error": repr(e)[:200]}
            try:
                try:
                    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                except Exception:                                  # noqa: BLE001
                    peaks_json = {}
                tc_func = tensor_core_func_block(dev, peaks_json)
            except Exception as e:                                 # noqa: BLE001
                tc_func = {"error": repr(e)[:200]}
            if not args.no_cpu_baseline:
                backend, pt, pn = best_cpu_backend()
                s = cpu_sample(60 if backend == "numpy" else 30, backend)
                cpu = {"value": s["value"], "unit": UNIT, "cores": s["threads"], "kind": "port", "sample": s["sample"],
                       "seconds": s["seconds"], "probe": {"torch_all_threads": pt["value"], "numpy_1_thread": pn["value"],
                                                          "host_cores": os.cpu_count()}}
        # roofline of the dominant kernel of the primary path
        fin_avg_ms = fin_ms.value / max(fin_cnt.value, 1)
        fin_bytes = FINALIZE_ELEMS * n * 8
        fin_gbs = fin_bytes / (fin_avg_ms * 1e-3) / 1e9 if fin_cnt.value else None
        finalize_roof = {"kernel": "k_rk_finalize<double,6> (error combine + norm + controller)", "achieved": fin_gbs,
                         "frac": (fin_gbs / peak) if fin_gbs else None, "algorithmic_bytes_per_launch": fin_bytes,
                         "avg_launch_ms": fin_avg_ms, "launches_timed": fin_cnt.value,
                         "note": "8 x 1.5 MiB read streams per launch: L2-resident and launch-latency bound at this size "
                                 "(eager pass; event pairs include the inter-launch gap); roofline_headline has the HBM-bound size"}
        if primary == "fused_rhs" and fus_cnt.value:
            ks = kern["fused_rhs"]
            fus_avg = fus_ms.value / fus_cnt.value
            # 352 B of HBM traffic per accepted fp64 element-step is what a func-external design must move;
            # the persistent kernel keeps state and k's in registers and only writes the solution slab
            alg = BYTES_PER_ELEM_STEP_FP64 * ks["n_accepted"] * n + 48 * NPTS * n
            slab = NPTS * n * 8 + n * 8
            roof = {"bound": "hbm", "kernel": "k_fused_adaptive<double, RhsLorenz<double>, 7> (one launch = one whole solve)",
                    "achieved": alg / (fus_avg * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": alg / (fus_avg * 1e-3) / 1e9 / peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture of the
                    # same launch (profiles/r01_fused_lorenz_65536x3_f64.md): the solution slab and nothing else
                    "traffic": 1517.4e6,
                    "algorithmic_bytes_per_launch": int(alg), "avg_launch_ms": fus_avg, "launches_timed": fus_cnt.value,
                    "bytes_actually_needed_per_launch": int(slab), "slab_write_GBps": slab / (fus_avg * 1e-3) / 1e9,
                    "note": "achieved uses SURVEY 8(d)'s per-unit bytes (the traffic of a design with func outside the kernel): "
                            "frac > 1 means the kernel avoids that traffic (state + k's in registers); its real HBM traffic is "
                            "the solution slab (`bytes_actually_needed_per_launch`), and its time is set by %d grid-wide "
                            "reductions (one per attempt), not by HBM" % (ks["n_accepted"] + ks["n_rejected"] + 2),
                    "finalize_kernel_generic_path": finalize_roof}
        else:
            roof = dict({"bound": "hbm", "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None}, **finalize_roof)
        func_desc = {"fused_rhs": "tfdiffeq_b200.rhs.Lorenz (library right-hand side: whole solve in one persistent kernel)",
                     "external_func_cuda_graph": "external PyTorch callable (tests/problems.py:Lorenz, 9 torch kernels per call), "
                                                 "options={'cuda_graph': True}",
                     "external_func_eager": "external PyTorch callable (tests/problems.py:Lorenz), eager launches"}[primary]
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "lorenz_b65536x3_f64_dopri5_1000pts", "per_gpu_batch": B, "dim": DIM, "rtol": RTOL,
                       "atol": ATOL, "n_out": NPTS, "path": primary, "func": func_desc,
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "parallelism": "batch shards, shared step via in-kernel NVLink mailbox exchange" if world > 1 else "single GPU",
                       "accepted_per_solve": main_res["n_acc"], "rejected_per_solve": main_res["n_rej"]},
            "roofline": roof,
            "roofline_headline": headline,
            "cpu_baseline": cpu,
            "e2e": {"value": main_e2e["value"], "unit": UNIT, "ms_per_step": main_e2e["ms_per_step"],
                    "h2d_bytes_per_step": int(B * DIM * 8 + NPTS * 8), "d2h_bytes_per_step": int(NPTS * B * DIM * 8)},
            "gpu_launches": main_res["launches"],
            "other_paths": others,
            "tensor_core_func": tc_func,
            "attempts_per_solve": main_res["n_acc"] + main_res["n_rej"],
            "wall_s_timed_region": main_res["wall"],
            "clocks": clocks,
        }
        print(json.dumps(line))
    if group is not None:
        group.close()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", default="fused_rhs", choices=["fused_rhs", "external_func_cuda_graph", "external_func_eager"],
                    help="which public-API path is the primary (timed) one; the others are reported under other_paths")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
