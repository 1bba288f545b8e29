#!/usr/bin/env python
"""Benchmark of the batched bounce solver (BASELINE.json metric: bounce actions / second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--mode auto|global|replicas]
                  [--impl ours|reference]

A "step" is one pass of the hot path (SingleFieldInstanton.__init__ + findProfile + findAction for every
instance) over one batch.  Workloads (BASELINE.json configs):
  c2  configs[1]: 1e5 quartic instances V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2, c_i = 0.02 + 0.475 (i + 1/2)/N,
      alpha = 2, all findProfile defaults.  The default at N = 1 (the configuration the metric is quoted on).
  c3  configs[2]: the same family, 1e6 instances, alpha = 3.  The default at N > 1.
  c5  configs[4] (one GPU's worth): 1000 parameter points x 256 temperatures of the one-field thermal model --
      phi_absMin(T) by the batched bounded minimiser, then one finite-T bounce (finite-difference dV, d2V) each.
  c4  configs[3]: model1 Vtot on the 4096 x 1024 (phi, T) grid (metric: grid points/s) + the 512-temperature bounce
      scan along the line between the phases.

N > 1 (torchrun, one rank per GPU), --mode global (default): ONE global batch, partitioned block-cyclically over
the ranks (cosmotransitions_b200.multigpu), every rank's results landing by DMA in a shared page-locked host
segment that rank 0 reads: strong scaling, no collective on the data path (torch.distributed: segment name,
barriers, max-over-ranks of the times).  --mode replicas: every rank owns one full batch (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    "c2": dict(n=100000, alpha=2, name="C2: 1e5 quartic thin->thick sweep, alpha=2, findProfile defaults",
               flop_per_bounce=7.3e4, ref_stride=16),
    "c3": dict(n=1000000, alpha=3, name="C3: 1e6 quartic thin->thick sweep, alpha=3, findProfile defaults",
               flop_per_bounce=9.0e4, ref_stride=160),
    "c5": dict(n=256000, alpha=2, name="C5 (one GPU's share): 1000 parameter points x 256 temperatures of the one-field "
                                       "thermal model, bounded minimiser + finite-T bounce per (point, T)",
               flop_per_bounce=1.8e6),
    "c5m": dict(n=10000 * 256, alpha=2, name="C5 on the two-field model (examples/testModel1.py): 10000 parameter points x "
                                             "256 temperatures, both phases followed on the device (ctb_trace_minima), "
                                             "then one bounce along the straight line between the minima wherever the "
                                             "high-T phase is metastable", flop_per_bounce=1.8e6),
    "c4": dict(n=4096 * 1024, alpha=2, name="C4: model1 Vtot on a 4096 x 1024 (phi, T) grid + 512-temperature bounce "
                                            "scan along the line between the phases"),
}
# slots (streams) of the end-to-end pipelines: with three batches in flight the thinning tails of one batch's kernels are
# filled by the heads of the next two (C2: 1.28e8 -> 1.42e8 bounces/s end to end; four slots: 1.35e8)
PIPE_DEPTH = 3
BYTES_PER_BOUNCE = 64.0  # 16-24 B in + ~48 B out
# dram__bytes_read.sum + dram__bytes_write.sum per STEP (every kernel of the step: setup + sort + shoot + save), from the
# `ncu --set full` captures under profiles/ (see profiles/README.md); None until captured for a workload
NCU_TRAFFIC_BYTES = {"c2": 25.7e6 + 1.3e6, "c3": 80.2e6 + 55.2e6 + 120.1e6 + 14.0e6 + 56e6}   # c3: the set-up kernel (not captured) scaled from c2
NCU_TRAFFIC_NOTE = ("whole step (setup + sort + shoot + save kernels), ncu dram__bytes_read+write summed over the "
                    "step's launches; algorithmic: 64 B x instances")


def sweep_c(n, rank=0, world=1):
    # replicas mode: every rank gets its own sweep over the same range (offset by a fraction of a grid step)
    return 0.02 + 0.475 * (np.arange(n) + (rank + 0.5) / world) / n


def quartic_params(c):
    p = np.zeros((c.size, 5))
    p[:, 2] = c / 2
    p[:, 3] = -(1 + c) / 3
    p[:, 4] = 0.25
    return p


def make_config(args, wl, world, mode):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm of the same command."""
    cfg = {"workload": wl["name"], "instances": wl["n"], "alpha": wl["alpha"], "order": args.order}
    if world > 1:
        cfg["partition"] = ("one global batch, block-cyclic over %d ranks" % world) if mode == "global" else \
                           ("%d independent replicas of the batch" % world)
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.windows = {}

    def mark(self, name, t0, t1):
        self.windows[name] = (t0, t1)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        # prefer samples taken inside the timed region; a region shorter than a few sampling periods
        # falls back to every sample taken while the same kernel loop kept the GPU busy
        window = "timed region"
        t0, t1 = self.windows.get("timed", (0, 0))
        sm, mx, reasons = summarise([x for x in self.rows if t0 <= x[0] <= t1])
        if len(sm) < 3:
            window = "under load (warm-up + timed region + end-to-end loop of the same step)"
            t0, t1 = self.windows.get("load", (0, float("inf")))
            sm, mx, reasons = summarise([x for x in self.rows if t0 <= x[0] <= t1])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref, pure Python, multiprocessing.Pool) and the C port of it

def cpu_port_rate(wl, threads, sample_stride, c=None):
    """The C restatement (oracle/ctb_oracle.c, pthreads) on a strided sample.  (bounces/s, results, idx)"""
    from oracle import oracle as O
    O.build()
    if c is None:
        c = sweep_c(wl["n"])
    idx = np.arange(0, c.size, sample_stride)
    cfg = O.make_config(alpha=wl["alpha"])
    p = quartic_params(c[idx])
    O.bounce_batch(cfg, p[:64], 1.0, 0.0, nthreads=1)  # warm the code path
    t0 = time.perf_counter()
    res = O.bounce_batch(cfg, p, 1.0, 0.0, nthreads=threads)
    dt = time.perf_counter() - t0
    return idx.size / dt, res, idx


def reference_subprocess(wl, stride, procs):
    """The unmodified Python reference under multiprocessing.Pool(procs), chunksize 25, in a
    subprocess (this process has CUDA initialised; the pool forks).  Returns the runner's dict or None."""
    runner = os.path.join(REPO, "oracle", "ref_runner.py")
    if not os.path.exists(os.path.join(REPO, "oracle", "_ref", "cosmoTransitions", "tunneling1D.py")):
        return None
    out = os.path.join(REPO, "gpurun_out", "ref_runner_%d.json" % os.getpid())
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [sys.executable, runner, "--n", str(wl["n"]), "--alpha", str(wl["alpha"]), "--stride", str(stride),
           "--procs", str(procs), "--out", out]
    try:
        subprocess.run(cmd, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(out) as f:
            return json.load(f)
    except Exception:
        return None
    finally:
        if os.path.exists(out):
            os.remove(out)


def run_reference(args, wl, world, mode):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same config / metric.
    oracle/_ref present (the unmodified Python reference, installed by oracle/make_ref.py): multiprocessing.Pool over
    a strided sample per step, analytic V / dV / d2V lambdas, chunksize 25 (kind "reference").  Otherwise the C port
    (kind "port").  Under torchrun rank 0 alone runs it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload not in ("c2", "c3"):
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm covers the quartic workloads c2 / c3"}))
        return
    from oracle import ref_runner
    threads = len(os.sched_getaffinity(0))
    n, alpha = wl["n"], wl["alpha"]
    total_steps = args.steps + args.warmup
    line_extra = {}
    if ref_runner.available() and not args.port:
        # ~27 bounces/s per core; whole run <= ~150 s
        per_step_s = args.ref_seconds if args.ref_seconds > 0 else min(6.0, 150.0 / total_steps)
        n_step = int(min(n // 16, max(4 * threads, 27.0 * threads * per_step_s)))
        stride = max(1, n // n_step)
        pool = ref_runner.make_pool(alpha, threads)
        try:
            S = None
            for k in range(args.warmup):
                idx = np.arange(k % stride, n, stride)
                ref_runner.quartic_pool(0.02 + 0.475 * (idx + 0.5) / n, alpha, threads, 25, pool=pool)
            t0 = time.perf_counter()
            count, solved = 0, 0
            for k in range(args.steps):
                idx = np.arange((args.warmup + k) % stride, n, stride)
                S, status, _ = ref_runner.quartic_pool(0.02 + 0.475 * (idx + 0.5) / n, alpha, threads, 25, pool=pool)
                count += idx.size
                solved += sum(s == "ok" for s in status)
            dt = time.perf_counter() - t0
        finally:
            pool.close()
            pool.join()
        kind = "reference"
        sample = ("every %d-th instance of the %d-instance sweep (%d per step, rotating offset), unmodified "
                  "cosmoTransitions.tunneling1D under multiprocessing.Pool(%d), chunksize 25" % (stride, n, count // args.steps, threads))
        value, solved_fraction = count / dt, solved / count
        import scipy
        line_extra["versions"] = {"numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0]}
    else:
        from oracle import oracle as O
        O.build()
        c = sweep_c(n)
        n_step = int(min(n, max(2000, 9000 * threads * 2)))     # ~2 s per step at ~9e3 bounces/s/thread
        stride = max(1, n // n_step)
        idx = np.arange(0, n, stride)
        cfg = O.make_config(alpha=alpha)
        p = quartic_params(c[idx])
        for _ in range(args.warmup):
            O.bounce_batch(cfg, p, 1.0, 0.0, nthreads=threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = O.bounce_batch(cfg, p, 1.0, 0.0, nthreads=threads)
        dt = time.perf_counter() - t0
        kind = "port"
        value, solved_fraction = idx.size * args.steps / dt, float((res["status"] <= 1).mean())
        sample = "every %d-th instance of the %d-instance sweep (%d per step), C port with pthreads" % (stride, n, idx.size)
    line = {
        "impl": "reference", "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and mode == "global") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": make_config(args, wl, world, mode),
        "cpu_baseline": {"value": value, "unit": "bounces/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "bounces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solved_fraction": solved_fraction,
    }
    line.update(line_extra)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------

def fp64_peak_tflops(_lib, torch, dev):
    """FP64 roofline denominator, measured here: register-resident DFMA loop (MEASURED_PEAKS.json has no FP64 entry)."""
    import ctypes as C
    lib = _lib.load()
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    ms = C.c_float(0)
    iters = 1 << 16
    blocks = max(1, lib.ctb_sm_count()) * 8
    best = 0.0
    for _ in range(3):
        _lib.check(lib.ctb_fp64_peak(iters, blocks, 256, sink.data_ptr(), C.byref(ms), None), "fp64_peak")
        best = max(best, blocks * 256 * iters * 16 * 2 / (ms.value * 1e-3) / 1e12)
    return best


def load_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except OSError:
        return {"hbm_gbs": 6650.0}, "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


class Env:
    pass


def setup_env(args):
    import torch
    import torch.distributed as dist
    from cosmotransitions_b200 import _lib
    e = Env()
    e.torch, e.dist, e._lib = torch, dist, _lib
    e.rank = int(os.environ.get("RANK", "0"))
    e.world = int(os.environ.get("WORLD_SIZE", "1"))
    e.local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(e.local)
    e.dev = torch.device("cuda", e.local)
    if e.world > 1:
        dist.init_process_group("nccl", device_id=e.dev)
    assert _lib.load().ctb_device_ok() == 0, "not an sm_100 device"
    e.flush = torch.empty(256 << 20, dtype=torch.uint8, device=e.dev)  # > 126 MB L2

    def barrier():
        if e.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=e.dev)
        if e.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e.barrier, e.max_over_ranks = barrier, max_over_ranks
    return e


def timed_device_loop(e, args, step, sampler, t_load0, settle_s=0.6):
    """W warm-up steps, then exactly K steps, each bracketed by CUDA events on the launching (current) stream, an L2
    flush before each; barrier + synchronize on both sides.  Returns (per-step ms list, wall seconds)."""
    torch = e.torch
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    while time.perf_counter() - t_load0 < settle_s:      # clocks settled, sampler reporting
        step()
        torch.cuda.synchronize()
    e.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e.flush.zero_()                 # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        step()
        ev[k][1].record()
    e.barrier()
    wall = time.perf_counter() - t0
    sampler.mark("timed", t0, t0 + wall)
    return [a.elapsed_time(b) for a, b in ev], wall


def quartic_roofline(e, args, wl, kern_avg_ms, n_in_step, scope):
    fp64_peak = fp64_peak_tflops(e._lib, e.torch, e.dev)
    achieved_tf = wl["flop_per_bounce"] * n_in_step / (kern_avg_ms * 1e-3) / 1e12
    peaks, peak_src = load_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_achieved = BYTES_PER_BOUNCE * n_in_step / (kern_avg_ms * 1e-3) / 1e9
    return {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if fp64_peak else None,
            "traffic": NCU_TRAFFIC_BYTES.get(args.workload), "traffic_note": NCU_TRAFFIC_NOTE,
            "peak_source": "ctb_fp64_peak register-resident DFMA loop, measured in this run "
                           "(MEASURED_PEAKS.json has no FP64 entry)",
            "algorithmic_flop_per_bounce": wl["flop_per_bounce"], "kernel_ms": kern_avg_ms, "scope": scope,
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                    "bytes_per_bounce": BYTES_PER_BOUNCE, "peak_source": peak_src}}


def cpu_baseline_and_parity(args, wl, c, out_np, want_reference=True):
    """Rank 0, after the timed regions: the reference (Python, Pool) on the strided sample of BASELINE.md s.4 and the C
    port on a larger one; S of the GPU run is checked against both on exactly those instances."""
    n = wl["n"]
    threads = len(os.sched_getaffinity(0))
    status, S_gpu = out_np["status"], out_np["S"]
    res = {}
    pstride = max(1, n // 100000)
    rate, pres, idx = cpu_port_rate(wl, threads, pstride, c)
    ok = (pres["status"] <= 1) & (status[idx] <= 1)
    rel = np.abs(S_gpu[idx][ok] - pres["S"][ok]) / np.abs(pres["S"][ok])
    port = {"value": rate, "unit": "bounces/s", "cores": threads, "kind": "port",
            "sample": "every %d-th instance of the workload (%d instances), one pass, C port (oracle/ctb_oracle.c) with "
                      "pthreads" % (pstride, idx.size)}
    parity = {"max_rel_err_S_vs_oracle": float(rel.max()) if rel.size else None,
              "status_equal": bool(np.array_equal(pres["status"], status[idx])), "checked": int(idx.size),
              "n_rkck_mismatches": int((out_np["n_rkck"][idx] != pres["n_rkck"]).sum()),
              "n_shots_mismatches": int((out_np["n_shots"][idx] != pres["n_shots"]).sum())}
    ref = reference_subprocess(wl, wl["ref_stride"], threads) if (want_reference and args.order == "natural") else None
    if ref is not None:
        ridx = np.arange(ref["offset"], n, ref["stride"])
        Sr = np.array(ref["S"])
        okr = np.array([s == "ok" for s in ref["status"]]) & (status[ridx] <= 1)
        relr = np.abs(S_gpu[ridx][okr] - Sr[okr]) / np.abs(Sr[okr])
        res["cpu_baseline"] = {
            "value": ref["rate"], "unit": "bounces/s", "cores": ref["procs"], "kind": "reference",
            "sample": "every %d-th instance of the workload (%d instances, %.1f s wall), the unmodified "
                      "cosmoTransitions.tunneling1D.SingleFieldInstanton (oracle/_ref) under multiprocessing.Pool(%d), "
                      "chunksize %d, analytic V / dV / d2V lambdas" % (ref["stride"], ref["count"], ref["wall_s"],
                                                                       ref["procs"], ref["chunksize"]),
            "versions": {"numpy": ref["numpy"], "scipy": ref["scipy"], "python": ref["python"]},
            "port": port}
        # knife-edge instances: where the reference's own S jumps by > 1e-6 when c moves in its 15th digit (one shot's
        # converged / overshoot classification flips), the GPU value must lie inside the band the reference itself spans
        outliers = []
        if relr.size and (relr >= 1e-6).any():
            from oracle import ref_runner
            for j in np.nonzero(okr)[0][relr >= 1e-6][:5]:
                ci = float(0.02 + 0.475 * (ridx[j] + 0.5) / n)
                lo, hi = ref_runner.quartic_band(ci, wl["alpha"])
                outliers.append({"c": ci, "S_gpu": float(S_gpu[ridx[j]]), "S_reference": float(Sr[j]),
                                 "reference_band_under_1e-13_perturbation_of_c": [lo, hi],
                                 "inside_band": bool(lo * (1 - 1e-6) <= S_gpu[ridx[j]] <= hi * (1 + 1e-6))})
        parity["live_reference"] = {"max_rel_err_S": float(relr.max()) if relr.size else None, "checked": int(okr.sum()),
                                    "above_1e-6": int((relr >= 1e-6).sum()), "knife_edge_outliers": outliers,
                                    "quantile_0.999_rel_err_S": float(np.quantile(relr, 0.999)) if relr.size else None,
                                    "reference_failures": int(len(ref["status"]) - sum(s == "ok" for s in ref["status"])),
                                    "gpu_failures_on_sample": int((status[ridx] > 1).sum())}
    else:
        res["cpu_baseline"] = port
    res["parity"] = parity
    return res


# ------------------------------------------------------------------------------------------------------------
# workloads c2 / c3 on one GPU, or as N replicas (weak)

def run_quartic_single(args, wl, e, mode):
    torch, _lib = e.torch, e._lib
    from cosmotransitions_b200 import batch
    n = wl["n"]
    sorted_sched = args.schedule == "sorted" or (args.schedule == "auto" and n >= batch.SORT_THRESHOLD)
    c = sweep_c(n, e.rank, e.world)
    if args.order == "reversed":
        c = c[::-1].copy()
    elif args.order == "shuffled":
        c = c[np.random.default_rng(1234).permutation(n)]
    params_h = torch.from_numpy(quartic_params(c)).pin_memory()
    abs_h = torch.ones(n, dtype=torch.float64).pin_memory()
    meta_h = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_params, d_abs, d_meta = params_h.to(e.dev), abs_h.to(e.dev), meta_h.to(e.dev)
    opts = batch.make_opts(alpha=wl["alpha"], block_threads=args.block)
    out = batch.alloc_out(torch, n, e.dev)

    def step():
        batch.launch_bounce(_lib.POT_POLY4, d_params, d_abs, d_meta, opts, out, schedule=args.schedule)

    sampler = ClockSampler(e.local)
    sampler.start()
    t_load0 = time.perf_counter()
    kern_ms, t_wall = timed_device_loop(e, args, step, sampler, t_load0)
    total_ms = e.max_over_ranks(sum(kern_ms))
    value = e.world * n * args.steps / (total_ms * 1e-3)
    out_np = {k: v.cpu().numpy() for k, v in out.items()}

    # ---- the same K device-resident steps with PIPE_DEPTH steps in flight (one stream each): what the GPU delivers when
    # the thinning tail of one batch's kernels is filled by the next batches (reported beside `value`, which stays the
    # per-step figure).  Inputs rotate over enough distinct device copies to exceed the L2 (no flush kernels between
    # overlapping steps); device-timed between one start and one end event, everything joined before the end event.
    overlapped = None
    if e.world == 1 and args.order == "natural":
        ncopy = max(PIPE_DEPTH + 1, int(130e6 // (n * 56)) + 1)          # 56 B of input per instance; > 126 MB in total
        copies = [(d_params.clone(), d_abs.clone(), d_meta.clone()) for _ in range(ncopy)]
        streams = [torch.cuda.Stream(device=e.dev) for _ in range(PIPE_DEPTH)]
        outs = [batch.alloc_out(torch, n, e.dev) for _ in range(PIPE_DEPTH)]

        def run_overlapped(steps):
            main = torch.cuda.current_stream(e.dev)
            for s_ in streams:
                s_.wait_stream(main)
            for k in range(steps):
                cp = copies[k % ncopy]
                with torch.cuda.stream(streams[k % PIPE_DEPTH]):
                    batch.launch_bounce(_lib.POT_POLY4, cp[0], cp[1], cp[2], opts, outs[k % PIPE_DEPTH], schedule=args.schedule)
            for s_ in streams:
                main.wait_stream(s_)

        run_overlapped(PIPE_DEPTH + 1)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        run_overlapped(args.steps)
        b_.record()
        torch.cuda.synchronize()
        ovl_ms = a_.elapsed_time(b_)
        same = bool(torch.equal(outs[(args.steps - 1) % PIPE_DEPTH]["S"], out["S"]))
        overlapped = {"value": n * args.steps / (ovl_ms * 1e-3), "unit": "bounces/s", "steps_in_flight": PIPE_DEPTH,
                      "ms_per_step": ovl_ms / args.steps, "input_copies": ncopy, "bit_identical_to_serial": same,
                      "how": "K device-resident steps on %d streams round-robin, inputs rotating over %d distinct device "
                             "copies (%.0f MB > L2), one CUDA event pair around all K steps" % (PIPE_DEPTH, ncopy, ncopy * n * 56 / 1e6)}
        del copies, outs

    # ---- end to end through the public API: host buffers in, host results out, every step ----
    # BouncePipeline: PIPE_DEPTH slots, each with its own stream -- the H2D copy of step k+1 and the D2H copy of step k-1
    # overlap the kernels of step k; every step's copies happen inside the timed region
    pipe = batch.BouncePipeline(_lib.POT_POLY4, n, depth=PIPE_DEPTH, device=e.dev, outputs=("S", "status"), alpha=wl["alpha"],
                                block_threads=args.block, schedule=args.schedule)

    def e2e_run(steps):
        checksum = 0.0
        for k in range(steps):
            if k >= pipe.depth:
                res = pipe.collect(pipe._submitted - pipe.depth)
                checksum += float(res["S"][0]) + int(res["status"][-1])     # the host reads every step's result
            pipe.submit(params_h, abs_h, meta_h)
        for t_ in range(max(pipe._submitted - pipe.depth, pipe._submitted - steps), pipe._submitted):
            res = pipe.collect(t_)
            checksum += float(res["S"][0]) + int(res["status"][-1])
        return checksum

    e2e_run(3)
    e.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e1.record()
    e.barrier()
    e2e_wall = time.perf_counter() - t0
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    e2e_ms = e.max_over_ranks(max(e0.elapsed_time(e1), e2e_wall * 1e3))
    e2e_value = e.world * n * args.steps / (e2e_ms * 1e-3)
    h2d = params_h.numel() * 8 + abs_h.numel() * 8 + meta_h.numel() * 8
    d2h = n * 8 + n * 4
    if e.rank != 0:
        return None
    kern_avg_ms = float(np.mean(kern_ms))
    line = {
        "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": e.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(args, wl, e.world, mode),
        "schedule": "work-sorted (setup kernel + device argsort + shoot kernel + save kernel, all inside the timed "
                    "step)" if sorted_sched else "natural",
        "timing": "CUDA events around each step's kernels on torch's current stream (the launching stream); 256 MiB L2 "
                  "flush between timed iterations; max over ranks",
        "wall_s_timed_region": t_wall,
        "solved_fraction": float((out_np["status"] <= 1).mean()),
        "roofline": quartic_roofline(e, args, wl, kern_avg_ms, n,
                                     "whole step: setup + sort + shoot + save kernels (the flop count per bounce covers "
                                     "shooting, dense output and quadrature)"),
        "e2e": {"value": e2e_value, "unit": "bounces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "batch.BouncePipeline (public API), pinned host buffers in / pinned host results out every step, %d "
                       "slots on as many CUDA streams so the copies and kernel tails of neighbouring steps overlap; wall "
                       "clock around the whole loop including the final synchronisation" % PIPE_DEPTH},
        # per step: setup_kernel + shoot kernel + save kernel (work-sorted, split solve); fused otherwise
        "gpu_launches": args.steps * (3 if sorted_sched else 1),
        "clocks": clocks,
    }
    if overlapped is not None:
        line["device_resident_overlapped"] = overlapped
    if getattr(args, "default_workload", False) and e.world == 1 and args.workload == "c2":
        # The same command at N > 1 partitions the C3 batch (BASELINE.json configs[2]: "at 1/2/4/8 B200"): its one-GPU
        # figure, measured here the same way as `value`, is the denominator a strong-scaling efficiency needs
        w3 = WORKLOADS["c3"]
        c3 = sweep_c(w3["n"])
        p3 = torch.from_numpy(quartic_params(c3)).to(e.dev)
        a3, m3 = torch.ones(w3["n"], dtype=torch.float64, device=e.dev), torch.zeros(w3["n"], dtype=torch.float64, device=e.dev)
        o3 = batch.alloc_out(torch, w3["n"], e.dev)
        opts3 = batch.make_opts(alpha=w3["alpha"], block_threads=args.block)
        ms3 = []
        for k in range(2 + 5):
            e.flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            batch.launch_bounce(_lib.POT_POLY4, p3, a3, m3, opts3, o3, schedule=args.schedule)
            b_.record()
            torch.cuda.synchronize()
            if k >= 2:
                ms3.append(a_.elapsed_time(b_))
        line["multi_gpu_workload_on_one_gpu"] = {
            "workload": w3["name"], "value": w3["n"] / (float(np.mean(ms3)) * 1e-3), "unit": "bounces/s", "steps": len(ms3),
            "ms_per_step": float(np.mean(ms3)), "solved_fraction": float((o3["status"] <= 1).float().mean()),
            "note": "`bench.py --gpus N` (N > 1) partitions THIS batch over the ranks; divide its `value` by N times this "
                    "figure for the strong-scaling efficiency (the N > 1 lines repeat the measurement as "
                    "single_gpu_same_workload)"}
        del p3, a3, m3, o3
    if not args.no_cpu_baseline:
        line.update(cpu_baseline_and_parity(args, wl, c, out_np))
    return line


# ------------------------------------------------------------------------------------------------------------
# workloads c2 / c3 as ONE global batch partitioned over the ranks (strong scaling)

def run_quartic_global(args, wl, e, mode):
    torch, _lib = e.torch, e._lib
    from cosmotransitions_b200 import batch, multigpu
    n = wl["n"]
    c = sweep_c(n)
    if args.order == "reversed":
        c = c[::-1].copy()
    elif args.order == "shuffled":
        c = c[np.random.default_rng(1234).permutation(n)]
    # every rank holds the same global input arrays in page-locked host memory (in production: one shared input)
    params_h = torch.from_numpy(quartic_params(c)).pin_memory()
    abs_h = torch.ones(n, dtype=torch.float64).pin_memory()
    meta_h = torch.zeros(n, dtype=torch.float64).pin_memory()
    outs = ("S", "status")
    db = multigpu.DistributedBounce(_lib.POT_POLY4, n, block=args.partition_block, depth=PIPE_DEPTH, outputs=outs, device=e.dev,
                                    alpha=wl["alpha"], block_threads=args.block, schedule=args.schedule)
    lay = db.layout
    n_local = lay.n_local
    sorted_sched = args.schedule == "sorted" or (args.schedule == "auto" and n_local >= batch.SORT_THRESHOLD)

    # ---- `value`: this rank's shard resident in HBM, kernels only, CUDA events; max over ranks ----
    t0 = db.submit(params_h, abs_h, meta_h)
    first = db.collect(t0)
    S_first = first["S"].copy() if e.rank == 0 else None
    st_first = first["status"].copy() if e.rank == 0 else None
    slot = db._sharded._slots[0]          # the shard's device-resident inputs (filled by the submit above)
    opts = db._sharded.opts
    out = batch.alloc_out(torch, n_local, e.dev)

    def step():
        batch.launch_bounce(_lib.POT_POLY4, slot["params"], slot["phi_abs"], slot["phi_meta"], opts, out,
                            schedule=args.schedule)

    sampler = ClockSampler(e.local)
    sampler.start()
    t_load0 = time.perf_counter()
    kern_ms, t_wall = timed_device_loop(e, args, step, sampler, t_load0)
    total_ms = e.max_over_ranks(sum(kern_ms))
    value = n * args.steps / (total_ms * 1e-3)

    # ---- `e2e`: the public multi-GPU API, global host arrays in, gathered global host result on rank 0, every step ----
    def e2e_run(steps):
        checksum = 0.0
        tickets = []
        for k in range(steps):
            if len(tickets) >= db.depth:
                res = db.collect(tickets.pop(0))
                if res is not None:
                    checksum += float(res["S"][0]) + float(res["S"][-1]) + int(res["status"][n // 2])
            tickets.append(db.submit(params_h, abs_h, meta_h))
        for t_ in tickets:
            res = db.collect(t_)
            if res is not None:
                checksum += float(res["S"][0]) + float(res["S"][-1]) + int(res["status"][n // 2])
        return checksum

    e2e_run(3)
    e.barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_wall_local = time.perf_counter() - t0
    e.barrier()
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    e2e_ms = e.max_over_ranks(e2e_wall_local * 1e3)
    e2e_value = n * args.steps / (e2e_ms * 1e-3)
    h2d = n_local * (5 * 8 + 8 + 8)  # the shard's rows of params, phi_absMin, phi_metaMin
    d2h = n_local * 12

    # ---- the same global batch on ONE GPU (rank 0), same run: the denominator of the strong-scaling efficiency ----
    single = None
    if e.rank == 0 and not args.no_single:
        d_params, d_abs, d_meta = params_h.to(e.dev), abs_h.to(e.dev), meta_h.to(e.dev)
        out1 = batch.alloc_out(torch, n, e.dev)
        opts1 = batch.make_opts(alpha=wl["alpha"], block_threads=args.block)
        ks = max(3, min(args.steps, 10))
        for _ in range(2):
            batch.launch_bounce(_lib.POT_POLY4, d_params, d_abs, d_meta, opts1, out1, schedule=args.schedule)
        torch.cuda.synchronize()
        ms1 = []
        for _ in range(ks):
            e.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            batch.launch_bounce(_lib.POT_POLY4, d_params, d_abs, d_meta, opts1, out1, schedule=args.schedule)
            b.record()
            torch.cuda.synchronize()
            ms1.append(a.elapsed_time(b))
        sb = multigpu.ShardedBounce(_lib.POT_POLY4, n, 1, 0, device=e.dev, depth=PIPE_DEPTH, outputs=outs, alpha=wl["alpha"],
                                    block_threads=args.block, schedule=args.schedule)
        res1 = {"S": torch.empty(n, dtype=torch.float64).pin_memory(), "status": torch.empty(n, dtype=torch.int32).pin_memory()}
        for _ in range(2):
            sb.submit(params_h, 1.0, 0.0, res1)
        sb.wait()
        tw = time.perf_counter()
        tk = []
        for k in range(ks):
            if len(tk) >= 2:
                sb.wait(tk.pop(0))
            tk.append(sb.submit(params_h, 1.0, 0.0, res1))
        for t_ in tk:
            sb.wait(t_)
        e2e1 = n * ks / (time.perf_counter() - tw)
        v1 = n / (float(np.mean(ms1)) * 1e-3)
        single = {"value": v1, "e2e": e2e1, "unit": "bounces/s", "steps": ks,
                  "strong_scaling_efficiency": {"value": value / (e.world * v1), "e2e": e2e_value / (e.world * e2e1)},
                  "bit_identical_to_partitioned": bool(np.array_equal(out1["S"].cpu().numpy(), S_first, equal_nan=True)
                                                       and np.array_equal(out1["status"].cpu().numpy(), st_first)),
                  "note": "the same 1-GPU code path as --gpus 1 --workload %s, run by rank 0 after the timed regions" % args.workload}
        out_np = {k: v.cpu().numpy() for k, v in out1.items()}
    e.barrier()
    db.close()
    if e.rank != 0:
        return None
    kern_avg_ms = total_ms / args.steps
    line = {
        "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": e.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(args, wl, e.world, mode),
        "partition": {"block": lay.block, "instances_per_rank": n_local, "how": "block-cyclic (multigpu.ShardLayout); "
                      "results gathered by each GPU's DMA engine into one shared page-locked host segment, no collective"},
        "schedule": "work-sorted (setup kernel + device argsort + shoot kernel + save kernel, all inside the timed "
                    "step)" if sorted_sched else "natural",
        "timing": "value: CUDA events around each step's kernels on the launching stream, shard resident in HBM, L2 flush "
                  "between iterations, MAX over ranks of the summed times; e2e: wall clock between barriers, max over ranks",
        "wall_s_timed_region": t_wall,
        "solved_fraction": float((st_first <= 1).mean()),
        "roofline": quartic_roofline(e, args, wl, kern_avg_ms, n_local,
                                     "one rank's shard per step (setup + sort + shoot + save kernels), slowest rank"),
        "e2e": {"value": e2e_value, "unit": "bounces/s", "h2d_bytes_per_step": h2d * e.world, "d2h_bytes_per_step": d2h * e.world,
                "how": "multigpu.DistributedBounce (public API): every rank DMAs its block-cyclic shard out of the pinned "
                       "GLOBAL input array, solves it, and DMAs S / status into the shared pinned GLOBAL result segment; "
                       "rank 0 reads every step's gathered result; three result slots rotate so consecutive steps overlap; "
                       "wall clock over all steps, max over ranks"},
        "gpu_launches": args.steps * (3 if sorted_sched else 1) * e.world,
        "clocks": clocks,
    }
    if single is not None:
        line["single_gpu_same_workload"] = single
        if not args.no_cpu_baseline:
            line.update(cpu_baseline_and_parity(args, wl, c, out_np, want_reference=False))
    return line


# ------------------------------------------------------------------------------------------------------------
# workload c5: finite-T bounces of the one-field thermal model (minimiser + bounce per (point, T))

def run_c5(args, wl, e, mode):
    torch, _lib = e.torch, e._lib
    from cosmotransitions_b200 import batch, scan
    P, nT = 1000, 256
    rng = np.random.default_rng(20260101)
    lam, Y, nb = 0.12 * rng.uniform(.8, 1.2, 10000), 0.35 * rng.uniform(.8, 1.2, 10000), 30. * rng.uniform(.8, 1.2, 10000)
    sel = np.arange(e.rank, 10000, max(1, 10000 // P))[:P] if e.world == 1 else np.arange(e.rank, 10000, e.world)[:P]
    lam, Y, nb = (torch.as_tensor(x[sel], dtype=torch.float64, device=e.dev) for x in (lam, Y, nb))
    T0, Tc = scan.model1f_temperature_window(lam, Y, nb)       # input preparation (not in the timed step)
    frac = (torch.arange(nT, dtype=torch.float64, device=e.dev) + 0.5) / nT
    T = T0[:, None] + frac[None, :] * (Tc - T0)[:, None]
    params = scan.model1f_params(lam[:, None], Y[:, None], nb[:, None], T).reshape(-1, 6).contiguous()
    n = params.shape[0]
    lo = torch.full((n,), 50.0, dtype=torch.float64, device=e.dev)
    hi = torch.full((n,), 500.0, dtype=torch.float64, device=e.dev)
    zero = torch.zeros(n, dtype=torch.float64, device=e.dev)
    opts = batch.make_opts(alpha=2, block_threads=args.block)
    out = batch.alloc_out(torch, n, e.dev)

    def step():
        phi_abs, _ = scan.minimize_1d(_lib.POT_MODEL1F, params, lo, hi)
        batch.launch_bounce(_lib.POT_MODEL1F, params, phi_abs, zero, opts, out, schedule=args.schedule)

    sampler = ClockSampler(e.local)
    sampler.start()
    t_load0 = time.perf_counter()
    kern_ms, t_wall = timed_device_loop(e, args, step, sampler, t_load0)
    total_ms = e.max_over_ranks(sum(kern_ms))
    value = e.world * n * args.steps / (total_ms * 1e-3)
    status = out["status"].cpu().numpy()
    # e2e: pinned host parameter rows in, S / status out, through the public functions
    params_h = params.cpu().pin_memory()
    S_h, st_h = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()

    def e2e_step():
        p = params_h.to(e.dev, non_blocking=True)
        phi_abs, _ = scan.minimize_1d(_lib.POT_MODEL1F, p, lo, hi)
        r = batch.find_bounce_batch(_lib.POT_MODEL1F, p, phi_abs, zero)
        S_h.copy_(r["S"], non_blocking=True)
        st_h.copy_(r["status"], non_blocking=True)
        torch.cuda.synchronize()
        return float(S_h[0]) + int(st_h[-1])

    e2e_step()
    e.barrier()
    t0 = time.perf_counter()
    ks = max(1, min(args.steps, 5))
    for _ in range(ks):
        e2e_step()
    e.barrier()
    e2e_ms = e.max_over_ranks((time.perf_counter() - t0) * 1e3)
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    if e.rank != 0:
        return None
    kern_avg_ms = float(np.mean(kern_ms))
    fp64_peak = fp64_peak_tflops(_lib, torch, e.dev)
    achieved = wl["flop_per_bounce"] * n / (kern_avg_ms * 1e-3) / 1e12
    return {
        "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": e.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(args, wl, e.world, mode),
        "timing": "CUDA events around each step (bounded minimiser + setup + sort + shoot + save kernels) on the launching "
                  "stream; 256 MiB L2 flush between iterations",
        "wall_s_timed_region": t_wall, "solved_fraction": float((status <= 1).mean()),
        "status_hist": np.bincount(status, minlength=12).tolist(),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "traffic": None, "algorithmic_flop_per_bounce": wl["flop_per_bounce"], "kernel_ms": kern_avg_ms,
                     "peak_source": "ctb_fp64_peak register-resident DFMA loop, measured in this run",
                     "scope": "whole step; SURVEY.md s.8d: ~6e3 Vtot evaluations x 300 flop per finite-T bounce"},
        "e2e": {"value": e.world * n * ks / (e2e_ms * 1e-3), "unit": "bounces/s", "h2d_bytes_per_step": n * 48,
                "d2h_bytes_per_step": n * 12, "steps": ks,
                "how": "pinned host parameter rows -> scan.minimize_1d -> batch.find_bounce_batch -> pinned host S / status"},
        "gpu_launches": args.steps * 4, "clocks": clocks,
    }


# ------------------------------------------------------------------------------------------------------------
# workload c5m: BASELINE.json configs[4] on the real two-field model1 -- phase following + straight-line bounces

def run_c5m(args, wl, e, mode):
    torch, _lib = e.torch, e._lib
    from cosmotransitions_b200 import phases, potentials
    P_all, nT = 10000, 256
    rng = np.random.default_rng(20260102)
    jit = rng.uniform(.9, 1.1, (P_all, 5))
    rows = np.stack([potentials.model1_model8(120. * a, 50. * b, 25. * c, .1 * d, .15 * f) for a, b, c, d, f in jit])
    sel = np.arange(e.rank, P_all, e.world)
    m8 = torch.as_tensor(rows[sel], dtype=torch.float64, device=e.dev).contiguous()
    P = m8.shape[0]
    T = torch.linspace(40.0, 135.0, nT, dtype=torch.float64, device=e.dev)
    res = {}

    def step():
        res["r"] = phases.model1_phase_scan(m8, T)

    sampler = ClockSampler(e.local)
    sampler.start()
    t_load0 = time.perf_counter()
    kern_ms, t_wall = timed_device_loop(e, args, step, sampler, t_load0)
    total_ms = e.max_over_ranks(sum(kern_ms))
    r = res["r"]
    nb = r["n_bounces"]
    st = r["status"].cpu().numpy()
    solved = int(((st == 0) | (st == 1)).sum())
    # e2e: pinned host parameter rows in, S / status / Tnuc out
    m8_h = m8.cpu().pin_memory()
    S_h = torch.empty((P, nT), dtype=torch.float64).pin_memory()
    Tn_h = torch.empty(P, dtype=torch.float64).pin_memory()

    def e2e_step():
        q = phases.model1_phase_scan(m8_h.to(e.dev, non_blocking=True), T)
        S_h.copy_(q["S"], non_blocking=True)
        Tn_h.copy_(q["Tnuc"], non_blocking=True)
        torch.cuda.synchronize()
        return float(Tn_h[0])

    e2e_step()
    e.barrier()
    t0 = time.perf_counter()
    ks = max(1, min(args.steps, 5))
    for _ in range(ks):
        e2e_step()
    e.barrier()
    e2e_ms = e.max_over_ranks((time.perf_counter() - t0) * 1e3)
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    if e.rank != 0:
        return None
    kern_avg_ms = float(np.mean(kern_ms))
    fp64_peak = fp64_peak_tflops(_lib, torch, e.dev)
    achieved = wl["flop_per_bounce"] * nb / (kern_avg_ms * 1e-3) / 1e12
    Tn = r["Tnuc"].cpu().numpy()
    return {
        "metric": "bounce actions/sec", "value": e.world * nb * args.steps / (total_ms * 1e-3), "unit": "bounces/s",
        "n_gpus": e.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, wl, e.world, mode),
        "timing": "CUDA events around each step (two ctb_trace_minima launches, the torch bookkeeping between them, setup + "
                  "sort + shoot + save kernels of the bounce batch) on the launching stream; 256 MiB L2 flush between iterations",
        "wall_s_timed_region": t_wall,
        "scan": {"parameter_points_this_rank": P, "temperatures": nT, "point_T_pairs_per_s": e.world * P * nT * args.steps / (total_ms * 1e-3),
                 "bounces_per_step_this_rank": nb, "solved": solved, "newton_iterations": r["newton_iters"],
                 "points_with_Tnuc": int(np.isfinite(Tn).sum()), "Tnuc_median": float(np.nanmedian(Tn)) if np.isfinite(Tn).any() else None,
                 "Tcrit_median": float(np.nanmedian(r["Tcrit"].cpu().numpy()))},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "traffic": None, "algorithmic_flop_per_bounce": wl["flop_per_bounce"], "kernel_ms": kern_avg_ms,
                     "peak_source": "ctb_fp64_peak register-resident DFMA loop, measured in this run",
                     "scope": "whole step, bounces only (the phase following is extra work on top): ~6e3 Vtot evaluations x 300 "
                              "flop per finite-T bounce (SURVEY.md s.8d)"},
        "e2e": {"value": e.world * nb * ks / (e2e_ms * 1e-3), "unit": "bounces/s", "h2d_bytes_per_step": P * 64,
                "d2h_bytes_per_step": P * nT * 8 + P * 8, "steps": ks,
                "how": "pinned host model rows -> phases.model1_phase_scan (public API) -> pinned host S[P, nT] and Tnuc[P]"},
        "gpu_launches": args.steps * 6, "clocks": clocks,
    }


# ------------------------------------------------------------------------------------------------------------
# workload c4: model1 Vtot on the 4096 x 1024 grid (+ the line bounce scan as a second figure)

def run_c4(args, wl, e, mode):
    torch, _lib = e.torch, e._lib
    from cosmotransitions_b200 import generic_potential as gp, scan
    m = gp.model1()
    x_low = np.array([286.39824989, 382.19727238])
    x_high = np.array([231.10296383, -136.82260069])
    L = np.linalg.norm(x_high - x_low)
    nhat = (x_high - x_low) / L
    s = torch.linspace(-0.2 * L, 1.2 * L, 4096, dtype=torch.float64, device=e.dev)
    T = torch.linspace(1, 250, 1024, dtype=torch.float64, device=e.dev)
    buf = torch.empty((4096, 1024), dtype=torch.float64, device=e.dev)

    def step():
        m.Vtot_grid(x_low, nhat, s, T, out=buf)

    sampler = ClockSampler(e.local)
    sampler.start()
    t_load0 = time.perf_counter()
    kern_ms, t_wall = timed_device_loop(e, args, step, sampler, t_load0)
    total_ms = e.max_over_ranks(sum(kern_ms))
    npts = 4096 * 1024
    value = e.world * npts * args.steps / (total_ms * 1e-3)
    # e2e: host axes in, host grid out
    s_h, T_h = s.cpu().numpy(), T.cpu().numpy()
    pinned = torch.empty((4096, 1024), dtype=torch.float64).pin_memory()
    m.Vtot_grid(x_low, nhat, s_h, T_h, host_out=pinned)
    t0 = time.perf_counter()
    ks = max(1, min(args.steps, 5))
    for _ in range(ks):
        g = m.Vtot_grid(x_low, nhat, s_h, T_h, host_out=pinned)
    e2e_s = (time.perf_counter() - t0) / ks
    assert float(g[7, 11]) == float(buf[7, 11])
    Ts = np.linspace(77.8, 109.2, 512)
    scan.model1_line_bounce_scan(m.model8, x_low, x_high, Ts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = scan.model1_line_bounce_scan(m.model8, x_low, x_high, Ts)
    torch.cuda.synchronize()
    t_scan = time.perf_counter() - t0
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    if e.rank != 0:
        return None
    peaks, peak_src = load_peaks()
    kern_avg_ms = float(np.mean(kern_ms))
    hbm = 8.0 * npts / (kern_avg_ms * 1e-3) / 1e9
    fp64_peak = fp64_peak_tflops(_lib, torch, e.dev)
    flop_pt = 40.0   # executed per point by the tracked kernel: 3 x (subtract, 2 compares, 3 FMA) + 3 multiplies + the final FMA
    return {
        "metric": "Vtot grid points/sec", "value": value, "unit": "points/s", "n_gpus": e.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(args, wl, e.world, mode),
        "timing": "CUDA events around each ctb_vtot_model1_grid launch; 256 MiB L2 flush between iterations",
        "wall_s_timed_region": t_wall,
        "roofline": {"bound": "hbm", "achieved": hbm, "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                     "frac": hbm / peaks.get("hbm_gbs", 6650.0), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_point": 8.0, "kernel_ms": kern_avg_ms,
                     "fp64": {"achieved": flop_pt * npts / (kern_avg_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "flop_per_point_executed": flop_pt,
                              "note": "per point: three tracked spline look-ups (interval test + cubic) and the final FMA; "
                                      "1/T^2 and T^4 are per THREAD, the three logs + sqrt of the field-dependent part "
                                      "per ROW; the kernel is latency / issue-bound table work (DESIGN.md s.4.6)"}},
        "e2e": {"value": npts / e2e_s, "unit": "points/s", "h2d_bytes_per_step": (4096 + 1024) * 8, "d2h_bytes_per_step": npts * 8,
                "how": "generic_potential.model1.Vtot_grid with numpy axes in, numpy grid out through a page-locked host buffer"},
        "line_bounce_scan": {"temperatures": len(Ts), "seconds": t_scan, "bounces_per_s": len(Ts) / t_scan,
                             "solved": int((r["status"].cpu().numpy() <= 1).sum()), "Tnuc_line": float(r["Tnuc"])},
        "gpu_launches": args.steps, "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="auto", choices=["auto", "global", "replicas"],
                    help="N > 1: one global batch partitioned over the ranks (default) or one full batch per rank")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--port", action="store_true", help="--impl reference: time the C port even if oracle/_ref exists")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: wall-time budget per step for the Python reference (default: <= 6 s, whole run <= ~150 s)")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--partition-block", type=int, default=256, help="instances per block of the block-cyclic partition")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="global mode: skip the 1-GPU run of the same batch on rank 0")
    ap.add_argument("--schedule", default="auto", choices=["auto", "sorted", "natural"],
                    help="instance scheduling inside the library (see batch.launch_bounce)")
    ap.add_argument("--order", default="natural", choices=["natural", "reversed", "shuffled"],
                    help="instance order inside the batch (SURVEY.md s.8d: shuffled exposes divergence sensitivity)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.default_workload = args.workload is None
    if args.workload is None:
        args.workload = "c2" if world == 1 else "c3"
    wl = WORKLOADS[args.workload]
    mode = args.mode
    if mode == "auto":
        mode = "global" if (world > 1 and args.workload in ("c2", "c3")) else "replicas"
    if world == 1:
        mode = "single"
    if args.impl == "reference":
        run_reference(args, wl, world, mode)
        return
    e = setup_env(args)
    if args.workload == "c5":
        line = run_c5(args, wl, e, mode)
    elif args.workload == "c5m":
        line = run_c5m(args, wl, e, mode)
    elif args.workload == "c4":
        line = run_c4(args, wl, e, mode)
    elif mode == "global":
        line = run_quartic_global(args, wl, e, mode)
    else:
        line = run_quartic_single(args, wl, e, mode)
    if e.rank == 0:
        print(json.dumps(line), flush=True)
    if e.world > 1:
        e.dist.destroy_process_group()


if __name__ == "__main__":
    main()
