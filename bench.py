#!/usr/bin/env python
"""Benchmark of the batched bounce solver (BASELINE.json metric: bounce actions / second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3] [--impl ours|reference]

A "step" is one pass of the hot path (SingleFieldInstanton.__init__ + findProfile + findAction for
every instance) over one batch.  Workload c2 = BASELINE.json configs[1]: 1e5 quartic instances
V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2, c_i = 0.02 + 0.475 (i + 1/2)/N, alpha = 2, all findProfile
defaults.  c3 = configs[2]: 1e6 instances, alpha = 3.

Under torchrun (N > 1) every rank owns one GPU and one full batch (weak scaling; the path has no
exchange step, so torch.distributed is used only for the barrier and the max-over-ranks of the time).
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
               flop_per_bounce=7.3e4),
    "c3": dict(n=1000000, alpha=3, name="C3: 1e6 quartic thin->thick sweep, alpha=3, findProfile defaults",
               flop_per_bounce=9.0e4),
}
BYTES_PER_BOUNCE = 64.0  # 16-24 B in + ~48 B out
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (the shoot kernel,
# bounce_kernel<.., MODE_SHOOT>: 0.63 of the 0.90 ms of kernel time per C2 step), from the `ncu --set full`
# captures committed as profiles/r1b_c2_newton_metrics.csv (8.06 MB read: 5.6 MB parameters/minima + 2.4 MB
# phi_bar/rscale/order; results and the 2.4 MB hand-over to the save kernel stay in L2) and
# profiles/r1b_c3_newton_metrics.csv (80.1 MB read + 51.4 MB written)
NCU_TRAFFIC_BYTES = {"c2": 8.056576e6, "c3": 131.504128e6}


def sweep_c(n, rank=0, world=1):
    # every rank gets its own sweep over the same range (offset by a fraction of a grid step)
    return 0.02 + 0.475 * (np.arange(n) + (rank + 0.5) / world) / n


def quartic_params(c):
    p = np.zeros((c.size, 5))
    p[:, 2] = c / 2
    p[:, 3] = -(1 + c) / 3
    p[:, 4] = 0.25
    return p


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
        # (warm-up + timed region + end-to-end loop)
        window = "timed region"
        t0, t1 = self.windows.get("timed", (0, 0))
        sm, mx, reasons = summarise([x for x in self.rows if t0 <= x[0] <= t1])
        if len(sm) < 3:
            window = "under load (warm-up + timed region + end-to-end loop of the same step)"
            t0, t1 = self.windows.get("load", (0, float("inf")))
            sm, mx, reasons = summarise([x for x in self.rows if t0 <= x[0] <= t1])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_oracle_rate(wl, threads, sample_stride, c=None):
    """Times the CPU oracle (C port of the reference algorithm, oracle/ctb_oracle.c) on a strided
    sample of the workload with `threads` host threads.  Returns (bounces/s, results dict, idx)."""
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


def run_reference(args, wl):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the reference
    itself is pure Python and is not on the GPU box) on all host threads, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = len(os.sched_getaffinity(0))
    c = sweep_c(wl["n"])
    # bounded sample per step: ~2 s of wall time at ~9e3 bounces/s/thread
    n_step = int(min(wl["n"], max(2000, 9000 * threads * 2)))
    stride = max(1, wl["n"] // n_step)
    idx = np.arange(0, wl["n"], stride)
    cfg = O.make_config(alpha=wl["alpha"])
    p = quartic_params(c[idx])
    for _ in range(args.warmup):
        O.bounce_batch(cfg, p, 1.0, 0.0, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = O.bounce_batch(cfg, p, 1.0, 0.0, nthreads=threads)
    dt = time.perf_counter() - t0
    value = idx.size * args.steps / dt
    sample = "every %d-th instance of the %d-instance sweep (%d per step)" % (stride, wl["n"], idx.size)
    line = {
        "impl": "reference", "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "instances_per_step": int(idx.size)},
        "cpu_baseline": {"value": value, "unit": "bounces/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "bounces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "solved_fraction": float((res["status"] <= 1).mean()),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--schedule", default="auto", choices=["auto", "sorted", "natural"],
                    help="instance scheduling inside the library (see batch.launch_bounce)")
    ap.add_argument("--order", default="natural", choices=["natural", "reversed", "shuffled"],
                    help="instance order inside the batch (SURVEY.md s.8d: shuffled exposes divergence sensitivity)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    from cosmotransitions_b200 import _lib, batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.load().ctb_device_ok() == 0, "not an sm_100 device"

    n = wl["n"]
    sorted_sched = args.schedule == "sorted" or (args.schedule == "auto" and n >= batch.SORT_THRESHOLD)
    c = sweep_c(n, rank, world)
    if args.order == "reversed":
        c = c[::-1].copy()
    elif args.order == "shuffled":
        c = c[np.random.default_rng(1234).permutation(n)]
    params_h = torch.from_numpy(quartic_params(c)).pin_memory()
    abs_h = torch.ones(n, dtype=torch.float64).pin_memory()
    meta_h = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_params, d_abs, d_meta = params_h.to(dev), abs_h.to(dev), meta_h.to(dev)
    opts = batch.make_opts(alpha=wl["alpha"], block_threads=args.block)
    out = batch.alloc_out(torch, n, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        batch.launch_bounce(_lib.POT_POLY4, d_params, d_abs, d_meta, opts, out, schedule=args.schedule)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # keep the same step running (untimed) until the clocks have settled and the sampler is reporting
    while time.perf_counter() - t_load0 < 0.6:
        step()
        torch.cuda.synchronize()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                 # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark("timed", t_wall0, t_wall0 + t_wall)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(kern_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)
    status = out["status"].cpu().numpy()
    S_gpu = out["S"].cpu().numpy()

    # ---- end to end through the public API: host buffers in, host results out, every step ----

    # BouncePipeline: two slots, each with its own stream -- the H2D copy of step k+1 and the D2H copy of step k-1
    # overlap the kernels of step k; every step's copies happen inside the timed region
    pipe = batch.BouncePipeline(_lib.POT_POLY4, n, depth=2, device=dev, outputs=("S", "status"), alpha=wl["alpha"],
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
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    sampler.mark("load", t_load0, time.perf_counter())
    clocks = sampler.stop()
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall * 1e3)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / (float(t.item()) * 1e-3)
    h2d = params_h.numel() * 8 + abs_h.numel() * 8 + meta_h.numel() * 8
    d2h = n * 8 + n * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- FP64 roofline denominator, measured here: register-resident FMA loop ----
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    import ctypes as C
    ms = C.c_float(0)
    iters = 1 << 16
    best = 0.0
    for _ in range(3):
        _lib.check(_lib.load().ctb_fp64_peak(iters, 148 * 8, 256, sink.data_ptr(), C.byref(ms), None), "fp64_peak")
        best = max(best, 148 * 8 * 256 * iters * 16 * 2 / (ms.value * 1e-3) / 1e12)
    fp64_peak = best
    kern_avg_ms = float(np.mean(kern_ms))
    achieved_tf = wl["flop_per_bounce"] * n / (kern_avg_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_achieved = BYTES_PER_BOUNCE * n / (kern_avg_ms * 1e-3) / 1e9

    line = {
        "metric": "bounce actions/sec", "value": value, "unit": "bounces/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "instances_per_gpu": n, "alpha": wl["alpha"], "order": args.order,
                   "schedule": "work-sorted (setup kernel + device argsort + shoot kernel + save kernel, all inside "
                               "the timed step)" if sorted_sched else "natural",
                   "timing": "CUDA events around each step's kernel on torch's current stream; 256 MiB L2 flush "
                             "between timed iterations", "wall_s_timed_region": t_wall},
        "solved_fraction": float((status <= 1).mean()),
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak if fp64_peak else None,
                     "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                     "traffic_note": "bytes per launch of the shoot kernel, ncu dram__bytes_read+write "
                                     "(profiles/r1b_c2_newton_metrics.csv, r1b_c3_newton_metrics.csv); algorithmic: 64 B x instances",
                     "peak_source": "ctb_fp64_peak register-resident DFMA loop, measured in this run "
                                    "(MEASURED_PEAKS.json has no FP64 entry)",
                     "algorithmic_flop_per_bounce": wl["flop_per_bounce"], "kernel_ms": kern_avg_ms,
                     "scope": "whole step: setup + sort + shoot + save kernels (the flop count per bounce covers "
                              "shooting, dense output and quadrature); the shoot kernel alone is 0.63 of the step's "
                              "kernel time (profiles/r1b_c2_launches.csv)",
                     "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_achieved / hbm_peak, "bytes_per_bounce": BYTES_PER_BOUNCE}},
        "e2e": {"value": e2e_value, "unit": "bounces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "batch.BouncePipeline (public API), pinned host buffers in / pinned host results out every step, two "
                       "slots on two CUDA streams so the copies of neighbouring steps overlap the kernels; wall clock "
                       "around the whole loop including the final synchronisation"},
        # per step: setup_kernel + shoot kernel + save kernel (work-sorted, split solve); fused otherwise
        "gpu_launches": args.steps * (3 if sorted_sched else 1),
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        stride = max(1, n // 100000)
        rate, res, idx = cpu_oracle_rate(wl, threads, stride, c)
        ok = (res["status"] <= 1) & (status[idx] <= 1)
        rel = np.abs(S_gpu[idx][ok] - res["S"][ok]) / np.abs(res["S"][ok])
        line["cpu_baseline"] = {"value": rate, "unit": "bounces/s", "cores": threads, "kind": "port",
                                "reference_python_note": "the reference itself is pure Python and does not travel to "
                                                         "the GPU box; measured in the build container (SURVEY.md "
                                                         "s.6.2): 28.2 bounces/s on 1 core, 221 bounces/s with "
                                                         "multiprocessing.Pool(8), same quartic family, alpha=2",
                                "sample": "every %d-th instance of the workload (%d instances), one pass, C oracle "
                                          "with pthreads" % (stride, idx.size)}
        line["parity"] = {"max_rel_err_S_vs_oracle": float(rel.max()) if rel.size else None,
                          "status_equal": bool(np.array_equal(res["status"], status[idx])), "checked": int(idx.size),
                          "n_rkck_mismatches": int((out["n_rkck"].cpu().numpy()[idx] != res["n_rkck"]).sum()),
                          "n_shots_mismatches": int((out["n_shots"].cpu().numpy()[idx] != res["n_shots"]).sum())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
