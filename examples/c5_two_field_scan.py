#!/usr/bin/env python
"""BASELINE.json configs[4] on the real two-field model (examples/testModel1.py of the reference): P parameter points x
256 temperatures.  Both phases of the first-order transition are followed on the device (ctb_trace_minima), then one
bounce per (point, T) is solved along the straight line between the minima; T_crit and T_nuc (S3/T = 140) per point.
The same through a run-time compiled copy of the model (jit.CompiledModel): what a user-defined model would run.

usage: python examples/c5_two_field_scan.py [P]          (torchrun --nproc-per-node G ...: points r, r + G, ...)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cosmotransitions_b200 import jit, phases, potentials  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
rng = np.random.default_rng(20260102)
jitter = rng.uniform(.9, 1.1, (P, 5))
rows = np.stack([potentials.model1_model8(120. * a, 50. * b, 25. * c, .1 * d, .15 * f) for a, b, c, d, f in jitter])[rank::world]
T = np.linspace(40.0, 135.0, 256)
phases.model1_phase_scan(rows[:8], T)                       # warm-up (module load, table upload)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = phases.model1_phase_scan(rows, T)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
Tn = r["Tnuc"].cpu().numpy()
print("rank %d: %d points x %d T in %.3f s: %d bounces (%d solved), %d points nucleate, median Tcrit %.2f, median Tnuc %.2f"
      % (rank, len(rows), len(T), dt, r["n_bounces"], int(((r["status"] == 0) | (r["status"] == 1)).sum()),
         int(np.isfinite(Tn).sum()), float(np.nanmedian(r["Tcrit"].cpu().numpy())), float(np.nanmedian(Tn))))
if rank == 0:
    m = jit.model1_compiled()                               # the same model as a CUDA string compiled at run time
    sub = rows[:min(64, len(rows))]
    v = np.sqrt(sub[:, 7])
    t0 = time.perf_counter()
    g = phases.phase_scan(m, sub, T, np.stack([v, v], -1))   # bounce kernels instantiated for the compiled model
    torch.cuda.synchronize()
    d = np.abs(g["Tnuc"].cpu().numpy() - Tn[:len(sub)])
    print("compiled model, %d points: %.3f s; |Tnuc - built-in| max %.3g K" % (len(sub), time.perf_counter() - t0, np.nanmax(d)))
