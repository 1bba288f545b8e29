#!/usr/bin/env python
"""BASELINE.json configs[3]: testModel1 finite-T potential -- batched Vtot with Jb/Jf on a 4096 x 1024
(phi, T) grid, plus a per-T 1-D bounce scan toward T_nuc along the straight line between the two phases
(SURVEY.md s.8d C4).  Prints timings and the S3/T = 140 crossing (parity of the same scan against the CPU oracle: tests/test_gpu_parity.py)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import generic_potential as gp, scan  # noqa: E402

m = gp.model1()
x_low = np.array([286.39824989, 382.19727238])
x_high = np.array([231.10296383, -136.82260069])
L = np.linalg.norm(x_high - x_low)
nhat = (x_high - x_low) / L
s = torch.linspace(-0.2 * L, 1.2 * L, 4096, dtype=torch.float64, device="cuda")
T = torch.linspace(1, 250, 1024, dtype=torch.float64, device="cuda")
buf = torch.empty((4096, 1024), dtype=torch.float64, device="cuda")
for _ in range(3):
    m.Vtot_grid(x_low, nhat, s, T, out=buf)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    m.Vtot_grid(x_low, nhat, s, T, out=buf)
torch.cuda.synchronize()
t_grid = (time.perf_counter() - t0) / 20

Ts = np.linspace(77.8, 109.2, 512)
scan.model1_line_bounce_scan(m.model8, x_low, x_high, Ts)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = scan.model1_line_bounce_scan(m.model8, x_low, x_high, Ts)
torch.cuda.synchronize()
t_scan = time.perf_counter() - t0
st = r["status"].cpu().numpy()
out = {"vtot_grid_s": t_grid, "vtot_pts_per_s": 4096 * 1024 / t_grid, "scan_T": len(Ts), "scan_s": t_scan,
       "bounces_per_s": len(Ts) / t_scan, "solved": int((st <= 1).sum()), "Tnuc_line": float(r["Tnuc"])}

print(json.dumps(out))
