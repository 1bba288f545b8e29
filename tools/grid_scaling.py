#!/usr/bin/env python
"""Time ctb_vtot_model1_grid against the number of field rows (fixed 1024 temperatures): separates the per-launch /
per-block fixed cost (table staging, first look-ups) from the per-point cost.  usage: python tools/grid_scaling.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import generic_potential as gp  # noqa: E402

dev = torch.device("cuda", 0)
m = gp.model1()
x_low = np.array([286.39824989, 382.19727238]); x_high = np.array([231.10296383, -136.82260069])
L = np.linalg.norm(x_high - x_low); nhat = (x_high - x_low) / L
T = torch.linspace(1, 250, 1024, dtype=torch.float64, device=dev)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for ns in (64, 256, 1024, 2048, 4096, 8192, 16384, 32768):
    s = torch.linspace(-0.2 * L, 1.2 * L, ns, dtype=torch.float64, device=dev)
    buf = torch.empty((ns, 1024), dtype=torch.float64, device=dev)
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.Vtot_grid(x_low, nhat, s, T, out=buf); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print("ns %6d  median %.1f us  min %.1f us  -> %.2f ns/row" % (ns, ts[len(ts) // 2], ts[0], ts[len(ts) // 2] * 1e3 / ns))
