#!/usr/bin/env python
"""One pass over the secondary kernels for an ncu capture (round 2): jspline (random and grid-shaped theta, 2^24),
vtot_model1_grid (4096 x 1024), the finite-T bounce path (minimiser + setup + shoot + save, 200 points x 256 T).
usage: python tools/prof_aux.py [what ...]   what in {jspline, grid, finiteT}   (default: all)"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import _lib, batch, finiteT, generic_potential as gp, scan  # noqa: E402

what = [a for a in sys.argv[1:] if not a.startswith("-")] or ["jspline", "grid", "finiteT"]
dev = torch.device("cuda", 0)
REPS = 2
if "jspline" in what:
    n = 1 << 24
    theta = (torch.rand(n, dtype=torch.float64, device=dev) * 2010.0) - 10.0
    phi_g = torch.linspace(0.0, 400.0, 4096, dtype=torch.float64, device=dev)
    T_g = torch.linspace(20.0, 250.0, 4096, dtype=torch.float64, device=dev)
    theta_g = ((0.1 * phi_g * phi_g)[:, None] / (T_g * T_g)[None, :]).contiguous().reshape(-1)
    for _ in range(REPS):
        finiteT.Jb_spline(theta)
        finiteT.Jb_spline(theta_g)
    torch.cuda.synchronize()
if "grid" in what:
    m = gp.model1()
    x_low = np.array([286.39824989, 382.19727238]); x_high = np.array([231.10296383, -136.82260069])
    L = np.linalg.norm(x_high - x_low); nhat = (x_high - x_low) / L
    s = torch.linspace(-0.2 * L, 1.2 * L, 4096, dtype=torch.float64, device=dev)
    T = torch.linspace(1, 250, 1024, dtype=torch.float64, device=dev)
    buf = torch.empty((4096, 1024), dtype=torch.float64, device=dev)
    for _ in range(REPS):
        m.Vtot_grid(x_low, nhat, s, T, out=buf)
    torch.cuda.synchronize()
if "finiteT" in what:
    P, nT = 200, 256
    rng = np.random.default_rng(20260101)
    lam, Y, nb = 0.12 * rng.uniform(.8, 1.2, P), 0.35 * rng.uniform(.8, 1.2, P), 30. * rng.uniform(.8, 1.2, P)
    lam, Y, nb = (torch.as_tensor(x, dtype=torch.float64, device=dev) for x in (lam, Y, nb))
    win = scan.model1f_temperature_window(lam, Y, nb)
    for _ in range(REPS):
        r = scan.model1f_bounce_scan(lam, Y, nb, nT=nT, window=win)
    torch.cuda.synchronize()
    print("finiteT solved", int((r["status"] <= 1).sum()), "of", P * nT)
print("prof_aux ok")
