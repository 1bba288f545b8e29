#!/usr/bin/env python
"""Small run that touches every kernel; meant to be executed under compute-sanitizer
(`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`, one tool per gpurun call)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import _lib, batch, finiteT, generic_potential as gp, potentials, scan  # noqa: E402
from scipy import interpolate  # noqa: E402

c = 0.02 + 0.478 * (np.arange(300) + 0.5) / 300
p = potentials.quartic_family_params(c)
for kern in ("thread", "warp"):
    for alpha in (2, 3):
        r = batch.find_bounce_batch(_lib.POT_POLY4, p, 1.0, 0.0, alpha=alpha, kernel=kern, return_profiles=(alpha == 2))
        assert (r["status"].numpy() <= 1).sum() > 250
r = batch.find_bounce_batch(_lib.POT_POLY4, np.tile(p, (20, 1)), 1.0, 0.0, schedule="sorted", kernel="thread")
lam, Y, n = np.full(3, 0.12), np.full(3, 0.35), np.full(3, 30.0)
out = scan.model1f_bounce_scan(lam, Y, n, nT=24)
assert np.isfinite(out["Tnuc"].cpu().numpy()).all()
m = gp.model1()
x_low, x_high = np.array([286.39824989, 382.19727238]), np.array([231.10296383, -136.82260069])
scan.model1_line_bounce_scan(m.model8, x_low, x_high, np.linspace(80, 100, 9))
m.Vtot_grid(x_low, (x_high - x_low) / np.linalg.norm(x_high - x_low), np.linspace(-50, 600, 77), np.linspace(1, 200, 53))
m.Vtot(np.random.default_rng(0).uniform(-300, 300, (101, 2)), np.linspace(0, 200, 101))
th = np.random.default_rng(1).uniform(-10, 2000, 5000)
for d in range(4):
    finiteT.Jb_spline(th, d); finiteT.Jf_spline(th, d)
x = np.linspace(0, 20, 333)
for d in range(4):
    finiteT.Jb(x, approx="high", deriv=d); finiteT.Jf(x, approx="high", deriv=d)
finiteT.Jb(x, approx="low"); finiteT.Jf(x, approx="low")
xs = np.linspace(-0.3, 1.3, 140)
tck = interpolate.splrep(xs, 0.25 * xs ** 4 - 0.4 * xs ** 3 + 0.1 * xs ** 2, s=0)
pot = potentials.SplinePotential(tck)
for kern in ("thread", "warp"):
    r = batch.find_bounce_batch(pot.kind, pot.params[None, :], [1.0], [0.0], kernel=kern)
    assert int(r["status"][0]) <= 1 and abs(float(r["S"][0]) - 6.6489883) < 1e-3, r["S"]
pot(np.linspace(0, 1, 50))
torch.cuda.synchronize()
print("sanitize smoke ok")
