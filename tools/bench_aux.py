#!/usr/bin/env python
"""Throughput of the kernels around the bounce solver (SURVEY.md s.8f rows + the rest of the tunneling1D / finiteT
surface): findAction on given profiles (HBM-bound), WallWithConstFriction, exact / series Jb, 1-D minimiser.
Prints one JSON object.  usage: python tools/bench_aux.py"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import _lib, batch, finiteT, potentials, scan  # noqa: E402

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


out = {}
# --- findAction on given profiles: 2e5 profiles x 750 points (3.6 GB of R / Phi / dPhi: larger than L2)
n, P = 200000, 750
c = 0.02 + 0.475 * (np.arange(n) + 0.5) / n
params = torch.from_numpy(potentials.quartic_family_params(c)).to(dev)
R = (torch.linspace(0.0, 30.0, P, dtype=torch.float64, device=dev)[None, :] * torch.ones((n, 1), dtype=torch.float64, device=dev)).contiguous()
Phi = (1.0 / (1.0 + torch.exp(R - 15.0))).contiguous()
dPhi = (-Phi * (1.0 - Phi)).contiguous()
meta = torch.zeros(n, dtype=torch.float64, device=dev)
ms = timeit(lambda: batch.find_action_batch(_lib.POT_POLY4, params, meta, R, Phi, dPhi, alpha=2))
byts = 24.0 * n * P + 8.0 * n
out["find_action"] = {"profiles": n, "points": P, "ms": ms, "profiles_per_s": n / ms * 1e3, "GBps": byts / ms * 1e-6,
                      "hbm_frac": byts / ms * 1e-6 / HBM, "bytes_per_point": 24}
del R, Phi, dPhi
# --- WallWithConstFriction, 1e5 quartic instances
n = 100000
c = 0.2 + 0.295 * (np.arange(n) + 0.5) / n
pw = torch.from_numpy(potentials.quartic_family_params(c)).to(dev)
one, zero = torch.ones(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
res = {}
ms = timeit(lambda: res.update(batch.find_wall_batch(_lib.POT_POLY4, pw, one, zero)), reps=3)
out["wall"] = {"instances": n, "ms": ms, "walls_per_s": n / ms * 1e3, "solved": int((res["status"] <= 1).sum())}
# --- Jb exact (device quadrature), series, and the bounded minimiser
x = torch.rand(1 << 20, dtype=torch.float64, device=dev) * 20.0
ms = timeit(lambda: finiteT.Jb(x, approx="exact"), reps=3)
out["jb_exact"] = {"n": x.numel(), "ms": ms, "evals_per_s": x.numel() / ms * 1e3}
th = torch.rand(1 << 20, dtype=torch.float64, device=dev) * 12.0 - 6.0
ms = timeit(lambda: finiteT.Jb_exact2(th), reps=3)
out["jb_exact2_mixed_sign"] = {"n": th.numel(), "ms": ms, "evals_per_s": th.numel() / ms * 1e3}
x = torch.rand(1 << 24, dtype=torch.float64, device=dev) * 20.0 + 0.1
ms = timeit(lambda: finiteT.Jb(x, approx="high", n=8))
out["jb_high_n8"] = {"n": x.numel(), "ms": ms, "evals_per_s": x.numel() / ms * 1e3}
n = 1 << 20
c = 0.02 + 0.475 * (np.arange(n) + 0.5) / n
pm = torch.from_numpy(potentials.quartic_family_params(c)).to(dev)
lo, hi = torch.full((n,), 0.6, dtype=torch.float64, device=dev), torch.full((n,), 1.5, dtype=torch.float64, device=dev)
ms = timeit(lambda: scan.minimize_1d(_lib.POT_POLY4, pm, lo, hi))
out["minimize_1d_quartic"] = {"n": n, "ms": ms, "minima_per_s": n / ms * 1e3}
# --- generic_potential.one_field_model.Vtot on a 4096 x 1024 (phi, T) grid (5 bosons incl. a thermal mass, 1 fermion)
from cosmotransitions_b200 import generic_potential as gpm  # noqa: E402
v2 = 246. ** 2
mdl = gpm.one_field_model([0.0, 0.0, -.5 * 0.13 * v2, 0.0, .25 * 0.13], v2,
                          bosons=[(-0.13 * v2, 0.39, 0.0, 1., 1.5), (-0.13 * v2, 0.13, 0.0, 3., 1.5), (0.0, 0.105, 0.0, 6., 5. / 6),
                                  (0.0, 0.1375, 0.0, 3., 5. / 6), (0.0, 0.9, 0.15, 12., 1.5)], fermions=[(0.0, 0.49, 12.)])
Xg = torch.linspace(-20.0, 320.0, 4096, dtype=torch.float64, device=dev)[:, None, None]
Tg = torch.linspace(1.0, 250.0, 1024, dtype=torch.float64, device=dev)[None, :]
ms = timeit(lambda: mdl.Vtot(Xg, Tg))
out["vtot_one_field_grid"] = {"points": 4096 * 1024, "ms": ms, "pts_per_s": 4096 * 1024 / ms * 1e3,
                              "note": "includes the broadcast / concat of the (phi, T) pairs in torch (24 B per point in and out)"}
print(json.dumps(out))
