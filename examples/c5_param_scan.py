#!/usr/bin/env python
"""BASELINE.json configs[4]: model-parameter scan, each point with a 256-step temperature bounce sweep
(one-field model1-style thermal potential, (lambda, Y, n) uniform +-20 % around
(0.12, 0.35, 30), default_rng(20260101)).  usage: c5_param_scan.py [P]
(parity of the same scan against the CPU oracle: tests/test_gpu_parity.py)
Under torchrun every rank takes the parameter points rank, rank+world, ... (no exchange between ranks)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import scan, multigpu  # noqa: E402

P = int(next((a for a in sys.argv[1:] if a.isdigit()), 1000))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
rng = np.random.default_rng(20260101)
lam, Y, n = 0.12 * rng.uniform(.8, 1.2, P), 0.35 * rng.uniform(.8, 1.2, P), 30. * rng.uniform(.8, 1.2, P)
idx = multigpu.partition(P, world, rank)
scan.model1f_bounce_scan(lam[idx][:8], Y[idx][:8], n[idx][:8], nT=16)   # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
win = scan.model1f_temperature_window(lam[idx], Y[idx], n[idx])
torch.cuda.synchronize()
t_win = time.perf_counter() - t0
t0 = time.perf_counter()
r = scan.model1f_bounce_scan(lam[idx], Y[idx], n[idx], nT=256, window=win)
torch.cuda.synchronize()
t_scan = time.perf_counter() - t0
st = r["status"].cpu().numpy()
Tn = r["Tnuc"].cpu().numpy()
out = {"rank": rank, "points": len(idx), "bounces": int(st.size), "window_s": t_win, "scan_s": t_scan,
       "bounces_per_s": st.size / t_scan, "solved": int((st <= 1).sum()),
       "status_hist": np.bincount(st.reshape(-1), minlength=11).tolist(),
       "Tnuc_found": int(np.isfinite(Tn).sum()), "Tnuc_mean": float(np.nanmean(Tn))}
print(json.dumps(out))
