#!/usr/bin/env python
"""Latency of small batches: thread-per-instance vs warp-per-instance kernel (quartic family, alpha=2).
usage: python tools/bench_small.py"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import _lib, batch, potentials  # noqa: E402

dev = torch.device("cuda", 0)
rows = []
for n in (1, 32, 256, 1024, 2048, 4096, 8192, 16384):
    c = 0.02 + 0.475 * (np.arange(n) + 0.5) / n
    p = torch.from_numpy(potentials.quartic_family_params(c)).to(dev)
    a, m = torch.ones(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    out = batch.alloc_out(torch, n, dev)
    row = {"n": n}
    for kern in ("thread", "warp"):
        opts = batch.make_opts(alpha=2, kernel=kern)
        for _ in range(3):
            batch.launch_bounce(_lib.POT_POLY4, p, a, m, opts, out, schedule="natural")
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); batch.launch_bounce(_lib.POT_POLY4, p, a, m, opts, out, schedule="natural"); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row[kern + "_ms"] = float(np.median(ts))
    row["speedup_warp"] = row["thread_ms"] / row["warp_ms"]
    rows.append(row)
    print(json.dumps(row), flush=True)
