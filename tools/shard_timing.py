#!/usr/bin/env python
"""What one rank of the partitioned C3 batch sees: time launch_bounce on the block-cyclic shard a rank owns at world
size G (1e6-instance quartic sweep, alpha = 3), per block size.  usage: [CTB_LIB=...] python tools/shard_timing.py [G ...]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import _lib, batch, potentials  # noqa: E402

dev = torch.device("cuda", 0)
N = 1000000
c = 0.02 + 0.475 * (np.arange(N) + 0.5) / N
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for G in [int(a) for a in sys.argv[1:]] or [8, 4, 2]:
    blk = np.arange(N) // 256
    idx = np.nonzero(blk % G == 0)[0]
    p = torch.as_tensor(potentials.quartic_family_params(c[idx]), device=dev)
    n = p.shape[0]
    d_abs, d_meta = torch.ones(n, dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    out = batch.alloc_out(torch, n, dev)
    for block in (0, 64, 128):
        opts = batch.make_opts(alpha=3, block_threads=block)
        ts = []
        for it in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); batch.launch_bounce(_lib.POT_POLY4, p, d_abs, d_meta, opts, out); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        print("G=%d n=%d block=%3d: %.3f ms -> %.3e bounces/s per GPU, x%d = %.3e" % (G, n, block, ts[len(ts) // 2], n / ts[len(ts) // 2] * 1e3, G, G * n / ts[len(ts) // 2] * 1e3), flush=True)
