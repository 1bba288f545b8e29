#!/usr/bin/env python
"""Wall-clock latency of the scalar drop-in API (one SingleFieldInstanton per call), as the reference's
callers use it (16 findProfile calls per model in model1.findAllTransitions)."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from cosmotransitions_b200 import potentials, tunneling1D as t1  # noqa: E402

pot = potentials.QuarticPotential(0.47)
for _ in range(5):
    inst = t1.SingleFieldInstanton(1.0, 0.0, pot)
    inst.findAction(inst.findProfile())
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for k in range(n):
    inst = t1.SingleFieldInstanton(1.0, 0.0, potentials.QuarticPotential(0.2 + 0.29 * k / n))
    S = inst.findAction(inst.findProfile())
dt = (time.perf_counter() - t0) / n
print("scalar API: %.3f ms per (construct + findProfile + findAction); reference ~16-47 ms" % (dt * 1e3))
