#!/usr/bin/env python
"""Runs the UNMODIFIED reference (oracle/_ref, see make_ref.py) on the host cores.

TEST INFRASTRUCTURE: only tests/, bench.py's `cpu_baseline` / `--impl reference` legs and smoke() use it.

As a module:   import_reference() -> dict of the reference's modules (matplotlib stubbed, it is absent here);
               quartic_pool(c, alpha, procs, chunksize) -> (S[], wall seconds) with multiprocessing.Pool, the way
               BASELINE.md s.4 prescribes: analytic V, dV, d2V lambdas, SingleFieldInstanton(1.0, 0.0, V, dV, d2V,
               alpha=alpha).findProfile() -> .findAction(), all defaults, chunksize 25.
As a script:   python oracle/ref_runner.py --n 100000 --alpha 2 --stride 16 [--procs P] [--out file.json]
               times the strided subsample c_i = 0.02 + 0.475 (i + 1/2) / n, i = offset, offset + stride, ... and prints
               one JSON object {rate, wall_s, procs, count, S: [...], status: [...]} (status: 0 solved, else the exception name).
               Run it as a subprocess from a CUDA-initialised process (the pool forks).
"""
import argparse
import json
import os
import sys
import time
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "cosmoTransitions", "tunneling1D.py"))


def import_reference():
    if not available():
        raise ImportError("oracle/_ref is not built (python oracle/make_ref.py in the build container)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    warnings.filterwarnings("ignore")
    import cosmoTransitions.finiteT as fT
    import cosmoTransitions.generic_potential as gp
    import cosmoTransitions.helper_functions as hf
    import cosmoTransitions.pathDeformation as pd
    import cosmoTransitions.transitionFinder as tf
    import cosmoTransitions.tunneling1D as t1
    return dict(t1=t1, hf=hf, fT=fT, gp=gp, pd=pd, tf=tf)


_ALPHA = [2]
_T1 = [None]


def _init_worker(alpha):
    _ALPHA[0] = alpha
    _T1[0] = import_reference()["t1"]


def _one_quartic(c):
    """dV = phi (phi - c)(phi - 1): V = phi^4/4 - (1+c)/3 phi^3 + c/2 phi^2."""
    t1 = _T1[0]
    a, b = (1.0 + c) / 3.0, c / 2.0
    V = lambda p: 0.25 * p ** 4 - a * p ** 3 + b * p ** 2
    dV = lambda p: p * (p - c) * (p - 1.0)
    d2V = lambda p: 3 * p * p - 2 * (1 + c) * p + c
    try:
        inst = t1.SingleFieldInstanton(1.0, 0.0, V, dV, d2V, alpha=_ALPHA[0])
        prof = inst.findProfile()
        return float(inst.findAction(prof)), "ok"
    except Exception as e:  # PotentialError / IntegrationError / brentq ValueError: reported, not fatal
        return float("nan"), type(e).__name__


def quartic_band(c, alpha, eps=(0.0, 1e-15, -1e-15, 3e-15, -3e-15, 1e-14, -1e-14, 3e-14, -3e-14, 1e-13, -1e-13)):
    """(min, max) of the reference's S over c (1 + eps): its own sensitivity to the last digits of the input.  At a
    few instances per ten thousand a shot sits on the edge between "converged" and "overshoot" (tunneling1D.py:
    511-513): one bisection step more or less moves S by up to ~1e-5 relative."""
    _init_worker(alpha)
    vals = [_one_quartic(c * (1.0 + e))[0] for e in eps]
    return float(np.nanmin(vals)), float(np.nanmax(vals))


def make_pool(alpha, procs):
    import multiprocessing as mp
    return mp.get_context("fork").Pool(procs, initializer=_init_worker, initargs=(alpha,))


def quartic_pool(c, alpha=2, procs=None, chunksize=25, pool=None):
    procs = procs or len(os.sched_getaffinity(0))
    own = pool is None
    if own:
        pool = make_pool(alpha, procs)
    try:
        pool.map(_one_quartic, [0.3] * procs, chunksize=1)     # workers up, reference imported
        t0 = time.perf_counter()
        res = pool.map(_one_quartic, [float(x) for x in c], chunksize=chunksize)
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    return np.array([r[0] for r in res]), [r[1] for r in res], wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--alpha", type=int, default=2)
    ap.add_argument("--stride", type=int, default=16)
    ap.add_argument("--offset", type=int, default=0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--chunksize", type=int, default=25)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import scipy
    procs = a.procs or len(os.sched_getaffinity(0))
    idx = np.arange(a.offset, a.n, a.stride)
    c = 0.02 + 0.475 * (idx + 0.5) / a.n
    S, status, wall = quartic_pool(c, a.alpha, procs, a.chunksize)
    out = {"rate": len(idx) / wall, "wall_s": wall, "procs": procs, "count": int(len(idx)), "chunksize": a.chunksize,
           "stride": a.stride, "offset": a.offset, "n": a.n, "alpha": a.alpha, "S": S.tolist(), "status": status,
           "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0]}
    text = json.dumps(out)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
