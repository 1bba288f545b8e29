#!/usr/bin/env python
"""Tuning sweep: runs bench.py for every (library build, block size) pair and prints one line each.
usage: python tools/tune.py [--workload c2] lib1.so[,lib2.so...] 32,64,128"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl, orders, sched = "c2", ["natural"], "auto"
for a in sys.argv[1:]:
    if a.startswith("--workload="):
        wl = a.split("=")[1]
    if a.startswith("--schedule="):
        sched = a.split("=")[1]
    if a.startswith("--orders="):
        orders = a.split("=")[1].split(",")
libs = args[0].split(",")
blocks = [int(b) for b in args[1].split(",")]
for lib in libs:
    for b, order in [(b, o) for b in blocks for o in orders]:
        env = dict(os.environ)
        if lib != "default":
            env["CTB_LIB"] = os.path.join(REPO, lib)
        out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "5", "--warmup", "3",
                              "--no-cpu-baseline", "--block", str(b), "--workload", wl, "--order", order, "--schedule", sched], env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        try:
            j = json.loads(out.strip().splitlines()[-1])
            print("%-40s block %4d %-9s %.3e bounces/s  kernel %.3f ms  e2e %.3e" % (
                lib, b, order, j["value"], j["roofline"]["kernel_ms"], j["e2e"]["value"]), flush=True)
        except Exception:
            print(lib, b, "FAILED", out[-400:], flush=True)
