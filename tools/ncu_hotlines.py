#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output by source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_hotlines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
cur, hdr, agg, tot, tots = None, None, {}, 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line, inst, samp = int(r[0]), int(r[7]), int(r[6] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur.split("/")[-1], line), [0, 0, r[1][:100]])
    a[0] += inst
    a[1] += samp
    tot += inst
    tots += samp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("total warp-instructions executed (source-attributed): %d ; stall samples: %d" % (tot, tots))
print("%-18s %5s %8s %8s  %s" % ("file", "line", "inst %", "samp %", "source"))
for (f, l), (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print("%-18s %5d %7.2f%% %7.2f%%  %s" % (f, l, 100.0 * inst / tot, 100.0 * samp / max(tots, 1), src))
