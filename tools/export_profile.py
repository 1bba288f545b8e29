#!/usr/bin/env python
"""Turn a .ncu-rep (gpurun_out/, scratch) into committed text summaries under profiles/.
usage: python tools/export_profile.py gpurun_out/prof.ncu-rep profiles/r1_name"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], stdout=subprocess.PIPE, text=True).stdout
open(out + "_details.txt", "w").write(det)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
with open(out + "_metrics.csv", "w") as f:
    w = csv.writer(f)
    hdr, units = rows[0], rows[1]
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")
        for k in keys:
            if k in d:
                w.writerow([name, k, units[hdr.index(k)], d[k]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
hot = subprocess.run([sys.executable, __file__.replace("export_profile.py", "ncu_hotlines.py"), "60"], input=src,
                     stdout=subprocess.PIPE, text=True).stdout
open(out + "_hotlines.txt", "w").write(hot)
print("wrote", out + "_{details.txt,metrics.csv,hotlines.txt}")
