#!/bin/bash
# Round-2 ncu captures of the shipping kernels (run on the B200 box through gpurun; outputs under gpurun_out/).
# usage: bash tools/capture_profiles.sh TAG      -> gpurun_out/TAG_*.{csv,ncu-rep,log}
# Every capture is bounded (-c) and runs under its own timeout; a number printed under ncu is never a bench value.
T=${1:-r2s}
O=gpurun_out
NCU="ncu --clock-control none"
B="--no-cpu-baseline --steps 2 --warmup 3"
# launch lists (per-launch durations of one bench command each)
for w in c2 c3 c5 c4 c5m; do
  timeout 240 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${T}_${w}_launches.csv python bench.py --workload $w $B > $O/${T}_${w}_launches.log 2>&1
done
# full captures of the dominant kernels
timeout 240 $NCU --set full --import-source on -k regex:"bounce_kernel|setup_kernel" --launch-skip 9 -c 3 -o $O/${T}_c2 -f python bench.py --workload c2 $B > $O/${T}_c2_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"bounce_kernel" --launch-skip 6 -c 2 -o $O/${T}_c3 -f python bench.py --workload c3 $B > $O/${T}_c3_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"bounce_kernel|rows_kernel" --launch-skip 9 -c 3 -o $O/${T}_c5 -f python bench.py --workload c5 $B > $O/${T}_c5_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"trace_minima" --launch-skip 2 -c 2 -o $O/${T}_c5m -f python bench.py --workload c5m $B > $O/${T}_c5m_ncu.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:"vtot_model1_grid|jspline" -c 6 -o $O/${T}_aux -f python tools/prof_aux.py jspline grid > $O/${T}_aux_ncu.log 2>&1
ls -la $O | grep ${T}_
