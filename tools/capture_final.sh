T=r2f
O=gpurun_out
NCU="ncu --clock-control none"
B="--no-cpu-baseline --steps 2 --warmup 3"
for w in c2 c3; do
  timeout 240 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${T}_${w}_launches.csv python bench.py --workload $w $B > $O/${T}_${w}_launches.log 2>&1
done
timeout 240 $NCU --set full --import-source on -k regex:"bounce_kernel|setup_kernel" --launch-skip 9 -c 3 -o $O/${T}_c2 -f python bench.py --workload c2 $B > $O/${T}_c2_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"bounce_kernel" --launch-skip 6 -c 2 -o $O/${T}_c3 -f python bench.py --workload c3 $B > $O/${T}_c3_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"bounce_kernel" --launch-skip 4 -c 2 -o $O/${T}_c5 -f python bench.py --workload c5 $B > $O/${T}_c5_ncu.log 2>&1
ls -la $O | grep ${T}_
