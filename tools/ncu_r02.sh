#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU; numbers printed by these runs are never bench values).
# Outputs under gpurun_out/: launch lists (csv) and `--set full` reports; summarised into profiles/ by
# profiles/summarize_launches.py and profiles/summarize_ncu.py.
cd "$(dirname "$0")/.."
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$LIST -c 8000 --log-file gpurun_out/launches_r02_build.csv $B --only build > gpurun_out/ncu_list_build.log 2>&1; echo "list build rc=$?"
$LIST -c 800 --log-file gpurun_out/launches_r02_query.csv $B --only query > gpurun_out/ncu_list_query.log 2>&1; echo "list query rc=$?"
for m in c1_train c1_transform c1_query c2_transform c4_assign; do
  $FULL -o gpurun_out/r02_$m python tools/ncu_targets.py $m > gpurun_out/ncu_$m.log 2>&1; echo "$m rc=$?"; tail -1 gpurun_out/ncu_$m.log
done
ls -la gpurun_out/*.ncu-rep
