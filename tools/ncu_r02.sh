#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU; numbers printed by these runs are never bench values).
# Launch lists (csv) and `--set full` reports of the named kernels; every report is summarised ON THE BOX
# (profiles/summarize_ncu.py) and then deleted -- gpurun_out/ only travels back below 64 MiB.
cd "$(dirname "$0")/.."
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$LIST -c 8000 --log-file gpurun_out/launches_r02_build.csv $B --only build > gpurun_out/ncu_list_build.log 2>&1; echo "list build rc=$?"
$LIST -c 800 --log-file gpurun_out/launches_r02_query.csv $B --only query > gpurun_out/ncu_list_query.log 2>&1; echo "list query rc=$?"
: > gpurun_out/ncu_r02_summary.txt
cap() {  # cap <mode> <kernel regex> <count>
  timeout 240 $FULL -k "regex:$2" -c $3 -o gpurun_out/r02_$1 python tools/ncu_targets.py $1 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/ncu_$1.log)"
  REPS="$REPS gpurun_out/r02_$1.ncu-rep"
}
cap c1_train "tc_pq_kernel|tc_filter_kernel|cluster_sort_kernel|update_stats_kernel|epilogue_kernel|pq_fallback_kernel|rerank_kernel" 14
cap c1_transform "tc_pq_kernel|tc_filter_kernel|rerank_kernel|pq_fallback_kernel|row_norms|residual" 10
cap c1_query "ivfpq_scan" 3
cap c2_transform "tc_filter_general_kernel|tc_pq_kernel|cand_exact_kernel" 5
cap c4_assign "tc_filter_general_kernel" 2
python profiles/summarize_ncu.py --traffic gpurun_out/ncu_traffic.json $REPS > gpurun_out/ncu_r02_summary.txt 2> gpurun_out/ncu_summarize.err
# keep the scan kernel's source-level page (small) and drop the reports
ncu -i gpurun_out/r02_c1_query.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/r02_c1_query_source.csv
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
