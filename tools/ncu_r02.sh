#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU; numbers printed by these runs are never bench values).
# Outputs under gpurun_out/: launch lists (csv) and `--set full` reports; summarised into profiles/ by
# profiles/summarize_launches.py and profiles/summarize_ncu.py.
set -x
cd "$(dirname "$0")/.."
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on -f"
$LIST -c 6000 --log-file gpurun_out/launches_r02_build.csv $B --only build > gpurun_out/ncu_list_build.log 2>&1
$LIST -c 600 --log-file gpurun_out/launches_r02_query.csv $B --only query > gpurun_out/ncu_list_query.log 2>&1
$FULL -o gpurun_out/r02_pq -k regex:"tc_pq_kernel|pq_fallback_kernel|residual_norms" -c 106 $B --only build > gpurun_out/ncu_pq.log 2>&1
$FULL -o gpurun_out/r02_ivf -k regex:"tc_filter_kernel|rerank_kernel|assign_tile_kernel|split_merge|row_norm" -c 170 $B --only build > gpurun_out/ncu_ivf.log 2>&1
$FULL -o gpurun_out/r02_update -k regex:"cluster_sort|update_stats|epilogue_kernel|group_kernel|hist_kernel" -s 12 -c 18 $B --only build > gpurun_out/ncu_update.log 2>&1
$FULL -o gpurun_out/r02_query -k regex:"ivfpq_scan|merge_kernel|select_probes|refine_kernel" -c 8 $B --only query > gpurun_out/ncu_query.log 2>&1
$FULL -o gpurun_out/r02_c2 -k regex:"tc_filter_general|cand_exact|gather_split|tc_pq_kernel|pq_fallback" -s 4 -c 14 python bench.py --config C2 --rows 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
$FULL -o gpurun_out/r02_c4 -k regex:"tc_filter_general|cand_exact|gather16|ivfflat_scan|group_vectors" -s 2 -c 10 python bench.py --config C4 --rows 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
