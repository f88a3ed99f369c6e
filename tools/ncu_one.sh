#!/bin/bash
# usage: tools/ncu_one.sh <ncu_targets mode> <kernel regex> <count> <tag>
# one `--set full` capture, summarised on the box (metrics + per-source-line stall samples); the report is kept
# only when it is small (gpurun_out/ travels back below 64 MiB)
cd "$(dirname "$0")/.."
MODE=$1; RE=$2; CNT=$3; TAG=$4
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k "regex:$RE" -c $CNT \
  -o gpurun_out/$TAG python tools/ncu_targets.py $MODE > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$? $(tail -1 gpurun_out/ncu_$TAG.log)"
python profiles/summarize_ncu.py gpurun_out/$TAG.ncu-rep > gpurun_out/${TAG}_summary.txt 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source.csv.gz
ls -la gpurun_out/
find gpurun_out -name "*.ncu-rep" -size +20M -delete
