#!/bin/bash
# usage: tools/multi_gpu_r02.sh N [small]   -- C1 scaling line + C3/C4/C5 sharded builds on N GPUs of one box
N=$1; MODE=${2:-full}
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/nccl_check.py 2>&1 | grep "nccl_check" | tail -2
$TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "C1 rc=$?"
if [ "$MODE" = small ]; then R3="--rows 2000000"; R4="--rows 1500000"; R5="--rows 4000000"; else R3=""; R4=""; R5=""; fi
$TR --master-port 29513 bench.py --gpus $N --config C3 $R3 --steps 1 --warmup 1 > gpurun_out/bench_r02_C3_${N}gpu.json 2> gpurun_out/bench_r02_C3_${N}gpu.err; echo "C3 rc=$?"
$TR --master-port 29514 bench.py --gpus $N --config C4 $R4 --steps 1 --warmup 1 > gpurun_out/bench_r02_C4_${N}gpu.json 2> gpurun_out/bench_r02_C4_${N}gpu.err; echo "C4 rc=$?"
$TR --master-port 29515 bench.py --gpus $N --config C5 $R5 --steps 1 --warmup 1 > gpurun_out/bench_r02_C5_${N}gpu.json 2> gpurun_out/bench_r02_C5_${N}gpu.err; echo "C5 rc=$?"
tail -2 gpurun_out/bench_r02_*_${N}gpu.err
