#!/bin/bash
# ncu launch list (duration + DRAM bytes) of ONE training step (TAG = $1).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
python tools/gpu_train_step_once.py > gpurun_out/train_once_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/train_once_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/train_launches_$TAG.csv python tools/gpu_train_step_once.py > gpurun_out/ncu_train_list.log 2>&1
echo "ncu train launch list $?"
