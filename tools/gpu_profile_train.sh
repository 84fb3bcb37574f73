#!/bin/bash
# Training-step profile: (1) plain run, (2) ncu launch list (duration + DRAM bytes) of ONE step, (3) `ncu --set full` of the
# largest tcgen05 weight-gradient launch.  Output in gpurun_out/ (TAG = $1, default r2).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
python tools/gpu_train_step_once.py > gpurun_out/train_once_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/train_once_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/train_launches_$TAG.csv python tools/gpu_train_step_once.py > gpurun_out/ncu_train_list.log 2>&1
echo "ncu train launch list $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 3 -c 1 -f \
    -o gpurun_out/prof_${TAG}_wgrad_tc python tools/gpu_train_step_once.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu wgrad_tc $?"
ls -la gpurun_out | tail -8
