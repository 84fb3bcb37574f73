#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LASS_B200_LIB=$PWD/lass_b200/_lib/liblass_b200.so LASS_NO_PROFILE_RUN=1
L="enc0.c1 32->32 @1024x512"
python tools/gpu_one_layer.py "$L" 16 > gpurun_out/one_layer.log 2>&1 || { tail -5 gpurun_out/one_layer.log; exit 1; }
tail -1 gpurun_out/one_layer.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -f -o gpurun_out/prof_r1b_enc0c1 python tools/gpu_one_layer.py "$L" 16 > gpurun_out/ncu_e.log 2>&1; echo "ncu $?"
L="enc0.c2 32->32+id @1024x512 4out"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -f -o gpurun_out/prof_r1b_enc0c2 python tools/gpu_one_layer.py "$L" 16 > gpurun_out/ncu_f.log 2>&1; echo "ncu $?"
