#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LASS_TIMING_FLAGS=0,1,2,16,256 timeout 300 python tools/gpu_conv_timing.py 16 > gpurun_out/conv_timing.log 2>&1; cat gpurun_out/conv_timing.log | cut -c1-150
for spec in "enc0c2|enc0.c2" "dec5up|dec5.up"; do
  tag=${spec%%|*}; L=${spec#*|}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -f -o gpurun_out/prof_r1c_$tag python tools/gpu_one_layer.py "$L" 16 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag $?"
done
