#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gpu_umma_bench2.py > gpurun_out/umma_bench2.log 2>&1; tail -3 gpurun_out/umma_bench2.log
python tools/gpu_conv_timing.py 16 dxn > gpurun_out/conv_timing_dxn.log 2>&1; tail -3 gpurun_out/conv_timing_dxn.log
