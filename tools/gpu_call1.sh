#!/bin/bash
# first GPU bring-up call: each section in its own process so a sticky CUDA error cannot hide the others
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
for sec in umma_probe stft mask_istft umma_probe_risky; do
  timeout 300 python tools/gpu_probe.py $sec > gpurun_out/probe_$sec.log 2>&1
  echo "section $sec exit $?" | tee -a gpurun_out/summary.txt
  tail -c 1500 gpurun_out/probe_$sec.log
done
