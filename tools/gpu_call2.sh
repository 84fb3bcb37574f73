#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sec in umma_probe mask_istft; do
  timeout 300 python tools/gpu_probe.py $sec > gpurun_out/probe_$sec.log 2>&1
  echo "section $sec exit $?" | tee -a gpurun_out/summary2.txt
  tail -c 1800 gpurun_out/probe_$sec.log
done
for pitch in 16 10; do
  timeout 300 python tools/gpu_conv_probe.py $pitch > gpurun_out/conv_probe_p$pitch.log 2>&1
  echo "conv pitch $pitch exit $?" | tee -a gpurun_out/summary2.txt
  tail -c 3000 gpurun_out/conv_probe_p$pitch.log
done
