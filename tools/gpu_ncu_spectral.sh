#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gpu_spectral_bench.py > gpurun_out/spectral_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mask_istft|stft_gemm" -s 4 -c 2 -f -o gpurun_out/prof_spectral \
  python tools/gpu_spectral_bench.py > gpurun_out/ncu_spectral.log 2>&1
echo "exit $?"; ls -la gpurun_out/
