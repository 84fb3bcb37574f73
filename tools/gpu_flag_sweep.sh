#!/bin/bash
# per-launch times of the whole model under several conv debug-flag settings: tools/gpu_flag_sweep.sh 0 8192 16384 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in "$@"; do
  LASS_CONV_FLAGS=$f timeout 300 python tools/gpu_layer_times.py 64 f$f > gpurun_out/layer_times_f$f.log 2>&1; echo "flags $f: $(tail -1 gpurun_out/layer_times_f$f.log)"
done
