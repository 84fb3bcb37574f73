#!/bin/bash
# dev loop for the CTA-pair (cta_group::2) conv path: parity of the pair cases first (short timeout: a protocol bug hangs),
# then per-launch times of the whole model with and without pairs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PAIR_CASES="c128_128 nopair_c128_128 c256_256 sc_big sc_pool12 c768_384 pair_c128_128_long pair_c256_256_long pair_c512_256 pair_c384_384_sc pair_slice nopair_pair_c512_256"
timeout 150 python tools/gpu_conv_probe.py 10 $PAIR_CASES > gpurun_out/pair_probe.log 2>&1
rc=$?
cat gpurun_out/pair_probe.log | cut -c1-220
if [ $rc -ne 0 ] || grep -q error gpurun_out/pair_probe.log; then echo "pair probe failed rc=$rc"; nvidia-smi --query-gpu=name,memory.used --format=csv; exit 0; fi
timeout 300 python tools/gpu_layer_times.py 64 pair > gpurun_out/layer_times_pair.log 2>&1; tail -33 gpurun_out/layer_times_pair.log
LASS_CONV_FLAGS=4096 timeout 300 python tools/gpu_layer_times.py 64 nopair > gpurun_out/layer_times_nopair.log 2>&1; tail -1 gpurun_out/layer_times_nopair.log
