#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LASS_DXN_MASK=0 timeout 300 python tools/gpu_layer_times.py 64 k > gpurun_out/layer_times_k.log 2>&1; tail -1 gpurun_out/layer_times_k.log
timeout 300 python tools/gpu_layer_times.py 64 dxn > gpurun_out/layer_times_dxn.log 2>&1; tail -1 gpurun_out/layer_times_dxn.log
for m in 0 0x3c00001; do
LASS_DXN_MASK=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_m$m.json 2>gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m$m.json')); print('$m', round(d['value']), d['stage_ms'], d['clocks'])"
done
