#!/bin/bash
# dev loop for the conv kernel: conv + forward parity tests, knock-out timings of plan-configured launches
# (TFLAGS / TFILTER), per-launch times of the whole model and a short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py -x -q > gpurun_out/pytest_conv.log 2>&1; tail -3 gpurun_out/pytest_conv.log
LASS_TIMING_FLAGS=${TFLAGS:-0,1,2} timeout 300 python tools/gpu_conv_timing.py 16 $TFILTER > gpurun_out/conv_timing.log 2>&1; cat gpurun_out/conv_timing.log | cut -c1-150
timeout 300 python tools/gpu_layer_times.py 64 cur > gpurun_out/layer_times_cur.log 2>&1; tail -1 gpurun_out/layer_times_cur.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(round(d['value']), d['stage_ms'], d['clocks'])"
