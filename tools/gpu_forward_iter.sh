#!/bin/bash
# dev loop: forward parity tests with the SNR printed, per-launch times, short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py -x -q -s 2>&1 | grep -i "snr\|passed\|failed" | cut -c1-200
timeout 300 python tools/gpu_layer_times.py 64 cur > gpurun_out/layer_times_cur.log 2>&1; head -2 gpurun_out/layer_times_cur.log; tail -1 gpurun_out/layer_times_cur.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(round(d['value']), round(d['e2e']['value']), d['stage_ms'], d['clocks'])"
