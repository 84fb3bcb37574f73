#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py tests/test_next_rows.py -x -q -m gpu > gpurun_out/pytest_spectral.log 2>&1; tail -15 gpurun_out/pytest_spectral.log
python tools/gpu_spectral_bench.py 2>&1 | tail -4
