#!/bin/bash
# Full GPU test suite on the box: python -m pytest tests -m gpu; log in gpurun_out/gpu_tests.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu "$@" > gpurun_out/gpu_tests.log 2>&1
rc=$?
tail -15 gpurun_out/gpu_tests.log
exit $rc
