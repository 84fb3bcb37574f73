#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
PY
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_igemm|stft|mask_istft|film|preconv" -s 111 -c 37 \
    --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/ncu_bench.log 2>&1
python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -3
