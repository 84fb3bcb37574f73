#!/bin/bash
# bench line + ncu launch list of the same command (per-launch device times; cold-cache, serialised)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_igemm|stft|mask_istft|film|preconv" -s 111 -c 160 \
    --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_bench.log
