#!/bin/bash
# `ncu --set full` captures of selected conv launches of one step (kept < 64 MiB so gpurun brings them back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BATCH=${1:-16}
CMD="python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline --no-spectral"
$CMD > gpurun_out/bench_b$BATCH.json 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_b$BATCH.json; exit 1; }
# conv launch order within a step: enc k -> 2k, 2k+1 ; dec j -> 14+3j (up), +1 (c1), +2 (c2)
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 96 -c 6 -f -o gpurun_out/prof_conv_enc012 $CMD > gpurun_out/ncu_full_a.log 2>&1
echo "ncu a exit $?"
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 122 -c 6 -f -o gpurun_out/prof_conv_dec45 $CMD > gpurun_out/ncu_full_b.log 2>&1
echo "ncu b exit $?"
ls -la gpurun_out/
