#!/bin/bash
# Round profile: (1) plain bench line + reference arm, (2) per-launch duration + DRAM bytes of one step (ncu, --clock-control
# none), (3) `ncu --set full` of representative launches: epilogue/store-bound enc0.c2, MMA-issue-bound dec5.c1 (resident-weight
# CTA pair), the N = 256 CTA pair dec2.c1, the N = 128 CTA pair dec3.c1, the resident pair dec4.c1 (single-layer runs of
# tools/gpu_one_layer.py at batch 16) and the spectral kernels K1 / K5.
# gpurun brings back at most 64 MiB: PARTS selects what runs (default "bench list enc0c2 dec5c1 dec2c1"; second call
# PARTS="dec3c1 dec4c1 k1 k5").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1d}
PARTS=${PARTS:-bench list enc0c2 dec5c1 dec2c1}
has() { [[ " $PARTS " == *" $1 "* ]]; }
if has bench; then
  python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo bench failed; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
fi
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral --no-extras"
if has list; then
  $CMD > gpurun_out/bench_short.json 2>&1 || { echo "short bench failed"; exit 1; }
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"conv_igemm|stft|mask_istft|film|preconv" -s 108 -c 36 --csv --log-file gpurun_out/step_launches_$TAG.csv $CMD > gpurun_out/ncu_a.log 2>&1
  echo "ncu launch list $?"
fi
for spec in "enc0c2|enc0.c2" "dec5c1|dec5.c1" "dec2c1|dec2.c1" "dec3c1|dec3.c1" "dec4c1|dec4.c1"; do
  tag=${spec%%|*}; L=${spec#*|}
  has $tag || continue
  timeout 300 python tools/gpu_one_layer.py "$L" 16 > gpurun_out/one_$tag.log 2>&1 || { echo "one_layer $tag failed"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$tag python tools/gpu_one_layer.py "$L" 16 > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag $?"
done
if has k1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stft_gemm" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_k1 python tools/gpu_spectral_bench.py 1024 160 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 $?"
fi
if has k5; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_istft" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_k5 python tools/gpu_spectral_bench.py 1024 160 > gpurun_out/ncu_k5.log 2>&1; echo "ncu k5 $?"
fi
ls -la gpurun_out | tail -15
