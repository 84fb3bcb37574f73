#!/bin/bash
# Round profile: (1) per-launch duration + DRAM bytes of one step, (2) `ncu --set full` of three representative conv
# launches (store-bound enc0.c2, MMA-count-bound dec5.c1, tensor-bound dec2.c1) and of the spectral kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-spectral"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { echo plain run failed; tail -5 gpurun_out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv_igemm|stft|mask_istft|film|preconv" -s 111 -c 37 --csv --log-file gpurun_out/step_dram.csv $CMD > gpurun_out/ncu_a.log 2>&1
echo "ncu a $?"
# conv launch indices within a step (0-based among conv launches): enc0.c2 = 1, dec5.c1 = 30, dec2.c1 = 21
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 97 -c 1 -f -o gpurun_out/prof_enc0c2 $CMD > gpurun_out/ncu_b.log 2>&1; echo "ncu b $?"
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 126 -c 1 -f -o gpurun_out/prof_dec5c1 $CMD > gpurun_out/ncu_c.log 2>&1; echo "ncu c $?"
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" -s 117 -c 1 -f -o gpurun_out/prof_dec2c1 $CMD > gpurun_out/ncu_d.log 2>&1; echo "ncu d $?"
ls -la gpurun_out
