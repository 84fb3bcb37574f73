#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 0x3 0x3c00000 0x3c00003 0x3c00001 0x3c0000f 0x3000003 0xc00000; do
  LASS_DXN_MASK=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spectral > gpurun_out/b_$m.json 2>gpurun_out/b_$m.err || tail -3 gpurun_out/b_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/b_$m.json')); print('$m', round(d['value']), round(d['stage_ms']['unet_convs'],2))"
done
LASS_DXN_MASK=0x3c00003 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -2
