#!/bin/bash
# dev loop for the streamed-weight conv paths (CTA pairs, row stages): parity of the affected cases first (short timeout: a
# protocol bug hangs), then per-launch times of the whole model under the flag settings given as arguments.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python - > gpurun_out/pair_probe.log 2>&1 <<'PY'
import json, sys
sys.path.insert(0, 'tools')
import gpu_conv_probe as g
names = [n for n in g.CASES if n.startswith(('pair_', 'nopair_', 'tapstage_', 'respair_')) or n in ('c128_128', 'c256_256', 'c256_384', 'c768_384', 'sc_big', 'sc_pool12', 'convT22_big')]
bad = 0
for n in names:
    r = g.run_case(n, **g.CASES[n])
    ok = all((v is True) or (isinstance(v, float) and v < 6e-3) for v in r.values())
    bad += not ok
    print(n, 'ok' if ok else 'BAD', json.dumps(r), flush=True)
print('cases', len(names), 'bad', bad)
sys.exit(1 if bad else 0)
PY
rc=$?
tail -45 gpurun_out/pair_probe.log | cut -c1-200
if [ $rc -ne 0 ]; then echo "probe failed rc=$rc"; exit 0; fi
if [ -n "$TFLAGS" ]; then
  LASS_B200_LIB=$PWD/lass_b200/_lib/liblass_b200.so LASS_NO_PROFILE_RUN=1 LASS_TIMING_FLAGS=$TFLAGS timeout 600 python tools/gpu_conv_timing.py 64 $TFILTER 2>&1 | tail -30
fi
for f in "$@"; do
  LASS_CONV_FLAGS=$f timeout 300 python tools/gpu_layer_times.py 64 f$f > gpurun_out/layer_times_f$f.log 2>&1; echo "flags $f: $(tail -1 gpurun_out/layer_times_f$f.log)"
done
