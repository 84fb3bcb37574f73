#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md), on the smallest case that covers every kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOL=${1:-memcheck}
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from lass_b200.models.resunet import ResUNet30
from oracle import factory
torch.manual_seed(0)
m = ResUNet30(1, 1, 512).eval()
m.load_state_dict(factory.fill_state_dict(m.state_dict(), seed=0))
m = m.cuda()
for L in (5157, 8000):          # ragged length (T not a multiple of 32, partial tiles) and a regular one
    mix, cond = factory.make_inputs(2, L, edge_clips=False)
    out = m({"mixture": mix.cuda(), "condition": cond.cuda()})["waveform"]
    torch.cuda.synchronize()
    print(L, float(out.abs().max()), bool(torch.isfinite(out).all()))
PY
python /tmp/san_case.py > gpurun_out/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "sanitizer exit $?"; tail -15 gpurun_out/sanitizer_$TOOL.log
