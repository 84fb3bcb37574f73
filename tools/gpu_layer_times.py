"""Per-launch times of the UNet stage at the bench shape (CUDA events between launches; no profiler).
Usage: python tools/gpu_layer_times.py [B] [tag]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200.models.resunet import ResUNet30
from oracle import factory

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tag = sys.argv[2] if len(sys.argv) > 2 else "cur"
L = 160000
if os.environ.get("LASS_CONV_FLAGS"):      # debug flags of the conv kernel, read when the launches are prepared (e.g. 4096: no CTA pairs)
    from lass_b200 import _cabi
    _cabi.check(_cabi.load().lass_debug_set_conv_flags(int(os.environ["LASS_CONV_FLAGS"], 0)))
torch.manual_seed(0)
m = ResUNet30(1, 1, 512).eval()
m.load_state_dict(factory.fill_state_dict(m.state_dict(), seed=0))
m = m.cuda()
mix, cond = factory.make_inputs(B, L, edge_clips=False)
mix, cond = mix.cuda(), cond.cuda()
for _ in range(2):
    m({"mixture": mix, "condition": cond})
torch.cuda.synchronize()
eng = m.base._get_engine(m.film)
names = []
for k in range(7):
    names += ["enc%d.c1" % k, "enc%d.c2" % k]
for j in range(6):
    names += ["dec%d.up" % j, "dec%d.c1" % j, "dec%d.c2" % j]
best = None
for rep in range(3):
    r = eng.time_unet_launches(B, L, mix.device)
    best = r if best is None else [(min(a[0], b[0]), a[1]) for a, b in zip(best, r)]
tot = sum(x[0] for x in best)
# the same 32 launches back to back between two events: the difference to the sum is launch gaps / tails
out_d = torch.empty(B, 1, L, device="cuda")
stage = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.forward_stages(mix, cond, out_d, 2)
    e1.record()
    torch.cuda.synchronize()
    stage = min(stage, e0.elapsed_time(e1) / 3)
out = {}
for nme, (ms, fl) in zip(names, best):
    out[nme] = {"ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}
    print("%-9s %8.4f ms  %7.1f TFLOP/s" % (nme, ms, fl / ms / 1e9))
print("stage back to back %.3f ms" % stage)
print("total %.3f ms" % tot)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"B": B, "total_ms": tot, "layers": out}, open(os.path.join(ROOT, "gpurun_out", "layer_times_%s.json" % tag), "w"), indent=1)
