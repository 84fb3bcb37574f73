"""Run ONE conv layer of tools/gpu_conv_timing.py a few times (for `ncu --set full` captures).
Usage: python tools/gpu_one_layer.py "<layer name>" [B] [algo]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv_saved = sys.argv[:]
name = sys.argv[1]
B = sys.argv[2] if len(sys.argv) > 2 else "16"
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sys.argv = [sys.argv[0], B, "none"]
import importlib.util
spec = importlib.util.spec_from_file_location("ct", os.path.join(ROOT, "tools", "gpu_conv_timing.py"))
ct = importlib.util.module_from_spec(spec)
ct.__dict__["__name__"] = "ct"
src = open(os.path.join(ROOT, "tools", "gpu_conv_timing.py")).read().split("\nout = {}\n")[0]
exec(compile(src, "gpu_conv_timing_head", "exec"), ct.__dict__)
ct.FLAGS = (0,)
cfg = dict(ct.LAYERS)
cfg["dec5.c2 32->32+sc64 @1024x512"] = (1024, 512, 32, 32, 64, 1, False)
cfg["enc1.c1 32->64 @512x256"] = (512, 256, 32, 64, 0, 1, False)
c = cfg[name]
if algo:
    c = tuple(c) + ((1, 1), 1)
os.environ["LASS_NO_PROFILE_RUN"] = "1"
print(name, ct.bench_layer(*c))
