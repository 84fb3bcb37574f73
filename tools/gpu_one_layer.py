"""Run ONE conv launch of tools/gpu_conv_timing.py a few times (for `ncu --set full` captures).
Usage: python tools/gpu_one_layer.py "<layer name substring>" [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["LASS_B200_LIB"] = os.path.join(ROOT, "lass_b200", "_lib", "liblass_b200.so")
os.environ["LASS_NO_PROFILE_RUN"] = "1"
os.environ["LASS_TIMING_FLAGS"] = "0"
name = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:3]
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_conv_timing as ct
match = [k for k in ct.LAYERS if name in k]
print(match[0], ct.bench_layer(ct.LAYERS[match[0]]))
