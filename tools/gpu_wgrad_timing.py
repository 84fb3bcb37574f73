"""Per-layer timing of the weight-gradient kernels at the training step's shapes (BASELINE config 4: 16 x 5 s -> 512 x 512 grid).
Usage: python tools/gpu_wgrad_timing.py [tc|mma] [fp16|bf16]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lass_b200 import train_kernels as K

K.WGRAD_TENSOR_CORE = (sys.argv[1] if len(sys.argv) > 1 else "tc") == "tc"
xdt = torch.float16 if (sys.argv[2] if len(sys.argv) > 2 else "fp16") == "fp16" else torch.bfloat16
B = 16
LAYERS = []          # name, co, ci, taps, H, W
H = [512, 256, 128, 64, 32, 16, 16]
W = [512, 256, 128, 64, 32, 16, 8]
C = [32, 64, 128, 256, 384, 384, 384]
for k in range(5):
    c = C[k]
    LAYERS.append(("lvl%d c%d->%d 3x3" % (k, c, c), c, c, 9, H[k], W[k]))
    LAYERS.append(("lvl%d c%d->%d 3x3" % (k, 2 * c, c), c, 2 * c, 9, H[k], W[k]))
    LAYERS.append(("lvl%d c%d->%d 1x1" % (k, 2 * c, c), c, 2 * c, 1, H[k], W[k]))
res = {}
for name, co, ci, taps, h, w in LAYERS:
    dy = (torch.randn(B, h, w, co, device="cuda") * 1e-3).to(torch.bfloat16)
    x = torch.randn(B, h, w, ci, device="cuda").to(xdt)
    dw = torch.zeros(taps * co * ci, device="cuda")
    for _ in range(2):
        K.wgrad(dy, 0, co, x, 0, ci, taps, dw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        K.wgrad(dy, 0, co, x, 0, ci, taps, dw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    fl = 2.0 * B * h * w * co * ci * taps
    by = B * h * w * (co + ci) * 2
    res[name] = {"us": round(us, 1), "tflops": round(fl / us / 1e6, 1), "gbs": round(by / us / 1e3, 1)}
    print("%-24s %8.1f us %7.1f TF/s %7.1f GB/s" % (name, us, fl / us / 1e6, by / us / 1e3), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/wgrad_timing_%s_%s.json" % (sys.argv[1] if len(sys.argv) > 1 else "tc", "fp16" if xdt == torch.float16 else "bf16"), "w"), indent=1)
