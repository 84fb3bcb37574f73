"""Latency of ResUNet30.forward through the public module API for small batches (B200)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
model = bench.build_model(torch.device("cuda"))
res = {}
for B in (1, 2, 4, 8, 16):
    mix, cond = bench.make_batch(B, 1)
    mix, cond = mix.cuda(), cond.cuda()
    for _ in range(3):
        model({"mixture": mix, "condition": cond})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        model({"mixture": mix, "condition": cond})
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[B] = {"ms": round(ms, 3), "audio_s_per_s": round(B * 10 / (ms * 1e-3))}
    print(B, res[B], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "latency.json"), "w"))
