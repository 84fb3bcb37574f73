"""Training-step timing at BASELINE config 4 (16 clips x 5 s per GPU): phases with CUDA events, per-kernel-kind totals via
torch.profiler (kernel names only; no ncu needed)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lass_b200 import training
from lass_b200.models.resunet import ResUNet30

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 80000
torch.manual_seed(0)
model = ResUNet30(1, 1, 512).cuda().train()
eng = training.TrainEngine(model)
g = torch.Generator().manual_seed(1)
mix = (0.1 * torch.randn(B, 1, L, generator=g)).cuda()
tgt = (0.05 * torch.randn(B, 1, L, generator=g)).cuda()
cond = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).cuda()
with torch.no_grad():
    for _ in range(3):
        eng.training_step(mix, cond, tgt, lr=1e-6)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    reps = 5
    t = [0.0] * 4
    for _ in range(reps):
        ev[0].record()
        wave = eng.forward(mix, cond)
        ev[1].record()
        ws = eng._last
        ws.loss_sum.zero_()
        eng.k.l1_loss(ws.wave, tgt.reshape(B, L), ws.dwave, ws.loss_sum)
        eng.backward(ws.dwave)
        ev[2].record()
        eng.optimizer_step(1e-6)
        ev[3].record()
        torch.cuda.synchronize()
        for i in range(3):
            t[i] += ev[i].elapsed_time(ev[i + 1]) / reps
    out = {"B": B, "L": L, "forward_ms": t[0], "loss_backward_ms": t[1], "optimizer_repack_ms": t[2], "step_ms": sum(t[:3]),
           "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.training_step(mix, cond, tgt, lr=1e-6)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        if e.device_time_total > 0:
            rows.append((e.device_time_total / 1e3, e.count, e.key[:90]))
    rows.sort(reverse=True)
    out["kernels_ms_count_name"] = rows[:25]
    out["kernel_total_ms"] = sum(r[0] for r in rows)
print(json.dumps(out, indent=1))
