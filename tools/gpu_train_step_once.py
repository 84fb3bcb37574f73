"""One training step at BASELINE config 4 (16 x 5 s) bracketed by cudaProfilerStart / Stop (the ncu target of
tools/gpu_profile_train.sh: `ncu --profile-from-start off`).  Usage: python tools/gpu_train_step_once.py [B L]"""
import os
import sys

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
    for _ in range(2):
        eng.training_step(mix, cond, tgt, lr=1e-6)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = eng.training_step(mix, cond, tgt, lr=1e-6)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("loss", float(loss))
