"""One eager (no CUDA graph) forward at batch B bracketed by cudaProfilerStart / Stop: the ncu target for small-batch launch lists.
Usage: python tools/gpu_forward_once.py [B] [L]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lass_b200.models.resunet import ResUNet30
from oracle import factory
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
torch.manual_seed(0)
m = ResUNet30(1, 1, 512).eval()
m.load_state_dict(factory.fill_state_dict(m.state_dict(), seed=0))
m = m.cuda()
mix, cond = factory.make_inputs(B, L, edge_clips=False)
mix, cond = mix.cuda(), cond.cuda()
eng = m.base._get_engine(m.film)
eng.use_graphs = False
for _ in range(3):
    m({"mixture": mix, "condition": cond})
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m({"mixture": mix, "condition": cond})["waveform"]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().max()))
