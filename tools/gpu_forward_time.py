"""Device-resident throughput of the whole forward at a given STFT shape (no bench contract, just the number).
Usage: python tools/gpu_forward_time.py n_fft hop [B]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200.models.resunet import ResUNet30  # noqa: E402
from oracle import factory  # noqa: E402

n_fft, hop = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
L = 160000
torch.manual_seed(0)
m = ResUNet30(1, 1, 512, window_size=n_fft, hop_size=hop).eval()
m.load_state_dict(factory.fill_state_dict(m.state_dict(), seed=0))
m = m.cuda()
mix, cond = factory.make_inputs(B, L, edge_clips=False)
inp = {"mixture": mix.cuda(), "condition": cond.cuda()}
for _ in range(3):
    m(inp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m(inp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out = {"n_fft": n_fft, "hop": hop, "clips": B, "ms_per_batch": ms, "audio_s_per_s": B * 10.0 / (ms * 1e-3)}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "forward_time_%d_%d.json" % (n_fft, hop)), "w"))
