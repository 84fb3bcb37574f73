"""GPU bring-up of the whole forward: ResUNet30 (B200 path) vs the fp32 oracle and the bf16 rounding-point model,
with per-buffer diagnostics.  Usage: python tools/gpu_e2e_probe.py [n_fft hop] [--time]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200.models.resunet import ResUNet30  # noqa: E402
from oracle import bf16_model, factory, resunet_oracle as O  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_fft, hop = (int(args[0]), int(args[1])) if len(args) >= 2 else (1024, 160)
report = {"n_fft": n_fft, "hop": hop}

torch.manual_seed(0)
model = ResUNet30(1, 1, 512, window_size=n_fft, hop_size=hop).eval()
sd = factory.fill_state_dict(model.state_dict(), seed=0)
model.load_state_dict(sd)
B, L = 3, 24000
mix, cond = factory.make_inputs(B, L)
t0 = time.time()
taps32, taps16 = {}, {}
ref = O.resunet30_forward(sd, mix, cond, hop=hop, taps=taps32)
emu = bf16_model.forward(sd, mix, cond, hop=hop, taps=taps16)
report["cpu_oracle_s"] = round(time.time() - t0, 2)
report["snr_emulation_vs_oracle"] = factory.snr_db(ref, emu).tolist()

model = model.cuda()
out = model({"mixture": mix.cuda(), "condition": cond.cuda()})["waveform"]
torch.cuda.synchronize()
out = out.cpu()
report["snr_gpu_vs_oracle"] = factory.snr_db(ref, out).tolist()
report["snr_gpu_vs_emulation"] = factory.snr_db(emu, out).tolist()
report["out_absmax"] = float(out.abs().max())
report["finite"] = bool(torch.isfinite(out).all())

eng = model.base._get_engine(model.film)
dev = torch.device("cuda")


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def nchw(t):
    return t.permute(0, 3, 1, 2)


bufs = {}
for name in ("mag", "cos", "sin"):
    bufs[name] = rel(eng.debug_buffer(B, L, dev, name), taps32[name].contiguous())
T = taps32["mag"].shape[2]
x_raw0 = eng.debug_buffer(B, L, dev, "x_raw0")
bufs["x_raw0"] = rel(nchw(x_raw0), taps32["pre_conv"])
for k, name in enumerate(("encoder_block1", "encoder_block2", "encoder_block3", "encoder_block4", "encoder_block5",
                          "encoder_block6")):
    cat = eng.debug_buffer(B, L, dev, "cat_raw%d" % k)
    c = cat.shape[3] // 2
    bufs["skip%d" % k] = rel(nchw(cat[..., c:]), taps16["base.%s.conv_block1:out" % name])
    if k < 5:
        xr = eng.debug_buffer(B, L, dev, "x_raw%d" % (k + 1))
        pooled = torch.nn.functional.avg_pool2d(taps16["base.%s.conv_block1:out" % name], 2)
        bufs["pool%d" % k] = rel(nchw(xr), pooled)
feat = eng.debug_buffer(B, L, dev, "feat")
bufs["feat"] = rel(feat[:, :, :T, :], taps16["feat"][..., :-1])
bufs["feat_vs_fp32"] = rel(feat[:, :, :T, :], taps32["feat"][..., :-1])
report["buffers"] = bufs
report["launches"] = eng.num_launches(B, L, dev)
print(json.dumps(report, indent=1))

if "--time" in sys.argv:
    times = {}
    for Bt, Lt in ((1, 160000), (8, 160000), (64, 160000)):
        mixt, condt = factory.make_inputs(Bt, Lt, edge_clips=False)
        mixt, condt = mixt.cuda(), condt.cuda()
        for _ in range(2):
            model({"mixture": mixt, "condition": condt})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            o = model({"mixture": mixt, "condition": condt})["waveform"]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        times["B%d" % Bt] = {"ms": ms, "audio_s_per_s": Bt * Lt / 16000 / (ms * 1e-3), "finite": bool(torch.isfinite(o).all()),
                             "mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        print("time", Bt, times["B%d" % Bt], flush=True)
    report["times"] = times

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "e2e_%d.json" % n_fft), "w") as f:
    json.dump(report, f, indent=1)
