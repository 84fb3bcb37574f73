"""The UNMODIFIED reference module on the same B200 (PyTorch eager / cuDNN: fp32, TF32, bf16 autocast) beside the lass_b200
forward, same weights, same 64 x 10 s batch: a same-device baseline (the bench's reference arm is the reference's CPU path).
Checks that the two agree (>= 40 dB vs the reference's fp32 output computed on the GPU) and that lass_b200 is faster than
every eager mode; writes the numbers to gpurun_out/reference_on_gpu.json (copied to profiles/ per round)."""
import json
import os

import pytest
import torch

from oracle import factory, reference_loader

from helpers import build_module, snr_ok

pytestmark = pytest.mark.gpu


def _time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def test_reference_module_on_the_same_gpu(repo_root):
    if not reference_loader.reference_available():
        pytest.skip("no reference tree (oracle/_ref is built by __graft_entry__.build() where /root/reference exists)")
    ref_mod = reference_loader.import_reference_resunet()
    model, sd = build_module(device="cuda")
    ref = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512).eval()
    ref.load_state_dict(sd)
    ref = ref.cuda()
    B, L = 64, 160000
    mix, cond = factory.make_inputs(B, L, seed=4, edge_clips=False)
    inp = {"mixture": mix.cuda(), "condition": cond.cuda()}
    res = {"batch": B, "clip_seconds": 10.0, "gpu": torch.cuda.get_device_name(0)}
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ms32, out32 = _time(lambda: ref(inp)["waveform"], 2)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_tf32, _ = _time(lambda: ref(inp)["waveform"], 3)
        torch.backends.cudnn.benchmark = True
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_bf16, out_bf16 = _time(lambda: ref(inp)["waveform"], 3)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ms_ours, out = _time(lambda: model(inp)["waveform"], 5)
    snr = snr_ok(out32.cpu(), out.cpu(), 40.0)
    snr_ac = factory.snr_db(out32.cpu(), out_bf16.float().cpu())
    for name, ms in (("reference_fp32", ms32), ("reference_tf32", ms_tf32), ("reference_bf16_autocast", ms_bf16),
                     ("lass_b200", ms_ours)):
        res[name] = {"ms_per_batch": ms, "audio_s_per_s": B * 10.0 / (ms * 1e-3)}
    res["speedup_vs_fp32"] = ms32 / ms_ours
    res["speedup_vs_tf32"] = ms_tf32 / ms_ours
    res["speedup_vs_bf16_autocast"] = ms_bf16 / ms_ours
    res["snr_db_lass_b200_vs_reference_fp32_min"] = float(snr.min())
    res["snr_db_reference_bf16_autocast_vs_fp32_min"] = float(snr_ac.min())
    res["reference_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    os.makedirs(os.path.join(repo_root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(repo_root, "gpurun_out", "reference_on_gpu.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    assert ms_ours < min(ms32, ms_tf32, ms_bf16)
