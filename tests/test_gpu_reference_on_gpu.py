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


def test_reference_training_step_on_the_same_gpu(repo_root):
    """BASELINE config 4 (16 clips x 5 s): the unmodified reference module in .train() with autograd, l1 and
    torch.optim.AdamW(amsgrad=True) -- fp32, TF32 and bf16 autocast -- beside lass_b200's fused training step, same weights and
    inputs.  First-step losses must agree (1e-3 relative, train-mode forward of the same parameters) and the fused step must be
    faster than every eager mode; the numbers are added to gpurun_out/reference_on_gpu.json."""
    if not reference_loader.reference_available():
        pytest.skip("no reference tree (oracle/_ref is built by __graft_entry__.build() where /root/reference exists)")
    from lass_b200 import training
    ref_mod = reference_loader.import_reference_resunet()
    B, L = 16, 80000
    mix, cond = factory.make_inputs(B, L, seed=11, edge_clips=False)
    tgt, _ = factory.make_inputs(B, L, seed=12, edge_clips=False)
    mix, cond, tgt = mix.cuda(), cond.cuda(), (0.5 * tgt).cuda()
    _, sd = build_module()

    def reference_steps(mode, reps):
        torch.manual_seed(0)
        ref = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512)
        ref.load_state_dict(sd)
        ref = ref.cuda().train()
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
        tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = mode == "bf16"
        losses = []

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                out = ref({"mixture": mix, "condition": cond})["waveform"]
            loss = torch.mean(torch.abs(out.float() - tgt))
            loss.backward()
            opt.step()
            losses.append(loss.detach())
            return loss

        ms, _ = _time(step, reps)
        first = float(losses[0])
        del ref, opt
        torch.cuda.empty_cache()
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return ms, first

    ms32, loss32 = reference_steps("fp32", 2)
    ms_tf32, _ = reference_steps("tf32", 3)
    ms_bf16, loss_bf16 = reference_steps("bf16", 3)
    peak_ref = torch.cuda.max_memory_allocated() / 1e9
    model, _ = build_module(device="cuda")
    model.train()
    eng = training.TrainEngine(model)
    first = []

    def ours():
        loss = eng.training_step(mix, cond, tgt, lr=1e-6)
        first.append(loss)
        return loss

    with torch.no_grad():
        ms_ours, _ = _time(ours, 5)
    loss_ours = float(first[0])
    assert abs(loss_ours - loss32) <= 1e-3 * loss32, (loss_ours, loss32)
    res = {"batch": B, "clip_seconds": 5.0, "reference_fp32_ms": ms32, "reference_tf32_ms": ms_tf32, "reference_bf16_autocast_ms": ms_bf16,
           "lass_b200_ms": ms_ours, "speedup_vs_fp32": ms32 / ms_ours, "speedup_vs_tf32": ms_tf32 / ms_ours,
           "speedup_vs_bf16_autocast": ms_bf16 / ms_ours, "first_step_loss": {"reference_fp32": loss32, "reference_bf16_autocast": loss_bf16,
                                                                          "lass_b200": loss_ours},
           "reference_peak_mem_gb": peak_ref}
    path = os.path.join(repo_root, "gpurun_out", "reference_on_gpu.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    allres = json.load(open(path)) if os.path.exists(path) else {}
    allres["train"] = res
    with open(path, "w") as f:
        json.dump(allres, f, indent=1)
    print(json.dumps(res))
    assert ms_ours < min(ms32, ms_tf32, ms_bf16)
