#!/usr/bin/env python
"""Benchmark of the separation hot path (BASELINE.json metric: separated audio-seconds per second, ResUNet30,
16 kHz, 1/2/4/8 B200; STFT/iSTFT HBM GB/s reported beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one ``ResUNet30.forward`` over a batch of 64 synthetic 10 s / 16 kHz clips PER GPU (weak scaling: clips
are sharded across ranks, weights replicated, no data-path collective — SURVEY.md §8e).  Random-init weights with
randomised BatchNorm statistics (no checkpoint offline), fixed unit-norm 512-d conditions (the CLAP text encoder is
off the hot path).  For N > 1 launch with ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``.

Output: ONE JSON line on rank 0 (see the keys below; contract in the task description).
  value    whole-job audio-s/s with inputs resident in HBM, timed with CUDA events, max over ranks
  e2e      same metric through the public module API with HOST (pinned) inputs: H2D + forward + D2H every step (copies
           on their own streams, double-buffered, overlapping the neighbouring steps' forwards)
  roofline the dominant kernel class (tcgen05 implicit-GEMM conv, 32 launches / step): algorithmic FLOPs of the
           UNet / CUDA-event time of the UNET stage, against the measured bf16 peak of MEASURED_PEAKS.json
  spectral BASELINE config 2 (STFT -> mask -> iSTFT round trip, 64 x 10 s): achieved HBM GB/s of K1 and K5
  cpu_baseline  the reference forward timed on this box's host cores (1 clip, best of 3): the UNMODIFIED reference when
           oracle/_ref travelled with the snapshot (kind "reference"), else the oracle port (kind "port")
  reference_on_this_gpu  a second baseline leg (N = 1): the UNMODIFIED reference module (oracle/_ref, PyTorch eager / cuDNN) on
           the same GPU with the same weights and 64 x 10 s batch, fp32 / TF32 / bf16 autocast; never the product path
  train    BASELINE config 4: one optimisation step (train-mode forward, l1_wav, backward, NCCL gradient all-reduce,
           fused AdamW-amsgrad) on 16 clips x 5 s per GPU, host inputs, loss read back; all-reduce timed alone as well
  north_star_shape / latency / strong_scaling / clap_standin  the other configurations SURVEY.md §8(d) asks for
``--impl reference`` times the reference's CPU implementation of the path with all host threads, one 10 s clip per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CLIP_SECONDS = 10.0
SAMPLE_RATE = 16000
L = int(CLIP_SECONDS * SAMPLE_RATE)
BATCH_PER_GPU = 64
N_FFT, HOP = 1024, 160            # the reference as shipped (models/resunet.py:271-272)
UNET_GFLOP_PER_CLIP = 233.35      # algorithmic conv FLOPs of one 10 s clip
METRIC = "separated audio-sec/sec (ResUNet30, 16 kHz)"
UNIT = "audio-s/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.samples:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device):
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    model = ResUNet30(input_channels=1, output_channels=1, condition_size=512, window_size=N_FFT, hop_size=HOP).eval()
    # randomised BatchNorm statistics / affine: identity BN would hide folding work
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return model.to(device)


def make_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    mix = (0.1 * torch.randn(batch, 1, L, generator=g)).clamp_(-1.0, 1.0)
    cond = torch.nn.functional.normalize(torch.randn(batch, 512, generator=g), dim=-1)
    return mix, cond


def time_reference_on_gpu(model, device):
    """A second BASELINE leg beside cpu_baseline (never the product path): the UNMODIFIED reference module (oracle/_ref, PyTorch
    eager / cuDNN) on THIS GPU with the weights of `model`, same 64 x 10 s batch -- fp32, TF32 and bf16 autocast, CUDA events, one
    warm-up + 2 timed forwards each.  {"unavailable": why} when the reference files did not travel."""
    from oracle import reference_loader
    if not reference_loader.reference_available():
        return {"unavailable": "oracle/_ref (the reference's own files, copied by __graft_entry__.build()) is not on this box"}
    try:
        ref = reference_loader.import_reference_resunet().ResUNet30(input_channels=1, output_channels=1, condition_size=512).eval()
        ref.load_state_dict(model.state_dict())
        ref = ref.to(device)
        mix, cond = make_batch(BATCH_PER_GPU, 0)
        inp = {"mixture": mix.to(device), "condition": cond.to(device)}
        out = {"batch": BATCH_PER_GPU, "note": "unmodified reference models/resunet.py, PyTorch eager / cuDNN, same weights and batch"}

        def timed():
            ref(inp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                ref(inp)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 2

        with torch.no_grad():
            for name, tf32, ac in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = ac
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    ms = timed()
                out[name] = {"ms_per_step": ms, "audio_s_per_s": BATCH_PER_GPU * CLIP_SECONDS / (ms * 1e-3)}
    except Exception as exc:                                 # a baseline leg must never take the bench line down
        out = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    finally:
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    try:
        del ref, inp
    except NameError:
        pass
    torch.cuda.empty_cache()
    return out


def cpu_reference_forward_time(repeats, threads=None):
    """The reference forward (fp32, eval, no_grad) on the host cores: (seconds per 10 s clip list, kind, description).
    kind "reference": the unmodified models/resunet.py of the reference (oracle/_ref, see oracle/build_ref.py) over the
    restated torchlibrosa; kind "port": the functional restatement oracle/resunet_oracle.py (same aten ops, same order)."""
    from lass_b200.models.resunet import ResUNet30
    from oracle import reference_loader, resunet_oracle
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in ResUNet30(1, 1, 512).state_dict().items()}
    mix, cond = make_batch(1, 1234)
    kind, what, fn = "port", "oracle port of the reference (oracle/resunet_oracle.py)", None
    if reference_loader.reference_available():
        try:
            ref_mod = reference_loader.import_reference_resunet()
            ref = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512).eval()
            ref.load_state_dict(sd)
            kind, what = "reference", "unmodified reference models/resunet.py (%s) over the restated torchlibrosa" % (
                os.path.relpath(reference_loader.REFERENCE_ROOT, ROOT) if reference_loader.REFERENCE_ROOT.startswith(ROOT)
                else reference_loader.REFERENCE_ROOT)

            def fn():
                with torch.no_grad():
                    return ref({"mixture": mix, "condition": cond})["waveform"]
        except Exception as exc:                         # noqa: BLE001 - any import problem falls back to the port, and says so
            what += " (reference import failed: %s)" % repr(exc)[:120]
    if fn is None:
        def fn():
            return resunet_oracle.resunet30_forward(sd, mix, cond, hop=HOP)
    times = []
    for i in range(repeats + 1):          # first pass is the warm-up
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times[1:], kind, what


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t, kind, what = cpu_reference_forward_time(args.warmup + args.steps - 1)
    t = t[-(args.steps):]
    per_step = sum(t) / len(t)
    value = CLIP_SECONDS / per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResUNet30 separation forward, 10 s @ 16 kHz clips, n_fft 1024 / hop 160 "
                               "(BASELINE.json configs[2] shape; reference arm runs 1 clip per step on the CPU)",
                   "clip_seconds": CLIP_SECONDS, "sample_rate": SAMPLE_RATE, "batch_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": "1 clip (10 s) per step, fp32 eval forward: " + what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def time_spectral(device, peaks):
    """BASELINE config 2: STFT -> mag/cos/sin -> masked iSTFT, 64 x 10 s, isolated (K1 and K5 timed separately)."""
    from lass_b200 import ops, packing
    from lass_b200.models.spectral import STFT
    out = {}
    for n_fft, hop in ((1024, 160), (2048, 320)):
        stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
        hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data.to(device), stft.conv_imag.weight.data.to(device))
        window, tw = packing.istft_tables(n_fft, device=device)
        B = BATCH_PER_GPU
        wave = (0.1 * torch.randn(B, L, device=device))
        T, F = L // hop + 1, n_fft // 2 + 1
        feat = torch.randn(B, 3, T, F, device=device)
        ws = torch.empty(ops._cabi.load().lass_stft_workspace_bytes(B, L, n_fft, hop), dtype=torch.uint8, device=device)
        mag, cos, sin = ops.stft_fwd(wave, hi, lo, n_fft, hop, 0, ws)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(3):
            ops.stft_fwd(wave, hi, lo, n_fft, hop, 0, ws)
            ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
        reps = 5
        ev[0].record()
        for _ in range(reps):
            ops.stft_fwd(wave, hi, lo, n_fft, hop, 0, ws)
        ev[1].record()
        for _ in range(reps):
            ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
        ev[2].record()
        torch.cuda.synchronize()
        t_fwd = ev[0].elapsed_time(ev[1]) / reps * 1e-3
        t_inv = ev[1].elapsed_time(ev[2]) / reps * 1e-3
        bytes_fwd = B * (4 * L + 3 * 4 * T * F)            # front 4L + 3*4*T*F per clip
        bytes_inv = B * (6 * 4 * T * F + 4 * L)            # back 6*4*T*F + 4L per clip
        out["%d/%d" % (n_fft, hop)] = {
            "stft_ms": t_fwd * 1e3, "stft_gbs": bytes_fwd / t_fwd / 1e9, "stft_frac_of_hbm": bytes_fwd / t_fwd / 1e9 / peaks["hbm_gbs"],
            "stft_tflops_3pass": 3 * 2.0 * T * n_fft * 2 * F * B / t_fwd / 1e12,
            # K1 is a tensor-core kernel by construction (DFT as a 3-pass bf16-split GEMM): its roof is the bf16 peak, not HBM
            "stft_frac_of_bf16_burst": 3 * 2.0 * T * n_fft * 2 * F * B / t_fwd / 1e12 / peaks.get("bf16_tflops", 1662.4),
            "mask_istft_ms": t_inv * 1e3, "mask_istft_gbs": bytes_inv / t_inv / 1e9,
            "mask_istft_frac_of_hbm": bytes_inv / t_inv / 1e9 / peaks["hbm_gbs"],
            "round_trip_audio_s_per_s": B * CLIP_SECONDS / (t_fwd + t_inv),
        }
    return out



# ---------------------------------------------------------------------------------------------------------------------
# Extra configurations of SURVEY.md §8(d) carried in the same JSON line
# ---------------------------------------------------------------------------------------------------------------------
TRAIN_CLIPS_PER_GPU = 16          # BASELINE config 4: 16 x 5 s mixtures per rank
TRAIN_SAMPLES = 80000


def clap_conditions(batch, device):
    """BASELINE config 3 'as stated': the 512-d conditions come from the random-init CLAP text stand-in (RoBERTa-base
    geometry, 124.6 M parameters, 512-token captions; PyTorch, off the hot path) — computed once, outside the timed hot path,
    and its time reported.  Returns (cond (batch, 512) on device, info dict)."""
    from lass_b200.models.clap_standin import RandomInitCLAPTextEncoder
    t0 = time.perf_counter()
    enc = RandomInitCLAPTextEncoder().to(device)
    build_s = time.perf_counter() - t0
    words = ("dog", "rain", "engine", "speech", "piano", "siren", "wind", "crowd", "bird", "door", "water", "drum")
    captions = ["the sound of a %s and a %s number %d" % (words[i % 12], words[(i * 5 + 3) % 12], i) for i in range(batch)]
    enc.get_query_embed(modality="text", text=captions[:2])          # warm-up (cuDNN / cuBLAS handles), then drop the cache
    enc._cache.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cond = enc.get_query_embed(modality="text", text=captions)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    info = {"ms_for_batch": ms, "captions": batch, "tokens_per_caption": 512, "build_s": build_s,
            "note": "random-init RoBERTa-base text tower + projection in PyTorch fp32, one pass per distinct caption, "
                    "outside the timed hot path"}
    del enc
    torch.cuda.empty_cache()
    return cond.contiguous(), info


def time_forward(model, batch, length, steps, warmup, device, sync_each=False):
    """audio-s/s of `steps` module forwards on device-resident inputs (CUDA events); sync_each = latency mode."""
    g = torch.Generator().manual_seed(99)
    mix = (0.1 * torch.randn(batch, 1, length, generator=g)).to(device)
    cond = torch.nn.functional.normalize(torch.randn(batch, 512, generator=g), dim=-1).to(device)
    inp = {"mixture": mix, "condition": cond}
    for _ in range(warmup):
        model(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        out = model(inp)["waveform"]
        if sync_each:
            out[0, 0, :1].cpu()                         # the caller waits for every result (what dcase_evaluator.py does)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return e0.elapsed_time(e1) * 1e-3 / steps, wall / steps


def time_segment_mixer(device, B, Ls, peaks):
    """The step in front of the training path (reference data/waveform_mixers.py:9-62, called at models/audiosep.py:76-78):
    lass_b200.data.waveform_mixers.SegmentMixer (C-ABI lass_segment_mix, two launches) on the training batch, device time per
    call, beside the UNMODIFIED reference module on this GPU (PyTorch eager, a Python loop over the batch) when oracle/_ref
    travels with the snapshot."""
    import random
    from lass_b200.data.waveform_mixers import SegmentMixer
    g = torch.Generator().manual_seed(99)
    wave = (0.1 * torch.randn(B, 1, Ls, generator=g)).to(device)
    mixer = SegmentMixer(max_mix_num=2, lower_db=-10, higher_db=10)      # config/audiosep_base.yaml: max_mix_num 2, -10 .. 10 dB
    random.seed(0)
    for _ in range(5):
        mixer(waveforms=wave)
    torch.cuda.synchronize()
    n = 100
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        mixer(waveforms=wave)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / n
    nbytes = 3 * B * Ls * 4                                             # read the batch once, write mixture and segment
    # the two launches alone (same draws every call, no host-side work between them): lass_segment_mix through ctypes
    from lass_b200 import _cabi
    from lass_b200.data.waveform_mixers import draw_plan
    lib = _cabi.load()
    plan = torch.from_numpy(draw_plan(B, 2, -10, 10)).to(device)
    scratch = torch.empty(lib.lass_segment_mix_scratch_bytes(B) // 4, dtype=torch.float32, device=device)
    flat, out_m, out_s = wave.reshape(B, Ls), torch.empty(B, Ls, device=device), torch.empty(B, Ls, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def launch():
        _cabi.check(lib.lass_segment_mix(flat.data_ptr(), B, Ls, 2, plan.data_ptr(), out_m.data_ptr(), out_s.data_ptr(),
                                         scratch.data_ptr(), scratch.numel() * 4, stream))
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        launch()
    b.record()
    torch.cuda.synchronize()
    us_kernels = a.elapsed_time(b) * 1e3 / n
    res = {"us_per_call": us, "us_per_call_kernels_back_to_back": us_kernels, "launches_per_call": 2, "algorithmic_bytes": nbytes,
           "gbs": nbytes / us_kernels / 1e3, "frac_of_hbm": nbytes / us_kernels / 1e3 / peaks["hbm_gbs"],
           "note": "15 MB per call (L2-resident): launch-bound, not HBM-bound; us_per_call is through the Python mirror with the "
                   "host-side random draws of every call, us_per_call_kernels_back_to_back the C-ABI call alone"}
    try:
        from oracle import reference_loader                             # baseline leg only (the unmodified reference)
        if reference_loader.mixers_available():
            ref = reference_loader.import_reference_mixers().SegmentMixer(max_mix_num=2, lower_db=-10, higher_db=10)
            for _ in range(2):
                ref(wave)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                ref(wave)
            torch.cuda.synchronize()
            res["reference_on_this_gpu_us"] = (time.perf_counter() - t0) * 1e6 / 5
    except Exception as exc:                                            # the baseline leg must never break the bench line
        res["reference_on_this_gpu_us"] = None
        res["reference_error"] = repr(exc)[:200]
    return res


def time_train(device, rank, world, steps, warmup, peaks):
    """BASELINE config 4 through TrainEngine.training_step: per step H2D of mixture / condition / target from pinned host
    memory, train-mode forward, l1_wav, backward, NCCL all-reduce of the flat gradient (three buckets, overlapped with the
    encoder's backward), fused AdamW-amsgrad + weight re-pack, loss copied back to the host."""
    import torch.distributed as dist
    from lass_b200 import sharding, train_kernels, training
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    model = ResUNet30(input_channels=1, output_channels=1, condition_size=512, window_size=N_FFT, hop_size=HOP).to(device).train()
    eng = training.TrainEngine(model)
    B, Ls = TRAIN_CLIPS_PER_GPU, TRAIN_SAMPLES
    g = torch.Generator().manual_seed(4321 + rank)
    mix_h = (0.1 * torch.randn(B, 1, Ls, generator=g)).pin_memory()
    tgt_h = (0.05 * torch.randn(B, 1, Ls, generator=g)).pin_memory()
    cond_h = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).pin_memory()
    loss_h = torch.zeros(steps + warmup, dtype=torch.float32).pin_memory()

    def step(i):
        mix, tgt, cond = (t.to(device, non_blocking=True) for t in (mix_h, tgt_h, cond_h))
        with torch.no_grad():
            loss = eng.training_step(mix, cond, tgt, lr=1e-5)
        loss_h[i:i + 1].copy_(loss.reshape(1), non_blocking=True)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    c0 = train_kernels.CALLS
    step(0)
    calls_per_step = train_kernels.CALLS - c0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_s = sharding.max_over_ranks(e0.elapsed_time(e1) * 1e-3 / steps, device)
    # phases without the collective (this rank), CUDA events
    mix, tgt, cond = (t.to(device) for t in (mix_h, tgt_h, cond_h))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ph = [0.0, 0.0, 0.0]
    reps = 3
    with torch.no_grad():
        for _ in range(reps):
            ev[0].record()
            eng.forward(mix, cond)
            ev[1].record()
            ws = eng._last
            ws.loss_sum.zero_()
            eng.k.l1_loss(ws.wave, tgt.reshape(B, Ls), ws.dwave, ws.loss_sum)
            eng.backward(ws.dwave)
            ev[2].record()
            eng.optimizer_step(1e-5)
            ev[3].record()
            torch.cuda.synchronize()
            for j in range(3):
                ph[j] += ev[j].elapsed_time(ev[j + 1]) / reps
    out = {"workload": "AudioSep training step (BASELINE.json configs[3]): %d clips x 5 s @ 16 kHz per GPU, n_fft %d / hop %d, "
                       "train-mode BatchNorm (batch statistics), l1_wav, backward, AdamW(amsgrad)" % (B, N_FFT, HOP),
           "clips_per_gpu": B, "global_clips": B * world, "ms_per_step": step_s * 1e3, "steps_per_s": 1.0 / step_s,
           "audio_s_per_s": B * world * Ls / SAMPLE_RATE / step_s, "steps": steps,
           "phases_ms_no_collective": {"forward": ph[0], "loss_backward": ph[1], "adamw_repack": ph[2]},
           "cabi_calls_per_step": calls_per_step, "loss_first": float(loss_h[0]), "loss_last": float(loss_h[steps + warmup - 1]),
           "h2d_bytes_per_step": (mix_h.numel() + tgt_h.numel() + cond_h.numel()) * 4, "d2h_bytes_per_step": 4,
           "grad_elements": int(eng.live_end), "storage": "forward tensors fp16, gradients bf16, parameters / optimizer state / "
                                                          "statistics fp32",
           "mem_gb": torch.cuda.max_memory_allocated(device) / 1e9}
    out["segment_mixer"] = time_segment_mixer(device, B, Ls, peaks)
    # algorithmic conv FLOPs of the step: forward + dgrad + wgrad = 3 x the UNet forward of 16 x 5 s
    flops = 3.0 * UNET_GFLOP_PER_CLIP * 1e9 * (B * Ls / float(L))
    out["conv_tflops_algorithmic"] = flops / step_s / 1e12
    out["frac_of_bf16_sustained"] = flops / step_s / 1e12 / peaks["bf16_tflops_sustained"]
    if world > 1:
        # the reference's own DDP configuration: sync_batchnorm True (config/audiosep_base.yaml:42) = one small all-reduce per
        # BatchNorm site in the forward (33) and in the backward (32) on top of the gradient buckets
        eng.sync_batchnorm = True
        for i in range(2):
            step(i)
        dist.barrier()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(steps):
            step(warmup + i)
        s1.record()
        dist.barrier()
        torch.cuda.synchronize()
        sync_s = sharding.max_over_ranks(s0.elapsed_time(s1) * 1e-3 / steps, device)
        out["sync_batchnorm"] = {"ms_per_step": sync_s * 1e3, "steps_per_s": 1.0 / sync_s, "small_allreduces_per_step": 65,
                                 "note": "BatchNorm statistics over all ranks (torch.nn.SyncBatchNorm semantics), NCCL all-reduce "
                                         "of (2, C) fp64 sums per site; the headline ms_per_step above is with per-rank statistics"}
        # the same with the exchange fused into the finalize kernels over NVLink peer memory (no NCCL call per site)
        try:
            eng.sync_transport = "p2p"
            for i in range(2):
                step(i)
            dist.barrier()
            torch.cuda.synchronize()
            s0.record()
            for i in range(steps):
                step(warmup + i)
            s1.record()
            dist.barrier()
            torch.cuda.synchronize()
            p2p_s = sharding.max_over_ranks(s0.elapsed_time(s1) * 1e-3 / steps, device)
            eng.check_sync_status()
            out["sync_batchnorm"]["peer_memory"] = {
                "ms_per_step": p2p_s * 1e3, "steps_per_s": 1.0 / p2p_s, "collective_calls_per_step_for_statistics": 0,
                "note": "lass_bn_finalize_p2p / lass_bn_bwd_finalize_p2p: every rank's sums read through NVLink peer pointers "
                        "(torch symmetric memory) inside the finalize kernel, epoch flags instead of a collective"}
        except Exception as exc:          # e.g. symmetric memory unavailable on this box: report, keep the line
            out["sync_batchnorm"]["peer_memory"] = {"unavailable": repr(exc)[:300]}
        eng.sync_transport = "nccl"
        eng.sync_batchnorm = False
        n = int(eng.live_end)
        for _ in range(2):
            dist.all_reduce(eng.G[:n])
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(eng.G[:n])
        a1.record()
        torch.cuda.synchronize()
        ar_s = sharding.max_over_ranks(a0.elapsed_time(a1) * 1e-3 / 5, device)
        out["allreduce"] = {"bytes": n * 4, "ms_alone": ar_s * 1e3, "algbw_gbs": n * 4 / ar_s / 1e9,
                            "busbw_gbs": n * 4 / ar_s / 1e9 * 2.0 * (world - 1) / world,
                            "exposed_ms_in_step": max(0.0, step_s * 1e3 - sum(ph)),
                            "note": "NCCL sum over NVLink of the flat fp32 gradient (three buckets in the step: the decoder and deep-encoder buckets "
                                    "overlaps the encoder backward); 1/world folded into the AdamW kernel"}
    del eng, model
    torch.cuda.empty_cache()
    return out


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to fd 1 when NCCL_DEBUG is
# set in the environment), so the process keeps a private copy of the real stdout for the result line and points fd 1 at stderr.
_RESULT_FD = None


def _claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spectral", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip train / north-star shape / latency / strong scaling / CLAP")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from lass_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (lass_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    peaks = load_peaks()
    model = build_model(device)
    B = args.batch
    n_clips_total = B * world
    lo, hi = sharding.shard_bounds(n_clips_total, rank, world)       # this rank's slab of the global batch
    mix_h, cond_h = make_batch(hi - lo, 1234 + rank)
    clap = None
    if not args.no_extras:
        cond_d0, clap = clap_conditions(hi - lo, device)          # config 3 as stated: conditions from the CLAP stand-in
        cond_h = cond_d0.cpu()
        del cond_d0
    mix_h, cond_h = mix_h.pin_memory(), cond_h.pin_memory()
    mix_d, cond_d = mix_h.to(device), cond_h.to(device)
    out_h = torch.empty(hi - lo, 1, L, dtype=torch.float32).pin_memory()
    inputs = {"mixture": mix_d, "condition": cond_d}
    engine = model.base._get_engine(model.film)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (also builds the plan / packs the weights) ----------------
    for _ in range(args.warmup):
        out_d = model(inputs)["waveform"]
    launches_per_step = engine.num_launches(hi - lo, L, device)
    unet_flops = engine.unet_flops(hi - lo, L, device)

    # ---------------- timed region 1: device-resident inputs (value) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_s = time.time()
    while not sampler.samples and time.time() - t_s < 3.0:     # nvidia-smi needs a moment before its first sample
        time.sleep(0.05)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out_d = model(inputs)["waveform"]
    e1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_s = sharding.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device)
    value = n_clips_total * CLIP_SECONDS * args.steps / dev_s

    # ---------------- timed region 2: end to end through the public API with host buffers ----------------
    # (directly after region 1, i.e. in the same thermal / power state as `value`; the >= 2 s roofline loop comes afterwards)
    # Every step copies ITS inputs from pinned host memory, calls the module, and copies ITS result back to pinned host
    # memory.  The copies run on their own streams with double-buffered device tensors, so the H2D of step i + 1 and the D2H
    # of step i - 1 overlap the forward of step i (what a serving loop does); the timed region ends when the last result
    # has arrived in host memory.
    s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    cur = torch.cuda.current_stream(device)
    mix_b = [torch.empty_like(mix_d) for _ in range(2)]
    cond_b = [torch.empty_like(cond_d) for _ in range(2)]
    out_hb = [torch.empty_like(out_h).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    def e2e_steps(n):
        for i in range(n):
            k = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_used[k])          # the forward of step i - 2 has consumed this input buffer
                mix_b[k].copy_(mix_h, non_blocking=True)
                cond_b[k].copy_(cond_h, non_blocking=True)
                ev_in[k].record(s_in)
            cur.wait_event(ev_in[k])
            w = model({"mixture": mix_b[k], "condition": cond_b[k]})["waveform"]
            ev_used[k].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_used[k])
                if i >= 2:
                    ev_out[k].synchronize()              # host side: the previous result in this pinned buffer has landed
                out_hb[k].copy_(w, non_blocking=True)
                w.record_stream(s_out)
                ev_out[k].record(s_out)
        cur.wait_stream(s_out)

    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    e2e_steps(max(args.warmup, 1))                       # untimed: W warm-up steps of the pipelined loop (the caching allocator needs
                                                         # a few rounds until the record_stream'd result blocks recycle without a cudaMalloc)
    torch.cuda.synchronize()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_steps(args.steps)
    e3.record()
    barrier()
    e2e_s = sharding.max_over_ranks(e2.elapsed_time(e3) * 1e-3, device)
    e2e_value = n_clips_total * CLIP_SECONDS * args.steps / e2e_s
    out_h.copy_(out_hb[(args.steps - 1) & 1])
    h2d = mix_h.numel() * 4 + cond_h.numel() * 4
    d2h = out_h.numel() * 4

    # ---------------- timed region 3: stage split with CUDA events on the launching stream ----------------
    stage_ms = [0.0, 0.0, 0.0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    for _ in range(args.steps):
        ev[0].record()
        engine.forward_stages(mix_d, cond_d, out_d, 1)
        ev[1].record()
        engine.forward_stages(mix_d, cond_d, out_d, 2)
        ev[2].record()
        engine.forward_stages(mix_d, cond_d, out_d, 4)
        ev[3].record()
        torch.cuda.synchronize()
        for i in range(3):
            stage_ms[i] += ev[i].elapsed_time(ev[i + 1]) / args.steps
    # roofline region: the UNet stage alone, looped for >= 2 s so that "sustained" clocks apply (a 0.4 s region runs at burst
    # clocks and would flatter the fraction); clocks sampled over exactly this region
    n_roof = max(args.steps, int(2.0 / max(stage_ms[1] * 1e-3, 1e-4)) + 1)
    sampler_r = ClockSampler(local_rank)
    sampler_r.start()
    t_s = time.time()
    while not sampler_r.samples and time.time() - t_s < 3.0:
        time.sleep(0.05)
    barrier()
    t_r0 = time.time()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(n_roof):
        engine.forward_stages(mix_d, cond_d, out_d, 2)
    r1.record()
    barrier()
    t_r1 = time.time()
    clocks_roof = sampler_r.stop(t_r0, t_r1)
    unet_s = sharding.max_over_ranks(r0.elapsed_time(r1) * 1e-3 / n_roof, device)
    n_conv_launches = launches_per_step - 4      # stft_prep, stft_gemm, film, mask_istft are the others
    achieved_tflops = unet_flops / unet_s / 1e12

    # ---------------- the other configurations (every rank takes part where ranks matter) ----------------
    extras = {}
    if not args.no_extras:
        if world > 1 and BATCH_PER_GPU % world == 0:
            # strong scaling: the SAME 64 global clips split over the ranks
            bs = BATCH_PER_GPU // world
            dev_t, _ = time_forward(model, bs, L, args.steps, args.warmup, device)
            barrier()
            dev_t = sharding.max_over_ranks(dev_t, device)
            extras["strong_scaling"] = {"global_clips": BATCH_PER_GPU, "clips_per_gpu": bs, "ms_per_step": dev_t * 1e3,
                                        "audio_s_per_s": BATCH_PER_GPU * CLIP_SECONDS / dev_t}
        extras["train"] = time_train(device, rank, world, args.steps, args.warmup, peaks)

    spectral = None
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_extras:
        # latency (the reference's shipped driver runs batch 1, dcase_evaluator.py:99-104) and small batches, through the module
        lat = {}
        for bsz in (1, 8):
            dev_t, wall_t = time_forward(model, bsz, L, 50 if bsz == 1 else 20, 5, device, sync_each=True)
            lat["batch%d" % bsz] = {"ms_per_call_wall": wall_t * 1e3, "ms_per_call_device": dev_t * 1e3,
                                    "audio_s_per_s": bsz * CLIP_SECONDS / wall_t}
        g = torch.Generator().manual_seed(7)
        long_mix = (0.1 * torch.randn(1, 1, 60 * SAMPLE_RATE, generator=g)).to(device)
        long_in = {"mixture": long_mix, "condition": cond_d[:1]}
        model.chunk_inference(long_in, rate=SAMPLE_RATE)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.chunk_inference(long_in, rate=SAMPLE_RATE)
        lat["chunk_inference_60s"] = {"ms": (time.perf_counter() - t0) * 1e3,
                                      "note": "60 s clip, 5 s windows hopping 3 s stacked on the batch axis, result as numpy"}
        extras["latency"] = lat
        # the north_star's STFT shape (n_fft 2048 / hop 320): whole forward, 64 x 10 s
        from lass_b200.models.resunet import ResUNet30
        engine.release()                                   # plans + workspace of the 1024 / 160 model (rebuilt on next use)
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        m2 = ResUNet30(input_channels=1, output_channels=1, condition_size=512, window_size=2048, hop_size=320).eval().to(device)
        dev_t, _ = time_forward(m2, BATCH_PER_GPU, L, args.steps, args.warmup, device)
        extras["north_star_shape"] = {"n_fft": 2048, "hop": 320, "clips": BATCH_PER_GPU, "ms_per_step": dev_t * 1e3,
                                      "audio_s_per_s": BATCH_PER_GPU * CLIP_SECONDS / dev_t}
        del m2
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras and not args.no_cpu_baseline:
        extras["reference_on_this_gpu"] = time_reference_on_gpu(model, device)
    if rank == 0 and world == 1:
        if not args.no_spectral:
            spectral = time_spectral(device, peaks)
        if not args.no_cpu_baseline:
            t, kind, what = cpu_reference_forward_time(3)
            best = min(t)
            cpu_baseline = {"value": CLIP_SECONDS / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                            "sample": "1 clip (10 s @ 16 kHz) of the same workload, fp32 eval forward, 1 warm-up + best of 3: "
                                      + what,
                            "seconds_per_clip": best}

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("conv_unet_dram_bytes_per_step")
        except Exception:
            traffic = None

    if rank == 0:
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ResUNet30 separation forward (BASELINE.json configs[2]): "
                                   "%d clips x 10 s @ 16 kHz per GPU, n_fft %d / hop %d, random-init weights, unit-norm 512-d "
                                   "conditions %s" % (B, N_FFT, HOP, "from the random-init CLAP text stand-in (computed once, "
                                                      "outside the timed region)" if clap else "(fixed random)"),
                       "clips_per_gpu": B, "global_clips": n_clips_total, "clip_seconds": CLIP_SECONDS,
                       "sample_rate": SAMPLE_RATE, "n_fft": N_FFT, "hop": HOP, "sharding": "clips across ranks, no collective",
                       "l2": "inputs_exceed_l2 (per-step activations are GBs, L2 is 126 MB)",
                       "storage": "activations/weights bf16, raw residual stream fp16, fp32 accumulate + epilogues"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches_per_step * args.steps * world,
            "launches_per_step_per_gpu": launches_per_step,
            "clocks": clocks,
            "stage_ms": {"front_stft_film": stage_ms[0], "unet_convs": stage_ms[1], "mask_istft": stage_ms[2]},
            "roofline": {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM, %d launches per step)" % n_conv_launches,
                         "bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak, "frac_burst": achieved_tflops / peaks["bf16_tflops"],
                         "peak_burst": peaks["bf16_tflops"], "traffic": traffic,
                         "traffic_source": "offline constant: ncu --set full capture of the 32 conv launches "
                                           "(profiles/traffic.json), NOT measured in this run",
                         "peak_source": "%s bf16_tflops_sustained (stage looped %d x = %.1f s, clocks below)" % (
                             peaks["source"], n_roof, unet_s * n_roof),
                         "region_clocks": clocks_roof,
                         "algorithmic_gflop_per_clip": unet_flops / (hi - lo) / 1e9,
                         "layerwise_roofline_us_per_clip": 227.8,
                         "layerwise_frac": 227.8e-6 * (hi - lo) / unet_s},
            "cpu_baseline": cpu_baseline,
            "spectral": spectral,
            "clap_standin": clap,
        }
        line.update(extras)
        if "train" in extras:
            line["gpu_launches_train_calls_per_step"] = extras["train"]["cabi_calls_per_step"]
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
