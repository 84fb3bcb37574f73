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
  cpu_baseline  the oracle port of the reference forward timed on this box's host cores (1 clip, best of 3)
``--impl reference`` times the reference's CPU implementation of the path (oracle port: /root/reference does not
exist on the GPU box and a Python reference cannot travel) with all host threads, one 10 s clip per step.
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


def cpu_reference_forward_time(repeats, threads=None):
    """Oracle port of the reference forward (fp32, eval, no_grad) on the host cores: seconds per 10 s clip."""
    from lass_b200.models.resunet import ResUNet30
    from oracle import resunet_oracle
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in ResUNet30(1, 1, 512).state_dict().items()}
    mix, cond = make_batch(1, 1234)
    times = []
    for i in range(repeats + 1):          # first pass is the warm-up
        t0 = time.perf_counter()
        resunet_oracle.resunet30_forward(sd, mix, cond, hop=HOP)
        times.append(time.perf_counter() - t0)
    return times[1:]


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t = cpu_reference_forward_time(args.warmup + args.steps - 1)[-(args.steps):]
    per_step = sum(t) / len(t)
    value = CLIP_SECONDS / per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResUNet30 separation forward, 10 s @ 16 kHz clips, n_fft 1024 / hop 160 "
                               "(BASELINE.json configs[2] shape; reference arm runs 1 clip per step on the CPU)",
                   "clip_seconds": CLIP_SECONDS, "sample_rate": SAMPLE_RATE, "batch_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "1 clip (10 s) per step, fp32 eval forward of the oracle port "
                                   "(oracle/resunet_oracle.py; /root/reference is not present on the GPU box)"},
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
            "mask_istft_ms": t_inv * 1e3, "mask_istft_gbs": bytes_inv / t_inv / 1e9,
            "mask_istft_frac_of_hbm": bytes_inv / t_inv / 1e9 / peaks["hbm_gbs"],
            "round_trip_audio_s_per_s": B * CLIP_SECONDS / (t_fwd + t_inv),
        }
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

    # ---------------- timed region 2: stage split with CUDA events on the launching stream ----------------
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
    unet_s = sharding.max_over_ranks(stage_ms[1] * 1e-3, device)
    n_conv_launches = launches_per_step - 4      # stft_prep, stft_gemm, film, mask_istft are the others
    achieved_tflops = unet_flops / unet_s / 1e12

    # ---------------- timed region 3: end to end through the public API with host buffers ----------------
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
    e2e_steps(min(2, args.warmup))                       # untimed: stream / allocator warm-up of the pipelined loop
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

    spectral = None
    cpu_baseline = None
    if rank == 0 and world == 1:
        if not args.no_spectral:
            spectral = time_spectral(device, peaks)
        if not args.no_cpu_baseline:
            t = cpu_reference_forward_time(3)
            best = min(t)
            cpu_baseline = {"value": CLIP_SECONDS / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": "1 clip (10 s @ 16 kHz) of the same workload, fp32 eval forward of the oracle port of "
                                      "the reference (oracle/resunet_oracle.py), 1 warm-up + best of 3",
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
            "config": {"workload": "ResUNet30 separation forward (BASELINE.json configs[2] minus the off-path CLAP encoder): "
                                   "%d clips x 10 s @ 16 kHz per GPU, n_fft %d / hop %d, random-init weights, fixed unit-norm "
                                   "512-d conditions" % (B, N_FFT, HOP),
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
                         "frac": achieved_tflops / peak, "traffic": traffic,
                         "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks["source"],
                         "algorithmic_gflop_per_clip": unet_flops / (hi - lo) / 1e9,
                         "layerwise_roofline_us_per_clip": 227.8,
                         "layerwise_frac": 227.8e-6 * (hi - lo) / unet_s},
            "cpu_baseline": cpu_baseline,
            "spectral": spectral,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
