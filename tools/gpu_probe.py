"""Round-1 GPU bring-up diagnostics (run under gpurun; writes gpurun_out/probe.json).

Not a test: prints which shared-memory descriptor hypotheses hold for tcgen05.mma and checks the first
spectral kernels against the oracle with verbose error reports, so that one GPU call answers many questions.
"""
import json
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
report = {}


WANT = sys.argv[1:]  # section names; empty = all


def section(name):
    def deco(fn):
        if WANT and name not in WANT and name != "env":
            return fn
        t0 = time.time()
        try:
            report[name] = fn()
        except Exception as exc:  # keep going: one call must answer as much as possible
            report[name] = {"error": repr(exc), "trace": traceback.format_exc()[-2000:]}
        report[name + "_s"] = round(time.time() - t0, 2)
        print("==", name, json.dumps(report[name], default=str)[:3000], flush=True)
        return fn
    return deco


@section("env")
def _env():
    p = torch.cuda.get_device_properties(0)
    return {"name": p.name, "sms": p.multi_processor_count, "mem_gb": round(p.total_memory / 2**30, 1),
            "cc": [p.major, p.minor], "torch": torch.__version__}


def _err(out, ref):
    return float((out - ref).abs().max())


RISKY = "umma_probe_risky" in WANT


@section("umma_probe_risky" if RISKY else "umma_probe")
def _probe():
    from lass_b200 import ops
    res = {}
    torch.manual_seed(0)
    dev = "cuda"
    for dt, dtn in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        for kc, sw, rowb in ((64, 2, 128), (32, 4, 64)):
            A = torch.randn(256, kc, device=dev).to(dt)
            for n in (32, 64, 128, 256):
                Bm = torch.randn(n, kc, device=dev).to(dt)
                full = A.float() @ Bm.float().t()      # (256, n)
                atom = 8 * rowb
                key = "%s_kc%d_n%d" % (dtn, kc, n)
                # E1 aligned
                out = ops.umma_probe(A, Bm, sw, 0, atom, 0, atom)
                res[key + "_aligned"] = _err(out, full[:128])
                if n != 64:
                    continue
                # E4 whole-atom shifts
                for dy in (1, 2, 5):
                    out = ops.umma_probe(A, Bm, sw, atom * dy, atom, 0, atom)
                    res[key + "_atomshift%d" % dy] = _err(out, full[8 * dy: 8 * dy + 128])
                # E5 SBO = 2 atoms (16-row pitch, 8 rows used)
                idx = torch.arange(128, device=dev)
                rows = (idx // 8) * 16 + idx % 8
                out = ops.umma_probe(A, Bm, sw, 0, 2 * atom, 0, atom)
                res[key + "_sbo2"] = _err(out, full[rows])
                # dense halo pitch: 10-row groups (SBO = 10 rows), start shifted by dx rows
                rows10 = (idx // 8) * 10 + idx % 8
                for dx in (0, 1, 2, 11, 22):
                    if int(rows10.max()) + dx < 256:
                        out = ops.umma_probe(A, Bm, sw, rowb * dx, 10 * rowb, 0, atom)
                        res[key + "_pitch10_shift%d" % dx] = _err(out, full[rows10 + dx])
                if not RISKY:
                    continue
                # E2/E3 row shifts (start address moved by dx rows) with base_offset 0 or dx
                for dx in (1, 2, 3, 7):
                    for bo in (0, dx):
                        out = ops.umma_probe(A, Bm, sw, rowb * dx, atom, bo, atom)
                        res[key + "_rowshift%d_bo%d" % (dx, bo)] = _err(out, full[dx: dx + 128])
                # row shift + 16-row pitch (halo tile): rows (g*16 + dx + r)
                for dx in (1, 2):
                    for bo in (0, dx):
                        out = ops.umma_probe(A, Bm, sw, rowb * dx, 2 * atom, bo, atom)
                        res[key + "_rowshift%d_sbo2_bo%d" % (dx, bo)] = _err(out, full[rows + dx])
    torch.cuda.synchronize()
    return res


def _spectral_setup(n_fft, hop, B, L, seed=0):
    from oracle.torchlibrosa.stft import STFT, ISTFT
    from oracle import factory
    stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=True, pad_mode="reflect")
    istft = ISTFT(n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=True, pad_mode="reflect")
    wave, _ = factory.make_inputs(B, L, seed=1234 + seed)
    return stft, istft, wave[:, 0].contiguous()


@section("stft")
def _stft():
    from lass_b200 import ops, packing
    from oracle import factory
    res = {}
    for n_fft, hop in ((1024, 160), (2048, 320), (512, 160), (256, 160)):
        B, L = 4, 32000
        stft, _, wave = _spectral_setup(n_fft, hop, B, L)
        with torch.no_grad():
            real, imag = stft(wave)
            mag_ref = torch.clamp(real ** 2 + imag ** 2, 1e-10, float("inf")) ** 0.5
            cos_ref, sin_ref = real / mag_ref, imag / mag_ref
            spec64 = torch.stft(wave.double(), n_fft, hop, n_fft, torch.hann_window(n_fft, periodic=True, dtype=torch.float64),
                                center=True, pad_mode="reflect", return_complex=True).transpose(1, 2)[:, None]
        hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data, stft.conv_imag.weight.data)
        hi, lo = hi.cuda(), lo.cuda()
        for mode in (0, 1):
            mag, cos, sin = ops.stft_fwd(wave.cuda(), hi, lo, n_fft, hop, precision_mode=mode)
            torch.cuda.synchronize()
            re_gpu = (mag * cos).cpu()
            im_gpu = (mag * sin).cpu()
            k = "n%d_h%d_m%d" % (n_fft, hop, mode)
            res[k] = {
                "mag_rel_vs_restated": factory.max_rel_err(mag_ref, mag.cpu()),
                "re_rel_vs_fp64": factory.max_rel_err(spec64.real, re_gpu),
                "im_rel_vs_fp64": factory.max_rel_err(spec64.imag, im_gpu),
                "mag_rel_vs_fp64": factory.max_rel_err(spec64.abs(), mag.cpu()),
                "cos_abs_noise_clip0": float((cos_ref[:2] - cos.cpu()[:2]).abs().max()),
                "restated_vs_fp64": factory.max_rel_err(spec64.abs(), mag_ref),
            }
    # timing at config-2 shapes
    for n_fft, hop in ((1024, 160), (2048, 320)):
        B, L = 64, 160000
        stft, _, wave = _spectral_setup(n_fft, hop, B, L)
        hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data, stft.conv_imag.weight.data)
        hi, lo, w = hi.cuda(), lo.cuda(), wave.cuda()
        ws = torch.empty(ops._cabi.load().lass_stft_workspace_bytes(B, L, n_fft, hop), dtype=torch.uint8, device="cuda")
        for mode in (0, 1):
            for _ in range(2):
                ops.stft_fwd(w, hi, lo, n_fft, hop, mode, ws)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.stft_fwd(w, hi, lo, n_fft, hop, mode, ws)
            e1.record()
            torch.cuda.synchronize()
            res["time_ms_n%d_m%d_B64" % (n_fft, mode)] = e0.elapsed_time(e1) / 5
    return res


@section("mask_istft")
def _istft():
    from lass_b200 import ops, packing
    from oracle import factory, resunet_oracle as O
    res = {}
    for n_fft, hop in ((1024, 160), (2048, 320), (512, 160)):
        B, L = 4, 32000
        stft, istft, wave = _spectral_setup(n_fft, hop, B, L)
        sd = {"base.stft.conv_real.weight": stft.conv_real.weight.data, "base.stft.conv_imag.weight": stft.conv_imag.weight.data,
              "base.istft.conv_real.weight": istft.conv_real.weight.data, "base.istft.conv_imag.weight": istft.conv_imag.weight.data,
              "base.istft.ola_window": istft.ola_window}
        with torch.no_grad():
            mag, cos, sin = [t.contiguous() for t in O.stft_mag_phase(sd, wave, n_fft, hop)]
            g = torch.Generator().manual_seed(7)
            feat = torch.randn(B, 3, mag.shape[2], mag.shape[3], generator=g)
            feat[:, :, :, -1] = 0.0   # models/resunet.py:573 zero Nyquist column
            ref = O.mask_to_wave(sd, feat, mag, cos, sin, L, n_fft, hop)[:, 0]
            # identity mask: x0 = +40 (sigmoid -> 1), x1 = +40 (tanh -> 1), x2 = 0 => waveform round trip
            feat_id = torch.zeros_like(feat)
            feat_id[:, 0] = 40.0
            feat_id[:, 1] = 40.0
            ref_id = O.mask_to_wave(sd, feat_id, mag, cos, sin, L, n_fft, hop)[:, 0]
        window, tw = packing.istft_tables(n_fft, device="cuda")
        out = ops.mask_istft(feat.cuda(), mag.cuda(), cos.cuda(), sin.cuda(), window, tw, n_fft, hop, L)
        out_id = ops.mask_istft(feat_id.cuda(), mag.cuda(), cos.cuda(), sin.cuda(), window, tw, n_fft, hop, L)
        torch.cuda.synchronize()
        k = "n%d_h%d" % (n_fft, hop)
        res[k] = {"rel_vs_oracle": factory.max_rel_err(ref, out.cpu()),
                  "rel_identity_vs_oracle": factory.max_rel_err(ref_id, out_id.cpu()),
                  "roundtrip_abs_vs_wave": float((out_id.cpu() - wave)[:2].abs().max())}
    for n_fft, hop in ((1024, 160), (2048, 320)):
        B, L = 64, 160000
        T, F = L // hop + 1, n_fft // 2 + 1
        feat = torch.randn(B, 3, T, F, device="cuda")
        mag = torch.rand(B, 1, T, F, device="cuda")
        ang = torch.rand(B, 1, T, F, device="cuda") * 6.28
        cos, sin = torch.cos(ang), torch.sin(ang)
        window, tw = packing.istft_tables(n_fft, device="cuda")
        for _ in range(2):
            ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = (6 * 4 * T * F + 4 * L) * B / 1e9
        res["time_ms_n%d_B64" % n_fft] = ms
        res["gbps_n%d_B64" % n_fft] = gb / (ms * 1e-3)
    return res


with open(os.path.join(OUT, "probe_%s.json" % ("_".join(WANT) or "all")), "w") as f:
    json.dump(report, f, indent=1, default=str)
print("probe done")
