"""Isolated spectral round trip at BASELINE config-2 shape (64 x 10 s): K1 then K5, timed with CUDA events (also the
ncu target).  Usage: python tools/gpu_spectral_bench.py [n_fft hop]   (default: both model shapes)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200 import ops, packing
from lass_b200.models.spectral import STFT
shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(1024, 160), (2048, 320)]
B, L = 64, 160000
for n_fft, hop in shapes:
    stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
    hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data.cuda(), stft.conv_imag.weight.data.cuda())
    window, tw = packing.istft_tables(n_fft, device="cuda")
    wave = 0.1 * torch.randn(B, L, device="cuda")
    T, F = L // hop + 1, n_fft // 2 + 1
    feat = torch.randn(B, 3, T, F, device="cuda")
    for _ in range(3):
        mag, cos, sin = ops.stft_fwd(wave, hi, lo, n_fft, hop, 0)
        out = ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(10):
        mag, cos, sin = ops.stft_fwd(wave, hi, lo, n_fft, hop, 0)
    ev[1].record()
    for _ in range(10):
        out = ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
    ev[2].record()
    torch.cuda.synchronize()
    t1, t5 = ev[0].elapsed_time(ev[1]) / 10, ev[1].elapsed_time(ev[2]) / 10
    b1 = B * (4 * L + 3 * 4 * T * F) / 1e9
    b5 = B * (6 * 4 * T * F + 4 * L) / 1e9
    print("n_fft %d hop %d: stft %.3f ms (%.0f GB/s)  mask_istft %.3f ms (%.0f GB/s)  max|out| %.4f" %
          (n_fft, hop, t1, b1 / t1 * 1e3, t5, b5 / t5 * 1e3, float(out.abs().max())))
