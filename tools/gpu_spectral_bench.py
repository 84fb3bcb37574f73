"""Isolated spectral round trip at BASELINE config-2 shape (64 x 10 s): K1 then K5, a few launches (ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200 import ops, packing
from lass_b200.models.spectral import STFT
n_fft, hop = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 160)
B, L = 64, 160000
stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data.cuda(), stft.conv_imag.weight.data.cuda())
window, tw = packing.istft_tables(n_fft, device="cuda")
wave = 0.1 * torch.randn(B, L, device="cuda")
T, F = L // hop + 1, n_fft // 2 + 1
feat = torch.randn(B, 3, T, F, device="cuda")
for _ in range(3):
    mag, cos, sin = ops.stft_fwd(wave, hi, lo, n_fft, hop, 0)
    out = ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
