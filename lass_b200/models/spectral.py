"""Parameter holders for the spectral front/back end.

The reference keeps the (frozen) windowed DFT / IDFT matrices of ``torchlibrosa.stft.STFT`` / ``ISTFT`` as conv
parameters, so they are part of every checkpoint (``base.stft.conv_real.weight`` ... ``base.istft.ola_window``,
SURVEY.md §5).  These classes reproduce those parameters (same names, shapes, values) so state dicts load both
ways.  They do no computation themselves: the B200 path feeds ``stft.conv_real/conv_imag`` to kernel K1 as the
GEMM basis and replaces the dense IDFT + fold of ``ISTFT`` by kernel K5 (shared-memory inverse FFT).
"""
import math

import torch
import torch.nn as nn


def _hann_periodic64(n: int) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64)
    return 0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)


def _pad_center(w: torch.Tensor, size: int) -> torch.Tensor:
    lpad = (size - w.numel()) // 2
    return torch.nn.functional.pad(w, (lpad, size - w.numel() - lpad))


def _phase_matrix(n: int) -> torch.Tensor:
    """angle[x, y] = 2*pi*((x*y) mod n)/n in float64 (exact integer reduction before the trig call)."""
    idx = torch.arange(n, dtype=torch.int64)
    return (2.0 * math.pi / n) * ((idx[:, None] * idx[None, :]) % n).to(torch.float64)


class STFT(nn.Module):
    """Holds ``conv_real`` / ``conv_imag`` = Re/Im(exp(-2 pi i x y / n) * window) of shape (n/2+1, 1, n)."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", freeze_parameters=True):
        super().__init__()
        assert window == "hann" and center and pad_mode == "reflect", "only the reference's configuration"
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length
        self.hop_length = int(self.win_length // 4) if hop_length is None else hop_length
        w = _pad_center(_hann_periodic64(self.win_length), n_fft)
        ang = _phase_matrix(n_fft)[:, : n_fft // 2 + 1]          # (n, F)
        out_channels = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, bias=False)
        self.conv_real.weight.data = (torch.cos(ang) * w[:, None]).T.to(torch.float32)[:, None, :].contiguous()
        self.conv_imag.weight.data = (-torch.sin(ang) * w[:, None]).T.to(torch.float32)[:, None, :].contiguous()
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False


class ISTFT(nn.Module):
    """Holds ``conv_real`` / ``conv_imag`` = Re/Im(exp(+2 pi i x y / n) / n * window) (n, n, 1) and ``ola_window``."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", freeze_parameters=True, onnx=False, frames_num=None, device=None):
        super().__init__()
        assert window == "hann" and center and pad_mode == "reflect", "only the reference's configuration"
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length
        self.hop_length = int(self.win_length // 4) if hop_length is None else hop_length
        w = _pad_center(_hann_periodic64(self.win_length), n_fft)
        ang = _phase_matrix(n_fft)
        self.conv_real = nn.Conv1d(n_fft, n_fft, 1, bias=False)
        self.conv_imag = nn.Conv1d(n_fft, n_fft, 1, bias=False)
        self.conv_real.weight.data = (torch.cos(ang) / n_fft * w[None, :]).T.to(torch.float32)[:, :, None].contiguous()
        self.conv_imag.weight.data = (torch.sin(ang) / n_fft * w[None, :]).T.to(torch.float32)[:, :, None].contiguous()
        self.register_buffer("ola_window", (w ** 2).to(torch.float32))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False
