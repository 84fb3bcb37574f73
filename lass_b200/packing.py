"""Host-side packing of reference-keyed parameters into the layouts the sm_100a kernels consume.

Runs once per set of weights (off the hot path) with plain torch ops on whatever device the parameters
live on.  State-dict keys are the reference's: the frozen DFT matrices
``base.stft.conv_real.weight`` / ``conv_imag.weight`` are honoured as the forward basis.
"""
import math

import torch


def split_bf16(x: torch.Tensor):
    """x (fp32) -> (hi, lo) bf16 with hi + lo ~= x to ~16 mantissa bits."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def pack_stft_basis(conv_real_w: torch.Tensor, conv_imag_w: torch.Tensor):
    """(F, 1, n_fft) x2 -> (hi, lo) bf16 of shape (ntiles*256, n_fft), F = n_fft/2 + 1.

    Tile j holds bins [128j, 128j+128): rows [0,128) real basis, rows [128,256) imaginary basis (kernel K1,
    ``lass_b200/csrc/stft.cu``).  The imaginary basis of bin 0 is identically zero (sin 0); its row (row 128 of
    tile 0) carries the REAL basis of the Nyquist bin n_fft/2 instead, whose imaginary basis is zero as well
    (sin(pi k)), so the n_fft/2 + 1 bins fit n_fft/256 tiles exactly.
    """
    F, _, n_fft = conv_real_w.shape
    half = n_fft // 2
    assert F == half + 1 and half % 128 == 0, "K1 needs n_fft a multiple of 256 and the full half spectrum"
    ntiles = half // 128
    re = conv_real_w.reshape(F, n_fft).to(torch.float32)
    im = conv_imag_w.reshape(F, n_fft).to(torch.float32)
    full = torch.zeros(ntiles, 2, 128, n_fft, dtype=torch.float32, device=re.device)
    full[:, 0] = re[:half].reshape(ntiles, 128, n_fft)
    full[:, 1] = im[:half].reshape(ntiles, 128, n_fft)
    full[0, 1, 0] = re[half]
    hi, lo = split_bf16(full.reshape(ntiles * 256, n_fft))
    return hi.contiguous(), lo.contiguous()


def hann_periodic(n: int, device=None) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64, device=device)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(torch.float32)


def istft_tables(n_fft: int, win_length: int = None, device=None):
    """(window (n_fft) fp32, twiddle (n_fft, 2) fp32 = (cos, sin)(2 pi j / n_fft)) for kernel K5."""
    win_length = n_fft if win_length is None else win_length
    w = hann_periodic(win_length, device)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (lpad, n_fft - win_length - lpad))
    j = torch.arange(n_fft, dtype=torch.float64, device=device)
    ang = 2.0 * math.pi * j / n_fft
    tw = torch.stack((torch.cos(ang), torch.sin(ang)), dim=-1).to(torch.float32)
    return w.contiguous(), tw.contiguous()


def pack_conv_weight(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """conv2d weight (Cout, Cin, kh, kw) -> (kh*kw, Cout, Cin) K-major 16-bit (tap = ky*kw + kx)."""
    cout, cin, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).to(dtype).contiguous()


def pack_convT_weight(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """conv_transpose2d weight (Cin, Cout, sh, sw) with kernel = stride -> (1, sh*sw*Cout, Cin):
    GEMM column n = (dy*sw + dx)*Cout + co."""
    cin, cout, sh, sw = w.shape
    return w.permute(2, 3, 1, 0).reshape(1, sh * sw * cout, cin).to(dtype).contiguous()
