"""TEST INFRASTRUCTURE ONLY — restatement of ``torchlibrosa==0.1.0`` ``torchlibrosa/stft.py``.

The reference pins ``torchlibrosa==0.1.0`` (``environment.yml:306``) but does not vendor it and
the wheel cannot be fetched offline.  This file restates the published algorithm of the three
symbols the hot path uses (``STFT``, ``ISTFT``, ``magphase``) so that the reference module runs
unmodified.  It is anchored on the reference's own call sites:

* constructor arguments    reference ``models/resunet.py:284-302``
* ``STFT.forward`` contract  reference ``models/base.py:83-88`` (returns ``(real, imag)``, each
  ``(B, 1, T, n_fft//2+1)``)
* ``ISTFT.forward`` contract reference ``models/resunet.py:510`` (``istft(real, imag, length)``)
* ``magphase`` contract      reference ``models/resunet.py:473``, ``models/base.py:147``
* state-dict keys          ``base.stft.conv_real.weight`` ``(F,1,n_fft)``, ``base.stft.conv_imag.weight``,
  ``base.istft.conv_real.weight`` ``(n_fft,n_fft,1)``, ``base.istft.conv_imag.weight``,
  buffer ``base.istft.ola_window``

Algorithm (published torchlibrosa 0.1.0): the DFT is a strided ``conv1d`` whose kernels are the
windowed DFT basis; the inverse is a 1x1 ``conv1d`` with the windowed IDFT basis followed by
``F.fold`` overlap-add and division by the folded squared window (clamped at 1e-11).
``librosa.filters.get_window('hann', n, fftbins=True)`` is the periodic Hann window
(= ``scipy.signal.get_window('hann', n, fftbins=True)``); ``librosa.util.pad_center`` centres it
in ``n_fft`` when ``win_length < n_fft``.

Cross-checked against fp64 ``torch.stft`` / ``torch.istft`` in ``tests/test_oracle_spectral.py``.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _hann_periodic(win_length: int) -> np.ndarray:
    # librosa.filters.get_window('hann', n, fftbins=True)
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def _get_window(window, win_length: int) -> np.ndarray:
    if isinstance(window, str):
        if window in ("hann", "hanning"):
            return _hann_periodic(win_length)
        import scipy.signal
        return scipy.signal.get_window(window, win_length, fftbins=True)
    return np.asarray(window, dtype=np.float64)


def _pad_center(data: np.ndarray, size: int) -> np.ndarray:
    # librosa.util.pad_center
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    if lpad < 0:
        raise ValueError("Target size ({}) must be at least input size ({})".format(size, n))
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


class DFTBase(nn.Module):
    def __init__(self):
        super(DFTBase, self).__init__()

    def dft_matrix(self, n):
        (x, y) = np.meshgrid(np.arange(n), np.arange(n))
        omega = np.exp(-2 * np.pi * 1j / n)
        W = np.power(omega, x * y)
        return W

    def idft_matrix(self, n):
        (x, y) = np.meshgrid(np.arange(n), np.arange(n))
        omega = np.exp(2 * np.pi * 1j / n)
        W = np.power(omega, x * y)
        return W


class STFT(DFTBase):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann",
                 center=True, pad_mode="reflect", freeze_parameters=True):
        super(STFT, self).__init__()
        assert pad_mode in ["constant", "reflect"]

        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.window = window
        self.center = center
        self.pad_mode = pad_mode

        if self.win_length is None:
            self.win_length = n_fft
        if self.hop_length is None:
            self.hop_length = int(self.win_length // 4)

        fft_window = _pad_center(_get_window(window, self.win_length), n_fft)

        self.W = self.dft_matrix(n_fft)
        out_channels = n_fft // 2 + 1

        self.conv_real = nn.Conv1d(in_channels=1, out_channels=out_channels, kernel_size=n_fft,
                                   stride=self.hop_length, padding=0, dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(in_channels=1, out_channels=out_channels, kernel_size=n_fft,
                                   stride=self.hop_length, padding=0, dilation=1, groups=1, bias=False)

        self.conv_real.weight.data = torch.Tensor(
            np.real(self.W[:, 0:out_channels] * fft_window[:, None]).T)[:, None, :]
        self.conv_imag.weight.data = torch.Tensor(
            np.imag(self.W[:, 0:out_channels] * fft_window[:, None]).T)[:, None, :]

        if freeze_parameters:
            for param in self.parameters():
                param.requires_grad = False

    def forward(self, input):
        """input: (batch_size, data_length) -> real, imag: (batch_size, 1, time_steps, n_fft // 2 + 1)"""
        x = input[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)
        imag = self.conv_imag(x)
        real = real[:, None, :, :].transpose(2, 3)
        imag = imag[:, None, :, :].transpose(2, 3)
        return real, imag


def magphase(real, imag):
    mag = (real ** 2 + imag ** 2) ** 0.5
    cos = real / torch.clamp(mag, 1e-10, np.inf)
    sin = imag / torch.clamp(mag, 1e-10, np.inf)
    return mag, cos, sin


class ISTFT(DFTBase):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann",
                 center=True, pad_mode="reflect", freeze_parameters=True,
                 onnx=False, frames_num=None, device=None):
        super(ISTFT, self).__init__()
        assert pad_mode in ["constant", "reflect"]

        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.window = window
        self.center = center
        self.pad_mode = pad_mode
        self.onnx = onnx

        if self.win_length is None:
            self.win_length = self.n_fft
        if self.hop_length is None:
            self.hop_length = int(self.win_length // 4)

        # (n_fft, n_fft) inverse basis, already divided by n_fft
        self.W = self.idft_matrix(n_fft) / n_fft

        self.conv_real = nn.Conv1d(in_channels=n_fft, out_channels=n_fft, kernel_size=1,
                                   stride=1, padding=0, dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(in_channels=n_fft, out_channels=n_fft, kernel_size=1,
                                   stride=1, padding=0, dilation=1, groups=1, bias=False)

        ifft_window = _pad_center(_get_window(window, self.win_length), n_fft)

        self.conv_real.weight.data = torch.Tensor(
            np.real(self.W * ifft_window[None, :]).T)[:, :, None]
        self.conv_imag.weight.data = torch.Tensor(
            np.imag(self.W * ifft_window[None, :]).T)[:, :, None]

        ola_window = torch.Tensor(ifft_window ** 2)
        self.register_buffer("ola_window", ola_window)

        if freeze_parameters:
            for param in self.parameters():
                param.requires_grad = False

    def forward(self, real_stft, imag_stft, length):
        """real_stft, imag_stft: (batch_size, 1, time_steps, n_fft // 2 + 1) -> (batch_size, length)"""
        assert real_stft.ndimension() == 4 and imag_stft.ndimension() == 4
        batch_size, _, frames_num, _ = real_stft.shape

        real_stft = real_stft[:, 0, :, :].transpose(1, 2)
        imag_stft = imag_stft[:, 0, :, :].transpose(1, 2)
        # (batch_size, n_fft // 2 + 1, time_steps)

        # Hermitian extension to the full n_fft bins
        full_real_stft = torch.cat((real_stft, torch.flip(real_stft[:, 1:-1, :], dims=[1])), dim=1)
        full_imag_stft = torch.cat((imag_stft, -torch.flip(imag_stft[:, 1:-1, :], dims=[1])), dim=1)

        # IDFT frame by frame; the synthesis window is folded into the conv weights
        s_real = self.conv_real(full_real_stft) - self.conv_imag(full_imag_stft)
        # (batch_size, n_fft, time_steps)

        # Overlap-add
        output_samples = (s_real.shape[-1] - 1) * self.hop_length + self.win_length
        y = F.fold(input=s_real, output_size=(1, output_samples),
                   kernel_size=(1, self.win_length), stride=(1, self.hop_length))
        y = y[:, 0, 0, :]

        # Overlap-add window sum
        window_matrix = self.ola_window[None, :, None].repeat(1, 1, frames_num)
        ifft_window_sum = F.fold(input=window_matrix, output_size=(1, output_samples),
                                 kernel_size=(1, self.win_length), stride=(1, self.hop_length))
        ifft_window_sum = ifft_window_sum.squeeze()
        ifft_window_sum = ifft_window_sum.clamp(1e-11, np.inf)

        y = y / ifft_window_sum[None, :]
        y = self._trim_edges(y, length)
        return y

    def _trim_edges(self, y, length):
        if length is None:
            if self.center:
                y = y[:, self.n_fft // 2: -self.n_fft // 2]
        else:
            start = self.n_fft // 2 if self.center else 0
            y = y[:, start: start + length]
        return y
