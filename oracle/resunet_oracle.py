"""TEST INFRASTRUCTURE ONLY — functional fp32 CPU restatement of the separation hot path.

This is the oracle that travels to the GPU box (``/root/reference`` does not exist there).  It is
a from-scratch functional restatement — state dict in, tensors out — of:

* ``FiLM.forward``                     reference ``models/resunet.py:59-81``
* ``Base.wav_to_spectrogram_phase``    reference ``models/base.py:83-113``
* ``ResUNet30_Base.forward``           reference ``models/resunet.py:522-595``
* ``ConvBlockRes.forward``             reference ``models/resunet.py:147-165``
* ``EncoderBlockRes1B.forward``        reference ``models/resunet.py:186-198``
* ``DecoderBlockRes1B.forward``        reference ``models/resunet.py:240-264``
* ``feature_maps_to_wav``              reference ``models/resunet.py:436-519``
* ``torchlibrosa.stft.{STFT,ISTFT,magphase}``  (see ``oracle/torchlibrosa/stft.py``)

It is pinned to the unmodified reference by ``tests/test_oracle_vs_reference.py`` (run in the build
container) and by the golden fixtures in ``tests/golden`` (checked everywhere).
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LRELU_SLOPE = 0.01

# (name, cin, cout, pool)   reference models/resunet.py:315-370
ENCODERS = (
    ("encoder_block1", 32, 32, (2, 2)),
    ("encoder_block2", 32, 64, (2, 2)),
    ("encoder_block3", 64, 128, (2, 2)),
    ("encoder_block4", 128, 256, (2, 2)),
    ("encoder_block5", 256, 384, (2, 2)),
    ("encoder_block6", 384, 384, (1, 2)),
    ("conv_block7a", 384, 384, (1, 1)),
)
# (name, cin, cout, upsample)   reference models/resunet.py:371-418
DECODERS = (
    ("decoder_block1", 384, 384, (1, 2)),
    ("decoder_block2", 384, 384, (2, 2)),
    ("decoder_block3", 384, 256, (2, 2)),
    ("decoder_block4", 256, 128, (2, 2)),
    ("decoder_block5", 128, 64, (2, 2)),
    ("decoder_block6", 64, 32, (2, 2)),
)
TIME_DOWNSAMPLE = 32  # reference models/resunet.py:282


BN_MOMENTUM = 0.01  # reference models/resunet.py:275


def _bn(sd, prefix, x, train=False):
    """eval: running statistics.  train: batch statistics (biased variance) and an in-place update of
    ``sd[prefix + '.running_*']`` with momentum 0.01 / unbiased variance, as nn.BatchNorm2d does in .train()."""
    if train:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], train, BN_MOMENTUM if train else 0.0, BN_EPS)


def _film(sd, cond, name):
    """One FiLM linear; ``name`` like 'encoder_block1->conv_block1->beta1' (models/resunet.py:51-57,76)."""
    w = sd["film." + name + ".weight"]
    b = sd["film." + name + ".bias"]
    return F.linear(cond, w, b)[:, :, None, None]


def _conv_block_res(sd, prefix, film_prefix, x, cond, taps=None, train=False):
    # reference models/resunet.py:147-165
    b1 = _film(sd, cond, film_prefix + "->beta1")
    b2 = _film(sd, cond, film_prefix + "->beta2")
    a1 = F.leaky_relu(_bn(sd, prefix + ".bn1", x, train) + b1, LRELU_SLOPE)
    h = F.conv2d(a1, sd[prefix + ".conv1.weight"], None, padding=1)
    a2 = F.leaky_relu(_bn(sd, prefix + ".bn2", h, train) + b2, LRELU_SLOPE)
    h2 = F.conv2d(a2, sd[prefix + ".conv2.weight"], None, padding=1)
    if (prefix + ".shortcut.weight") in sd:
        res = F.conv2d(x, sd[prefix + ".shortcut.weight"], sd[prefix + ".shortcut.bias"])
    else:
        res = x
    out = res + h2
    if taps is not None:
        taps[prefix + ":a1"] = a1
        taps[prefix + ":a2"] = a2
        taps[prefix + ":out"] = out
    return out


def stft_mag_phase(sd, wave, n_fft, hop, eps=1e-10, prefix="base."):
    """wave (B, L) -> mag, cos, sin each (B, 1, T, F).   reference models/base.py:83-88."""
    x = wave[:, None, :]
    x = F.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    real = F.conv1d(x, sd[prefix + "stft.conv_real.weight"], stride=hop)
    imag = F.conv1d(x, sd[prefix + "stft.conv_imag.weight"], stride=hop)
    real = real[:, None].transpose(2, 3)
    imag = imag[:, None].transpose(2, 3)
    mag = torch.clamp(real ** 2 + imag ** 2, eps, math.inf) ** 0.5
    return mag, real / mag, imag / mag


def magphase(real, imag):
    """torchlibrosa.stft.magphase (different clamp from ``stft_mag_phase``)."""
    mag = (real ** 2 + imag ** 2) ** 0.5
    den = torch.clamp(mag, 1e-10, math.inf)
    return mag, real / den, imag / den


def istft(sd, real, imag, length, n_fft, hop, prefix="base."):
    """real, imag (B, 1, T, F) -> (B, length).   torchlibrosa 0.1.0 ISTFT.forward."""
    frames = real.shape[2]
    re = real[:, 0].transpose(1, 2)
    im = imag[:, 0].transpose(1, 2)
    full_re = torch.cat((re, torch.flip(re[:, 1:-1], dims=[1])), dim=1)
    full_im = torch.cat((im, -torch.flip(im[:, 1:-1], dims=[1])), dim=1)
    s = F.conv1d(full_re, sd[prefix + "istft.conv_real.weight"]) - \
        F.conv1d(full_im, sd[prefix + "istft.conv_imag.weight"])
    out_len = (frames - 1) * hop + n_fft
    y = F.fold(s, (1, out_len), (1, n_fft), stride=(1, hop))[:, 0, 0]
    wmat = sd[prefix + "istft.ola_window"][None, :, None].repeat(1, 1, frames)
    wsum = F.fold(wmat, (1, out_len), (1, n_fft), stride=(1, hop)).squeeze().clamp(1e-11, math.inf)
    y = y / wsum[None, :]
    return y[:, n_fft // 2: n_fft // 2 + length]


def mask_to_wave(sd, feat, mag, cos_in, sin_in, length, n_fft, hop, prefix="base."):
    """feat (B, 3, T, F) -> waveform (B, 1, L) for input_channels = output_channels = 1.

    reference models/resunet.py:436-519 with target_sources_num = 1, K = 3.
    """
    mask_mag = torch.sigmoid(feat[:, 0:1])
    _, mask_cos, mask_sin = magphase(torch.tanh(feat[:, 1:2]), torch.tanh(feat[:, 2:3]))
    out_cos = cos_in * mask_cos - sin_in * mask_sin
    out_sin = sin_in * mask_cos + cos_in * mask_sin
    out_mag = F.relu(mag * mask_mag)
    y = istft(sd, out_mag * out_cos, out_mag * out_sin, length, n_fft, hop, prefix)
    return y[:, None, :]


def infer_stft_params(sd, prefix="base."):
    w = sd[prefix + "stft.conv_real.weight"]
    return int(w.shape[2])


def resunet30_forward_impl(sd: Dict[str, torch.Tensor], mixture: torch.Tensor, condition: torch.Tensor,
                           hop: int = 160, taps: Optional[dict] = None, train: bool = False) -> torch.Tensor:
    """``ResUNet30.forward`` for input_channels = output_channels = 1; autograd-transparent.

    sd: reference-keyed state dict (``base.*``, ``film.*``); mixture (B, 1, L); condition (B, 512).
    Returns waveform (B, 1, L).  ``taps`` (optional dict) receives named intermediates.  ``train`` = the module in
    ``.train()`` (reference ``models/audiosep.py:99-100``): BatchNorm batch statistics, running stats in ``sd`` updated.
    """
    assert mixture.shape[1] == 1, "oracle restates the single-channel configuration (config yaml: 1/1/512)"
    n_fft = infer_stft_params(sd)
    length = mixture.shape[2]
    mag, cos_in, sin_in = stft_mag_phase(sd, mixture[:, 0], n_fft, hop)

    # bn0 runs over the frequency axis (models/resunet.py:537-539)
    # (mag.contiguous(): torch 2.11's CPU batch_norm BACKWARD returns wrong weight / bias gradients when its input and the
    #  incoming gradient have different memory formats, which happens for this (B, F, T, 1) view when `mag` keeps the
    #  F-major strides of the conv1d output while the gradient arrives (B, 1, T, F)-contiguous.  The reference's `mag` comes
    #  out of torch.cat, i.e. contiguous, and its gradients agree with fp64 finite differences; same layout here.)
    x = _bn(sd, "base.bn0", mag.contiguous().transpose(1, 3), train).transpose(1, 3)
    frames = x.shape[2]
    pad = int(math.ceil(frames / TIME_DOWNSAMPLE)) * TIME_DOWNSAMPLE - frames
    x = F.pad(x, (0, 0, 0, pad))
    x = x[..., : x.shape[-1] - 1]
    x = F.conv2d(x, sd["base.pre_conv.weight"], sd["base.pre_conv.bias"])
    if taps is not None:
        taps["mag"], taps["cos"], taps["sin"], taps["pre_conv"] = mag, cos_in, sin_in, x

    skips = []
    for name, _cin, _cout, pool in ENCODERS:
        full = _conv_block_res(sd, "base.%s.conv_block1" % name, "%s->conv_block1" % name, x, condition, taps, train)
        skips.append(full)
        x = F.avg_pool2d(full, kernel_size=pool)
    # conv_block7a's pooled output (pool 1x1 == identity) feeds the decoder; its `full` is unused
    skips.pop()

    for name, _cin, _cout, up in DECODERS:
        p = "base." + name
        b1 = _film(sd, condition, name + "->beta1")
        # NB: default negative_slope (0.01) at reference models/resunet.py:255
        a = F.leaky_relu(_bn(sd, p + ".bn1", x, train) + b1, LRELU_SLOPE)
        u = F.conv_transpose2d(a, sd[p + ".conv1.weight"], None, stride=up)
        cat = torch.cat((u, skips.pop()), dim=1)
        if taps is not None:
            taps[p + ":up"] = u
        x = _conv_block_res(sd, p + ".conv_block2", name + "->conv_block2", cat, condition, taps, train)

    feat = F.conv2d(x, sd["base.after_conv.weight"], sd["base.after_conv.bias"])
    feat = F.pad(feat, (0, 1))[:, :, :frames, :]
    if taps is not None:
        taps["feat"] = feat
    return mask_to_wave(sd, feat, mag, cos_in, sin_in, length, n_fft, hop)


@torch.no_grad()
def resunet30_forward(sd, mixture, condition, hop: int = 160, taps: Optional[dict] = None) -> torch.Tensor:
    """Eval-mode forward (reference ``dcase_evaluator.py:104``)."""
    return resunet30_forward_impl(sd, mixture, condition, hop, taps, False)


@torch.no_grad()
def chunk_inference(sd, mixture, condition, hop: int = 160, rate: int = 32000):
    """Serial restatement of ``ResUNet30.chunk_inference`` (reference models/resunet.py:655-714): 5 s windows
    (NL = 1 s, NC = 3 s, NR = 1 s at the hard-coded RATE = 32000), hop NC, numpy stitching, batch 1."""
    import numpy as np
    NL, NC, NR = int(1.0 * rate), int(3.0 * rate), int(1.0 * rate)
    L = mixture.shape[2]
    out = np.zeros([1, L])
    WINDOW = NL + NC + NR
    cur = 0
    while cur + WINDOW < L:
        chunk = resunet30_forward(sd, mixture[:, :, cur:cur + WINDOW], condition, hop=hop)[0].numpy()
        if cur == 0:
            out[:, cur:cur + WINDOW - NR] = chunk[:, :-NR] if NR != 0 else chunk
        else:
            out[:, cur + NL:cur + WINDOW - NR] = chunk[:, NL:-NR] if NR != 0 else chunk[:, NL:]
        cur += NC
        if cur < L:
            chunk = resunet30_forward(sd, mixture[:, :, cur:cur + WINDOW], condition, hop=hop)[0].numpy()
            seg_len = chunk.shape[1]
            out[:, cur + NL:cur + seg_len] = chunk[:, NL:]
    return out
