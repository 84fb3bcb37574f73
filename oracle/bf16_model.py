"""TEST INFRASTRUCTURE ONLY — CPU model of the B200 path's *rounding points*.

The CUDA path stores activations as NHWC bf16, multiplies bf16 operands on tcgen05 with fp32
accumulation in TMEM and does every epilogue (BN-affine, FiLM shift, LeakyReLU, residual, pooling,
after_conv, mask) in fp32.  This file restates the reference forward (``oracle/resunet_oracle.py``)
with exactly those rounding points inserted, so that

* the >= 40 dB SNR bar of BASELINE.json can be checked without a GPU (tests/test_bf16_model.py), and
* per-layer GPU outputs can be compared against a model that should agree to ~1 bf16 ulp.

Stored tensors per ConvBlockRes (see DESIGN.md "Data layout"):
  raw(x)  fp16   input of the shortcut / residual tap (saturating; 11-bit mantissa keeps the residual stream accurate)
  act(x)  bf16   lrelu(bn1(x) + beta1), computed from the fp32 value *before* rounding
"""
import math

import torch
import torch.nn.functional as F

from . import resunet_oracle as O


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def fp16(x: torch.Tensor) -> torch.Tensor:
    """Raw residual / skip stream: saturating fp16 (DESIGN.md "Numerics")."""
    return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


def fold_bn(sd, prefix):
    a = sd[prefix + ".weight"] / torch.sqrt(sd[prefix + ".running_var"] + O.BN_EPS)
    b = sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * a
    return a, b


def _act(sd, bn_prefix, film_name, cond, x32, lo=None, hi=None):
    """lrelu(a*x + b + beta) in fp32, rounded to bf16.  lo:hi selects a channel slice of the BN."""
    a, b = fold_bn(sd, bn_prefix)
    beta = F.linear(cond, sd["film." + film_name + ".weight"], sd["film." + film_name + ".bias"])
    if lo is not None:
        a, b, beta = a[lo:hi], b[lo:hi], beta[:, lo:hi]
    y = x32 * a[None, :, None, None] + (b[None, :] + beta)[:, :, None, None]
    return bf16(F.leaky_relu(y, O.LRELU_SLOPE))


def _block(sd, prefix, film_prefix, cond, raw_b, act_b):
    """raw_b/act_b: bf16-rounded raw and activated block input.  Returns fp32 block output."""
    h = F.conv2d(act_b, bf16(sd[prefix + ".conv1.weight"]), None, padding=1)
    a2 = _act(sd, prefix + ".bn2", film_prefix + "->beta2", cond, h)
    h2 = F.conv2d(a2, bf16(sd[prefix + ".conv2.weight"]), None, padding=1)
    if (prefix + ".shortcut.weight") in sd:
        res = F.conv2d(raw_b, fp16(sd[prefix + ".shortcut.weight"]), sd[prefix + ".shortcut.bias"])
    else:
        res = raw_b
    return res + h2


@torch.no_grad()
def forward(sd, mixture, condition, hop=160, taps=None):
    n_fft = O.infer_stft_params(sd)
    length = mixture.shape[2]
    mag, cos_in, sin_in = O.stft_mag_phase(sd, mixture[:, 0], n_fft, hop)
    x = O._bn(sd, "base.bn0", mag.transpose(1, 3)).transpose(1, 3)
    frames = x.shape[2]
    pad = int(math.ceil(frames / O.TIME_DOWNSAMPLE)) * O.TIME_DOWNSAMPLE - frames
    x = F.pad(x, (0, 0, 0, pad))[..., :-1]
    x32 = F.conv2d(x, sd["base.pre_conv.weight"], sd["base.pre_conv.bias"])

    skips = []
    for name, _ci, _co, pool in O.ENCODERS:
        p = "base.%s.conv_block1" % name
        fp = "%s->conv_block1" % name
        # encoder_block1's identity residual is regenerated from the fp32 magnitude in the conv epilogue (exact)
        raw_b = x32 if name == "encoder_block1" else fp16(x32)
        act_b = _act(sd, p + ".bn1", fp + "->beta1", condition, x32)
        full32 = _block(sd, p, fp, condition, raw_b, act_b)
        if taps is not None:
            taps[p + ":out"] = full32
        skips.append(full32)
        x32 = F.avg_pool2d(full32, kernel_size=pool)
    skips.pop()

    for name, _ci, cout, up in O.DECODERS:
        p = "base." + name
        a = _act(sd, p + ".bn1", name + "->beta1", condition, x32)
        u32 = F.conv_transpose2d(a, bf16(sd[p + ".conv1.weight"]), None, stride=up)
        s32 = skips.pop()
        cb = p + ".conv_block2"
        fp = name + "->conv_block2"
        raw_b = torch.cat((fp16(u32), fp16(s32)), dim=1)
        act_b = torch.cat((_act(sd, cb + ".bn1", fp + "->beta1", condition, u32, 0, cout),
                           _act(sd, cb + ".bn1", fp + "->beta1", condition, s32, cout, 2 * cout)), dim=1)
        x32 = _block(sd, cb, fp, condition, raw_b, act_b)
        if taps is not None:
            taps[cb + ":out"] = x32

    feat = F.conv2d(x32, sd["base.after_conv.weight"], sd["base.after_conv.bias"])
    feat = F.pad(feat, (0, 1))[:, :, :frames, :]
    if taps is not None:
        taps["feat"] = feat
    return O.mask_to_wave(sd, feat, mag, cos_in, sin_in, length, n_fft, hop)
