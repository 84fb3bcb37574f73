"""TEST INFRASTRUCTURE ONLY — seeded weights and inputs shared by every parity test.

Follows SURVEY.md §8(d): BatchNorm running statistics and affine parameters are randomised
(with ``init_bn`` defaults — reference ``models/base.py:18-21`` — every BN-folding bug would be
invisible), conv / linear weights are Xavier-uniform like ``init_layer`` (reference
``models/base.py:9-15``) but biases are made non-zero so that bias handling is exercised.

Values depend only on (key name, shape, seed) so the same weights can be produced in the build
container (to be loaded into the unmodified reference) and on the GPU box (to be loaded into the
B200 module and the travelling oracle) without shipping a 118 MB state dict.
"""
import math
import zlib

import torch

_FROZEN_KEYS = ("stft.conv_real.weight", "stft.conv_imag.weight",
                "istft.conv_real.weight", "istft.conv_imag.weight", "istft.ola_window")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode("utf-8")) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def is_frozen_dft_key(key: str) -> bool:
    return key.endswith(_FROZEN_KEYS)


def fill_state_dict(template: dict, seed: int = 0) -> dict:
    """Return a new state dict with the template's keys/shapes and seeded values.

    The deterministic DFT / IDFT / ola_window entries are copied from the template unchanged.
    """
    out = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if is_frozen_dft_key(key):
            out[key] = ref.detach().clone()
        elif leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif ".bn" in key and leaf == "weight":
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif ".bn" in key and leaf == "bias":
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight":
            # Xavier-uniform; conv: (out, in, kh, kw); convT: (in, out, kh, kw); linear: (out, in)
            recept = 1
            for s in shape[2:]:
                recept *= s
            fan_a, fan_b = shape[0] * recept, shape[1] * recept
            bound = math.sqrt(6.0 / (fan_a + fan_b))
            out[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        elif leaf == "bias":
            out[key] = 0.05 * torch.randn(shape, generator=g)
        else:
            raise KeyError("factory does not know how to fill %s %s" % (key, shape))
        out[key] = out[key].to(ref.dtype)
    return out


def make_inputs(batch: int, length: int, seed: int = 1234, edge_clips: bool = True,
                condition_size: int = 512, channels: int = 1):
    """mixture (B, C, L) = 0.1*N(0,1) clipped to [-1,1]; condition (B, 512) unit-norm.

    With ``edge_clips`` the last two clips (when B >= 3) are replaced by all-zeros (exercises both
    magnitude clamps) and a full-scale 1 kHz sine at 16 kHz (peaky spectrum) — SURVEY.md §8(d).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    mixture = (0.1 * torch.randn(batch, channels, length, generator=g)).clamp_(-1.0, 1.0)
    condition = torch.nn.functional.normalize(torch.randn(batch, condition_size, generator=g), dim=-1)
    if edge_clips and batch >= 3:
        mixture[-2].zero_()
        t = torch.arange(length, dtype=torch.float64) / 16000.0
        mixture[-1] = torch.sin(2.0 * math.pi * 1000.0 * t).to(torch.float32)[None, :]
    return mixture, condition


def snr_db(ref: torch.Tensor, est: torch.Tensor) -> torch.Tensor:
    """Per-clip SNR = 10 log10(sum ref^2 / sum (ref-est)^2) (formula of reference ``utils.py:148-169``)."""
    ref = ref.double().reshape(ref.shape[0], -1)
    est = est.double().reshape(est.shape[0], -1)
    num = (ref ** 2).sum(-1).clamp_min(1e-30)
    den = ((ref - est) ** 2).sum(-1).clamp_min(1e-30)
    return 10.0 * torch.log10(num / den)


def max_rel_err(ref: torch.Tensor, est: torch.Tensor) -> float:
    """max|ref-est| / max|ref| (the spectral parity metric of SURVEY.md §8(d))."""
    denom = float(ref.double().abs().max())
    if denom == 0.0:
        return float((ref.double() - est.double()).abs().max())
    return float((ref.double() - est.double()).abs().max()) / denom
