"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ``SegmentMixer`` (``data/waveform_mixers.py:9-62``) and its
helpers ``get_energy`` / ``get_energy_ratio`` / ``rescale_to_match_energy`` / ``dynamic_loudnorm`` (``:65-95``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this; the product (``lass_b200/``) never does.
Parity pin: ``tests/test_segment_mixer.py`` runs this restatement against the UNMODIFIED reference module (imported with the
``oracle/pyloudnorm`` stub) on the same seeded inputs with the same ``random`` seed, bit for bit, and against the committed
golden fixture ``tests/golden/segment_mixer_b6_l4000.npz`` that ``oracle/make_golden.py`` generated from the reference.

fp32 torch arithmetic in the reference's operation order, so the pin is exact; the CUDA kernel differs only in the summation
order of the two energies (tolerance stated in the GPU test).
"""
import random

import numpy as np
import torch


def get_energy_ratio(segment1: torch.Tensor, segment2: torch.Tensor) -> torch.Tensor:
    # data/waveform_mixers.py:72-82
    energy1 = torch.mean(segment1 ** 2)
    energy2 = max(torch.mean(segment2 ** 2), 1e-10)
    ratio = (energy1 / energy2) ** 0.5
    return torch.clamp(ratio, 0.02, 50)


def dynamic_loudnorm(audio, reference, lower_db=-10, higher_db=10):
    # data/waveform_mixers.py:85-93 (rescale_to_match_energy :65-69 inlined)
    rescaled = audio / get_energy_ratio(audio, reference)
    delta_loudness = random.randint(lower_db, higher_db)
    gain = np.power(10.0, delta_loudness / 20.0)
    return gain * rescaled


def segment_mix(waveforms: torch.Tensor, max_mix_num: int, lower_db: int, higher_db: int):
    """waveforms (B, ..., L) fp32 on the CPU -> (mixture, segment), consuming Python's ``random`` like the reference
    (``data/waveform_mixers.py:19-62``)."""
    batch_size = waveforms.shape[0]
    segments, mixtures = [], []
    for n in range(batch_size):
        segment = waveforms[n].clone()
        noise = torch.zeros_like(segment)
        mix_num = random.randint(2, max_mix_num)
        for i in range(1, mix_num):
            noise += dynamic_loudnorm(waveforms[(n + i) % batch_size], segment, lower_db, higher_db)
        noise = dynamic_loudnorm(noise, segment, lower_db, higher_db)
        mixture = segment + noise
        max_value = torch.max(torch.abs(mixture))
        if max_value > 1:
            segment *= 0.9 / max_value
            mixture *= 0.9 / max_value
        segments.append(segment)
        mixtures.append(mixture)
    return torch.stack(mixtures, dim=0), torch.stack(segments, dim=0)


def make_waveforms(B: int, L: int, seed: int = 0, channel_dim: bool = True) -> torch.Tensor:
    """Seeded training-batch stand-in: clips of very different loudness (0.003 .. 1.2 rms-ish), clip 1 silent (exercises the
    1e-10 energy clamp and the 0.02 / 50 ratio clamps), clip 2 near full scale (exercises the de-clipping branch)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(B, L, generator=g) * (0.003 * 3.0 ** torch.arange(B, dtype=torch.float32) % 1.3)[:, None]
    t = torch.arange(L, dtype=torch.float32)
    w += 0.05 * torch.sin(2 * np.pi * 440.0 * t / 16000.0)[None, :] * torch.rand(B, 1, generator=g)
    if B > 1:
        w[1] = 0.0
    if B > 2:
        w[2] = (0.9 * torch.sin(2 * np.pi * 1000.0 * t / 16000.0) + 0.05 * torch.randn(L, generator=g)).clamp(-1, 1)
    return w[:, None, :].contiguous() if channel_dim else w.contiguous()
