"""Multi-resolution STFT front end (BASELINE config 5's runnable part).

Mirrors reference ``scripts/precompute_stfts.py:19-58`` (``calculate_stft_components``): for each window length
``n_fft = win_length in {256, 512, 2048}`` with hop 160 it returns magnitude, cos and sin of the torchlibrosa STFT with
``magphase`` semantics, each ``(B, 1, T, n_fft//2 + 1)``.  Every resolution is one launch of kernel K1 (``lass_stft_fwd``,
tcgen05 DFT-GEMM) with ``magphase_mode = 1``; there is no PyTorch fallback.  The reference's multi-resolution trunk
(``models/resunet_with_multistft.py``) is non-functional as shipped, so only the front end is offered.
"""
from typing import Dict, Sequence, Tuple

import torch

from . import ops, packing
from .models.spectral import STFT

_BASIS_CACHE: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def _basis(n_fft: int, hop: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    key = (n_fft, hop, str(device))
    if key not in _BASIS_CACHE:
        stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
        _BASIS_CACHE[key] = packing.pack_stft_basis(stft.conv_real.weight.data.to(device),
                                                     stft.conv_imag.weight.data.to(device))
    return _BASIS_CACHE[key]


def calculate_stft_components(waveform: torch.Tensor, n_fft: int, hop_length: int, win_length: int = None,
                              window: str = "hann", center: bool = True, pad_mode: str = "reflect"):
    """Same contract as reference ``scripts/precompute_stfts.py:19-58``: waveform (B, 1, L) or (B, L) on a CUDA
    device -> (magnitude, cos_phase, sin_phase), each (B, 1, T, n_fft//2 + 1) fp32 contiguous."""
    if win_length is not None and win_length != n_fft:
        raise NotImplementedError("the reference only uses n_fft == win_length (scripts/precompute_stfts.py:573-582)")
    if window != "hann" or not center or pad_mode != "reflect":
        raise NotImplementedError("only the reference's STFT configuration (hann, center, reflect)")
    if waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    waveform = waveform.float().contiguous()
    with torch.cuda.device(waveform.device):
        hi, lo = _basis(n_fft, hop_length, waveform.device)
        return ops.stft_fwd(waveform, hi, lo, n_fft, hop_length, precision_mode=0, magphase_mode=1)


def multires_stft(waveform: torch.Tensor, win_lengths: Sequence[int] = (256, 512, 2048), hop_length: int = 160):
    """{win_length: (mag, cos, sin)} for the reference's three resolutions (``config/audiosep_base.yaml:17-21``) from one
    kernel launch (up to three resolutions per launch; longer lists are processed three at a time)."""
    if waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    waveform = waveform.float().contiguous()
    wins = [int(w) for w in win_lengths]
    out = {}
    with torch.cuda.device(waveform.device):
        for i in range(0, len(wins), 3):
            group = wins[i:i + 3]
            bases = [_basis(w, hop_length, waveform.device) for w in group]
            res = ops.stft_multi_fwd(waveform, bases, group, hop_length, precision_mode=0, magphase_mode=1)
            out.update(dict(zip(group, res)))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Pre-computed STFT shards: the data format on the far side of this front end (reference scripts/precompute_stfts.py:60-123
# `save_batch_precomputed_data`, item layout :596-622, common parameters :700-705; consumed by
# models/audiosep_with_neg_query.py:45-90).  One `.pt` file per batch = a LIST with one dict per item:
#   {'stfts': {'mixture': {win_len: (mag, cos, sin)}, 'segment': {win_len: (mag, cos, sin)}},     each (1, 1, T, win_len//2 + 1) CPU
#    'target_waveform': (1, L) CPU, 'text': str, 'mixture_component_texts': [str], 'stft_common_params': {...},
#    'stft_win_lengths': [int]}
# ---------------------------------------------------------------------------------------------------------------------
def shard_common_params(hop_length: int = 160) -> dict:
    """The 'stft_common_params' entry (reference scripts/precompute_stfts.py:700-705)."""
    return {"hop_length": hop_length, "window": "hann", "center": True, "pad_mode": "reflect"}


def build_shard_items(mixtures: torch.Tensor, segments: torch.Tensor, texts, mixture_component_texts,
                      win_lengths: Sequence[int] = (256, 512, 2048), hop_length: int = 160) -> list:
    """Per-item dicts of one batch, as the reference assembles them (scripts/precompute_stfts.py:570-622): mixtures / segments
    (B, 1, L) on a CUDA device; both go through ONE multi-resolution K1 launch each; item k holds the [k:k+1] slices."""
    assert mixtures.shape == segments.shape and mixtures.dim() == 3 and mixtures.shape[1] == 1
    B = mixtures.shape[0]
    assert len(texts) == B and len(mixture_component_texts) == B
    wins = [int(w) for w in win_lengths]
    mix = multires_stft(mixtures, wins, hop_length)
    seg = multires_stft(segments, wins, hop_length)
    common = shard_common_params(hop_length)
    items = []
    for k in range(B):
        items.append({
            "stfts": {"mixture": {w: tuple(t[k:k + 1] for t in mix[w]) for w in wins},
                      "segment": {w: tuple(t[k:k + 1] for t in seg[w]) for w in wins}},
            "target_waveform": segments[k],
            "text": texts[k],
            "mixture_component_texts": list(mixture_component_texts[k]),
            "stft_common_params": common,
            "stft_win_lengths": wins,
        })
    return items


def _to_cpu(value):
    if isinstance(value, torch.Tensor):
        return value.detach().cpu()
    if isinstance(value, tuple):
        return tuple(_to_cpu(v) for v in value)
    if isinstance(value, dict):
        return {k: _to_cpu(v) for k, v in value.items()}
    return value


def save_batch_precomputed_data(output_dir, batch_index: int, batch_data_list: list) -> int:
    """Same contract as reference scripts/precompute_stfts.py:60-123: writes ``batch_{index:06d}.pt`` (a list of item dicts with
    every tensor detached on the CPU) and returns the number of items saved; an empty list writes nothing and returns 0."""
    import os
    if not batch_data_list:
        return 0
    os.makedirs(str(output_dir), exist_ok=True)
    filename = os.path.join(str(output_dir), "batch_%06d.pt" % batch_index)
    torch.save([_to_cpu(d) for d in batch_data_list], filename)
    return len(batch_data_list)


def load_batch_precomputed_data(filename) -> list:
    """Reads a shard written by this module or by the reference script."""
    return torch.load(str(filename), map_location="cpu", weights_only=False)
