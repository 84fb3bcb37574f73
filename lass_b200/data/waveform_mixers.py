"""``SegmentMixer`` — drop-in for the reference's ``data/waveform_mixers.py:9-62`` on the GPU.

The reference builds every training mixture in a Python loop over the batch (per clip: draw ``mix_num``, rescale the next
``mix_num - 1`` clips of the batch to the segment's energy with a random dB offset, sum them, rescale the sum once more, add,
de-clip).  Same constructor, same call, same results here — but the arithmetic is ONE C-ABI call (``lass_segment_mix``: two
kernel launches for the whole batch, ``lass_b200/csrc/mixer.cu``).  The random draws stay on the host and consume Python's
``random`` stream in exactly the reference's order (``mix_num`` at ``:35``, one offset per mixed-in clip at ``:89`` via ``:40``,
one for the summed noise via ``:44``), so ``random.seed(batch_idx)`` in ``AudioSep.training_step`` (``models/audiosep.py:69``)
pins the mixtures of a step on every rank exactly as it does in the reference.

There is no CPU path: a non-CUDA tensor raises.
"""
import random

import numpy as np
import torch

from .. import _cabi


def draw_plan(batch_size: int, max_mix_num: int, lower_db: int, higher_db: int) -> np.ndarray:
    """The reference's random draws for one batch as the ``(B, max_mix_num + 1)`` fp32 table ``lass_segment_mix`` takes:
    column 0 = ``mix_num``, columns ``1..mix_num-1`` = gains of the mixed-in clips, last column = gain of the summed noise.
    Gains are ``np.power(10.0, dB / 20.0)`` (``data/waveform_mixers.py:91``) rounded to fp32, which is what multiplying a
    float32 tensor by that scalar does in the reference."""
    if max_mix_num < 2:
        raise ValueError("max_mix_num must be >= 2 (the reference asserts mix_num >= 2, data/waveform_mixers.py:36)")
    plan = np.zeros((batch_size, max_mix_num + 1), dtype=np.float32)
    gains = {db: np.float32(np.power(10.0, db / 20.0)) for db in range(lower_db, higher_db + 1)}
    randint = random.randint
    for n in range(batch_size):
        mix_num = randint(2, max_mix_num)
        row = plan[n]
        row[0] = mix_num
        for i in range(1, mix_num):
            row[i] = gains[randint(lower_db, higher_db)]
        row[max_mix_num] = gains[randint(lower_db, higher_db)]
    return plan


class SegmentMixer(torch.nn.Module):
    def __init__(self, max_mix_num, lower_db, higher_db):
        super().__init__()
        self.max_mix_num = max_mix_num
        self.loudness_param = {"lower_db": lower_db, "higher_db": higher_db}
        self._scratch = {}          # device index -> fp32 scratch of the energy partials
        self._ring = {}             # (device index, B) -> ring of pinned plan buffers + the events of their last copies

    _RING = 8

    def _upload_plan(self, plan_np, device):
        """Asynchronous host -> device copy of the draws through a small ring of pinned buffers (a slot is rewritten only after
        the copy that last read it has completed), so the call never blocks on the GPU."""
        key = (device.index, plan_np.shape)
        ring = self._ring.get(key)
        if ring is None:
            ring = self._ring[key] = {"next": 0, "slots": [(torch.empty(plan_np.shape, dtype=torch.float32).pin_memory(),
                                                            torch.cuda.Event()) for _ in range(self._RING)]}
        host, event = ring["slots"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % self._RING
        event.synchronize()                         # no-op unless the copy of _RING calls ago is still in flight
        host.numpy()[...] = plan_np
        dev = host.to(device, non_blocking=True)
        event.record()
        return dev

    def __call__(self, waveforms):
        """waveforms ``(B, 1, L)`` (the training batch, ``models/audiosep.py:56-58``) or ``(B, L)`` float32 on a CUDA device
        -> ``(mixture, segment)`` of the same shape."""
        if not isinstance(waveforms, torch.Tensor) or not waveforms.is_cuda:
            raise RuntimeError("lass_b200 SegmentMixer needs a CUDA tensor (no CPU fallback)")
        if waveforms.dtype != torch.float32:
            raise TypeError("SegmentMixer: float32 waveforms expected, got %s" % waveforms.dtype)
        if waveforms.dim() == 3 and waveforms.shape[1] != 1 or waveforms.dim() not in (2, 3):
            raise ValueError("SegmentMixer: waveforms (B, 1, L) or (B, L), got %s" % (tuple(waveforms.shape),))
        B, L = waveforms.shape[0], waveforms.shape[-1]
        wave = waveforms.contiguous()
        with torch.cuda.device(wave.device):
            plan = self._upload_plan(draw_plan(B, self.max_mix_num, **self.loudness_param), wave.device)
            lib = _cabi.load()
            need = lib.lass_segment_mix_scratch_bytes(B)
            scratch = self._scratch.get(wave.device.index)
            if scratch is None or scratch.numel() * 4 < need:
                scratch = torch.empty((need + 3) // 4, dtype=torch.float32, device=wave.device)
                self._scratch[wave.device.index] = scratch
            mixture, segment = torch.empty_like(wave), torch.empty_like(wave)
            _cabi.check(lib.lass_segment_mix(wave.data_ptr(), B, L, self.max_mix_num, plan.data_ptr(), mixture.data_ptr(),
                                             segment.data_ptr(), scratch.data_ptr(), scratch.numel() * 4,
                                             torch.cuda.current_stream().cuda_stream))
        return mixture, segment
