"""Time SegmentMixer (lass_segment_mix: two launches) at BASELINE config 4's per-rank batch against the UNMODIFIED reference
module running on the same GPU (oracle/_ref, PyTorch eager) and on the host cores.  Output: one JSON line."""
import json
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200.data.waveform_mixers import SegmentMixer  # noqa: E402
from oracle import reference_loader  # noqa: E402
from oracle.segment_mixer_oracle import make_waveforms  # noqa: E402


def timed(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters, (time.perf_counter() - t0) * 1e6 / iters


def main():
    out = {}
    for B, L, mm in ((16, 80000, 2), (16, 80000, 5), (64, 160000, 2)):
        wave = make_waveforms(B, L, seed=1)
        dev = wave.cuda()
        mixer = SegmentMixer(mm, -10, 10)
        random.seed(0)
        for _ in range(5):
            mixer(dev)
        dev_us, wall_us = timed(lambda: mixer(dev), 200)
        row = {"lass_b200_device_us": dev_us, "lass_b200_wall_us": wall_us,
               "algorithmic_bytes": 3 * B * L * 4, "gbs_at_device_time": 3 * B * L * 4 / dev_us / 1e3}
        if reference_loader.mixers_available():
            ref = reference_loader.import_reference_mixers().SegmentMixer(mm, -10, 10)
            for _ in range(2):
                ref(dev)
            row["reference_on_this_gpu_us"] = timed(lambda: ref(dev), 10)[1]
            t0 = time.perf_counter()
            for _ in range(3):
                ref(wave)
            row["reference_cpu_us"] = (time.perf_counter() - t0) * 1e6 / 3
        out["%dx%d_mix%d" % (B, L, mm)] = row
    print(json.dumps(out))


if __name__ == "__main__":
    main()
