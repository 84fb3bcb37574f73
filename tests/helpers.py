"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

from oracle import factory

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def build_module(n_fft=1024, hop=160, seed=0, device="cpu"):
    """lass_b200 ResUNet30 with the key-seeded factory weights; returns (module, state_dict on CPU)."""
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    model = ResUNet30(1, 1, 512, window_size=n_fft, hop_size=hop).eval()
    sd = factory.fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return model.to(device), sd


def check_factory_checksums(sd):
    """The key-seeded weight factory must reproduce the weights the golden fixtures were generated with."""
    with open(os.path.join(GOLDEN, "factory_seed0_checksums.json")) as f:
        sums = json.load(f)
    assert sorted(sums) == sorted(sd)
    for k, v in sd.items():
        s, a = float(v.double().sum()), float(v.double().abs().sum())
        assert abs(s - sums[k][0]) <= 1e-6 * max(1.0, abs(sums[k][1])), k
        assert abs(a - sums[k][1]) <= 1e-6 * max(1.0, abs(sums[k][1])), k


def snr_ok(ref, est, min_db, zero_tol=1e-6):
    """Per-clip SNR >= min_db; an all-zero reference clip requires a (near) all-zero estimate."""
    snr = factory.snr_db(ref, est)
    for i in range(ref.shape[0]):
        if float(ref[i].abs().max()) == 0.0:
            assert float(est[i].abs().max()) <= zero_tol, "clip %d: reference is silent, estimate is not" % i
        else:
            assert float(snr[i]) >= min_db, "clip %d: SNR %.2f dB < %.1f dB" % (i, float(snr[i]), min_db)
    return snr
