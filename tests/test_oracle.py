"""Pin the travelling oracle (oracle/resunet_oracle.py + oracle/torchlibrosa) against

* the golden fixtures generated from the unmodified reference (tests/golden, oracle/make_golden.py),
* the unmodified reference itself when /root/reference is present (build container only),
* fp64 torch.stft / torch.istft as an independent spectral oracle.
"""
import math

import numpy as np
import pytest
import torch

from oracle import bf16_model, factory, resunet_oracle as O
from oracle.reference_loader import import_reference_resunet, reference_available
from oracle.torchlibrosa.stft import ISTFT, STFT

from helpers import build_module, check_factory_checksums, golden, snr_ok


@pytest.fixture(scope="module")
def sd():
    _, sd = build_module()
    return sd


def test_factory_weights_match_golden_checksums(sd):
    check_factory_checksums(sd)


def test_oracle_forward_matches_reference_golden(sd):
    g = golden("resunet30_fwd_b3_l24000.npz")
    B, L, n_fft, hop, seed, _ = [int(v) for v in g["meta"]]
    mix, cond = factory.make_inputs(B, L, seed=seed)
    taps = {}
    wav = O.resunet30_forward(sd, mix, cond, hop=hop, taps=taps)
    ref = torch.from_numpy(g["waveform"])
    # same aten ops in the same order as the reference: agreement to fp32 round-off
    assert factory.max_rel_err(ref, wav) <= 1e-5
    assert factory.max_rel_err(torch.from_numpy(g["mag"]), taps["mag"]) <= 1e-6
    assert float((torch.from_numpy(g["cos_clip0"]) - taps["cos"][0]).abs().max()) <= 1e-5
    assert float(wav[1].abs().max()) == 0.0          # silent clip stays silent (both clamps exercised)


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_forward_matches_unmodified_reference(sd):
    ref_mod = import_reference_resunet()
    net = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512).eval()
    net.load_state_dict(sd)
    mix, cond = factory.make_inputs(2, 16000, seed=3, edge_clips=False)
    with torch.no_grad():
        ref = net({"mixture": mix, "condition": cond})["waveform"]
    assert factory.max_rel_err(ref, O.resunet30_forward(sd, mix, cond)) <= 1e-6


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_module_state_dict_is_interchangeable_with_reference():
    ref_mod = import_reference_resunet()
    torch.manual_seed(0)
    ref = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    mine = ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b) and len(a) == 332
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        # same constructor order => same RNG draws; DFT matrices agree to fp32 round-off
        assert float((a[k].double() - b[k].double()).abs().max()) <= 1e-6, k
    mine.load_state_dict(a)
    ref.load_state_dict(b)
    assert mine.film_meta == ref.film_meta


@pytest.mark.parametrize("n_fft,hop", [(1024, 160), (2048, 320), (512, 160), (256, 160)])
def test_restated_stft_matches_golden_and_fp64(n_fft, hop):
    g = golden("stft_%d_%d_l8000.npz" % (n_fft, hop))
    wave, _ = factory.make_inputs(2, 8000, seed=5, edge_clips=False)
    stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
    istft = ISTFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
    with torch.no_grad():
        re, im = stft(wave[:, 0])
        back = istft(re, im, 8000)
    assert factory.max_rel_err(torch.from_numpy(g["real"]), re) <= 1e-6
    assert factory.max_rel_err(torch.from_numpy(g["imag"]), im) <= 1e-6
    assert factory.max_rel_err(torch.from_numpy(g["roundtrip"]), back) <= 1e-6
    spec = torch.stft(wave[:, 0].double(), n_fft, hop, n_fft, torch.hann_window(n_fft, periodic=True, dtype=torch.float64),
                      center=True, pad_mode="reflect", return_complex=True).transpose(1, 2)[:, None]
    assert factory.max_rel_err(spec.real, re) <= 1e-5
    assert factory.max_rel_err(spec.imag, im) <= 1e-5
    assert float((back - wave[:, 0]).abs().max()) <= 1e-5        # perfect reconstruction


def test_oracle_chunk_inference_matches_reference_golden(sd):
    g = golden("chunk_inference_l230000.npz")
    _, L, seed = [int(v) for v in g["meta"]]
    mix, cond = factory.make_inputs(1, L, seed=seed, edge_clips=False)
    out = O.chunk_inference(sd, mix, cond)
    ref = g["waveform"]
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_bf16_rounding_model_meets_the_40db_bar(sd):
    """BASELINE.json: separated-waveform SNR vs the fp32 reference >= 40 dB for the bf16 path.  The CPU model of
    the B200 path's rounding points (raw residual stream fp16, activations / weights bf16, fp32 accumulate)."""
    mix, cond = factory.make_inputs(3, 24000)
    ref = O.resunet30_forward(sd, mix, cond)
    snr_ok(ref, bf16_model.forward(sd, mix, cond), 40.0)
