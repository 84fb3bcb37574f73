"""SegmentMixer (reference data/waveform_mixers.py:9-62), the step in front of the training path (models/audiosep.py:76-78).

not gpu: the restatement ``oracle/segment_mixer_oracle.py`` is pinned bit for bit to the golden fixture generated from the
UNMODIFIED reference and (in the build container) to the reference module itself; the host side of the product (the random plan)
consumes Python's ``random`` stream exactly like the reference and encodes the draws the kernel needs.
gpu: ``lass_b200.data.waveform_mixers.SegmentMixer`` (C-ABI ``lass_segment_mix``) against the oracle / golden on the same seeded
inputs; floating point -- the kernel sums the two energies in another order than ``torch.mean``, so the bar is
``max|d| <= 1e-5 max|ref|`` per clip (observed ~2e-7), stated below as TOL.
"""
import random

import numpy as np
import pytest
import torch

from helpers import golden
from oracle import reference_loader
from oracle.segment_mixer_oracle import make_waveforms, segment_mix

TOL = 1e-5


def _golden_case():
    g = golden("segment_mixer_b6_l4000.npz")
    B, L, max_mix_num, lower_db, higher_db, seed, wave_seed = [int(v) for v in g["meta"]]
    return g, make_waveforms(B, L, seed=wave_seed), (max_mix_num, lower_db, higher_db), seed


def test_oracle_matches_golden_of_the_reference():
    g, wave, cfg, seed = _golden_case()
    random.seed(seed)
    mixture, segment = segment_mix(wave, *cfg)
    assert mixture.shape == wave.shape and segment.shape == wave.shape
    assert np.array_equal(mixture.numpy(), g["mixture"]) and np.array_equal(segment.numpy(), g["segment"])
    # the fixture exercises what it should: a silent clip, the de-clipping branch (max exactly 0.9) and the plain branch
    peak = np.abs(g["mixture"]).reshape(wave.shape[0], -1).max(axis=1)
    assert float(np.abs(g["segment"][1]).max()) == 0.0
    assert (np.abs(peak - 0.9) < 1e-6).sum() >= 2 and (peak < 0.89).sum() >= 2


@pytest.mark.skipif(not reference_loader.mixers_available(), reason="reference tree not present")
@pytest.mark.parametrize("B,L,max_mix_num,db,channel_dim", [(6, 4000, 4, 10, True), (1, 1000, 2, 10, True), (2, 333, 2, 3, True),
                                                            (5, 2048, 7, 20, False), (16, 8000, 2, 10, True)])
def test_oracle_matches_unmodified_reference(B, L, max_mix_num, db, channel_dim):
    ref = reference_loader.import_reference_mixers()
    wave = make_waveforms(B, L, seed=B + L, channel_dim=channel_dim)
    random.seed(B)
    want = ref.SegmentMixer(max_mix_num=max_mix_num, lower_db=-db, higher_db=db)(wave.clone())
    after_ref = random.random()
    random.seed(B)
    got = segment_mix(wave, max_mix_num, -db, db)
    assert random.random() == after_ref                      # same number of draws
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def _apply_plan(wave, plan):
    """What the plan table means (include/lass_b200.h, lass_segment_mix), in torch fp32 on the CPU, reference operation order."""
    B = wave.shape[0]
    flat = wave.reshape(B, -1)
    mixtures, segments = [], []

    def ratio(a, b):
        return torch.clamp((torch.mean(a ** 2) / max(torch.mean(b ** 2), 1e-10)) ** 0.5, 0.02, 50)
    for n in range(B):
        seg = flat[n].clone()
        noise = torch.zeros_like(seg)
        for i in range(1, int(plan[n, 0])):
            nxt = flat[(n + i) % B]
            noise += float(plan[n, i]) * (nxt / ratio(nxt, seg))
        noise = float(plan[n, -1]) * (noise / ratio(noise, seg))
        mix = seg + noise
        mx = mix.abs().max()
        if mx > 1:
            seg *= 0.9 / mx
            mix *= 0.9 / mx
        mixtures.append(mix)
        segments.append(seg)
    return torch.stack(mixtures).reshape(wave.shape), torch.stack(segments).reshape(wave.shape)


def test_plan_consumes_random_like_the_reference_and_encodes_the_draws():
    from lass_b200.data.waveform_mixers import draw_plan
    g, wave, cfg, seed = _golden_case()
    random.seed(seed)
    segment_mix(wave, *cfg)
    after_oracle = random.random()
    random.seed(seed)
    plan = draw_plan(wave.shape[0], *cfg)
    assert random.random() == after_oracle
    assert plan.shape == (wave.shape[0], cfg[0] + 1) and plan.dtype == np.float32
    assert ((plan[:, 0] >= 2) & (plan[:, 0] <= cfg[0])).all()
    for n in range(plan.shape[0]):                           # unused gain slots stay zero
        assert (plan[n, int(plan[n, 0]):cfg[0]] == 0).all() and plan[n, -1] > 0
    mixture, segment = _apply_plan(wave, plan)
    assert np.array_equal(mixture.numpy(), g["mixture"]) and np.array_equal(segment.numpy(), g["segment"])
    with pytest.raises(ValueError):
        draw_plan(4, 1, -10, 10)


def test_segment_mixer_refuses_cpu_tensors():
    from lass_b200.data.waveform_mixers import SegmentMixer
    m = SegmentMixer(max_mix_num=2, lower_db=-10, higher_db=10)
    assert m.max_mix_num == 2 and m.loudness_param == {"lower_db": -10, "higher_db": 10}      # reference attributes (:13-17)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 100))


# --------------------------------------------------------------------------------------------------------------- GPU
def _close(got, want):
    got, want = got.detach().cpu(), want.detach().cpu()
    B = want.shape[0]
    err = (got - want).reshape(B, -1).abs().max(dim=1).values
    ref = want.reshape(B, -1).abs().max(dim=1).values
    assert bool((err <= TOL * ref + 1e-12).all()), (err, ref)
    return float((err / ref.clamp_min(1e-30)).max())


@pytest.mark.gpu
def test_cuda_mixer_matches_golden_and_oracle():
    from lass_b200.data.waveform_mixers import SegmentMixer
    g, wave, cfg, seed = _golden_case()
    mixer = SegmentMixer(*cfg)
    random.seed(seed)
    mixture, segment = mixer(wave.cuda())
    after = random.random()
    assert mixture.shape == wave.shape and mixture.is_cuda and mixture.dtype == torch.float32
    _close(mixture, torch.from_numpy(g["mixture"]))
    _close(segment, torch.from_numpy(g["segment"]))
    assert float(segment[1].abs().max()) == 0.0                          # silent segment stays silent
    random.seed(seed)
    segment_mix(wave, *cfg)
    assert random.random() == after                                      # same draws consumed as the reference
    # deterministic: the clip-wide reductions run in a fixed order
    random.seed(seed)
    again = mixer(wave.cuda())
    assert torch.equal(again[0], mixture) and torch.equal(again[1], segment)


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,max_mix_num,db,channel_dim", [
    (16, 80000, 2, 10, True),       # BASELINE config 4 per rank, the reference's max_mix_num (config/audiosep_base.yaml)
    (16, 80000, 5, 10, True),
    (1, 1000, 2, 10, True),         # a clip mixed with itself ((n + i) % 1)
    (3, 4001, 6, 20, False),        # odd length (rows not 16-byte aligned), mix_num > B wraps around, (B, L) input
    (2, 7, 2, 3, True),             # fewer samples than CTAs in the cluster
    (2, 420000, 3, 10, True),       # longer than the shared-memory noise cache: the recomputing variant
    (64, 160000, 3, 10, True),      # 64 x 10 s
])
def test_cuda_mixer_matches_oracle(B, L, max_mix_num, db, channel_dim):
    from lass_b200.data.waveform_mixers import SegmentMixer
    wave = make_waveforms(B, L, seed=B + L, channel_dim=channel_dim)
    random.seed(B + 1)
    want_m, want_s = segment_mix(wave, max_mix_num, -db, db)
    random.seed(B + 1)
    got_m, got_s = SegmentMixer(max_mix_num, -db, db)(wave.cuda())
    _close(got_m, want_m)
    _close(got_s, want_s)
    # properties that hold at any size: never above full scale after de-clipping; segment is the input up to one scalar per clip
    assert float(got_m.abs().max()) <= 1.0 + 1e-6
    scale = (got_s.reshape(B, -1) * wave.cuda().reshape(B, -1)).sum(1) / (wave.cuda().reshape(B, -1) ** 2).sum(1).clamp_min(1e-30)
    assert bool(((scale > 0.0) & (scale <= 1.0 + 1e-6) | (wave.reshape(B, -1).abs().sum(1).cuda() == 0)).all())


@pytest.mark.gpu
def test_cuda_mixer_feeds_the_training_shell():
    """AudioSep.training_step's first two statements (models/audiosep.py:69-78) with the GPU mixer: seeded by batch_idx."""
    from lass_b200.data.waveform_mixers import SegmentMixer
    wave = make_waveforms(4, 16000, seed=11).cuda()
    mixer = SegmentMixer(max_mix_num=2, lower_db=-10, higher_db=10)
    random.seed(5)
    a = mixer(waveforms=wave)
    random.seed(5)
    b = mixer(waveforms=wave)
    random.seed(6)
    c = mixer(waveforms=wave)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and not torch.equal(a[0], c[0])


@pytest.mark.gpu
def test_cabi_argument_errors():
    from lass_b200 import _cabi
    lib = _cabi.load()
    x = torch.zeros(2, 100, device="cuda")
    plan = torch.zeros(2, 3, device="cuda")
    out = torch.zeros(2, 2, 100, device="cuda")
    scratch = torch.zeros(64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    assert lib.lass_segment_mix(x.data_ptr(), 2, 100, 1, plan.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), scratch.data_ptr(), 256, s) != 0
    assert lib.lass_segment_mix(x.data_ptr(), 2, 100, 2, plan.data_ptr(), out[0].data_ptr(), out[0].data_ptr(), scratch.data_ptr(), 256, s) != 0
    assert lib.lass_segment_mix(x.data_ptr(), 2, 100, 2, plan.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), scratch.data_ptr(), 8, s) != 0
    assert b"lass_segment_mix" in lib.lass_last_error()
