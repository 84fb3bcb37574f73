"""Host-side logic of the drop-in module (no GPU): reference-compatible surface, loud failures, chunk stitching."""
import numpy as np
import pytest
import torch

from lass_b200.models import get_model_class
from lass_b200.models.resunet import FiLM, ResUNet30, get_film_meta

from helpers import build_module


def test_get_model_class_contract():
    assert get_model_class("ResUNet30") is ResUNet30
    with pytest.raises(NotImplementedError):
        get_model_class("HResUNet")


def test_film_meta_and_parameter_inventory():
    m = ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    meta = m.film_meta
    assert meta["encoder_block1"]["conv_block1"] == {"beta1": 32, "beta2": 32}
    assert meta["decoder_block6"]["beta1"] == 64 and meta["decoder_block6"]["conv_block2"] == {"beta1": 64, "beta2": 32}
    total = 0

    def walk(d):
        nonlocal total
        for v in d.values():
            if isinstance(v, dict):
                walk(v)
            else:
                total += v
    walk(meta)
    assert total == 9856
    assert sum(p.numel() for p in m.parameters()) == 29594693
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 26446917
    assert len(m.state_dict()) == 332
    assert get_film_meta(m.base) == meta
    # FiLM module returns the reference's nested (B, C, 1, 1) dict
    fd = m.film(conditions=torch.randn(2, 512))
    assert fd["encoder_block3"]["conv_block1"]["beta2"].shape == (2, 128, 1, 1)
    assert isinstance(m.film, FiLM)


def test_forward_fails_loudly_without_cuda():
    m, _ = build_module()
    x = {"mixture": torch.zeros(1, 1, 8000), "condition": torch.zeros(1, 512)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x)
    m.train()                       # the training-step engine has no CPU path either
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x)


def test_engine_film_rows_match_c_layout():
    from lass_b200 import _cabi, engine
    m, _ = build_module()
    sites = engine.film_sites(m.base)
    lib = _cabi.load()
    off = 0
    for i, (bn, name) in enumerate(sites):
        assert lib.lass_resunet30_film_offset(i) == off
        assert getattr(m.film, name).out_features == bn.num_features
        off += bn.num_features
    assert off == lib.lass_resunet30_film_rows() == 8256    # 9856 minus the six dead decoder beta2 (1600)
    assert lib.lass_resunet30_workspace_bytes(1, 160000, 1024, 160) > 400e6
    assert lib.lass_resunet30_workspace_bytes(1, 100, 1024, 160) == 0   # L <= n_fft/2: reflect padding impossible


class _IdentityEngine:
    """Stands in for the CUDA engine so the stitching logic can be checked on the CPU."""

    def forward(self, mixtures, conditions):
        return mixtures.clone()


@pytest.mark.parametrize("L", [9 * 800, 13 * 800 + 123, 5 * 800 + 1, 5 * 800, 4000])
def test_chunk_inference_stitching_equals_reference_loop(monkeypatch, L):
    """With a separator that returns its input, chunk_inference must reproduce the reference's stitching
    (models/resunet.py:671-714): serial loop restated here with rate 800 instead of 32000."""
    m, _ = build_module()
    monkeypatch.setattr(type(m.base), "_get_engine", lambda self, film: _IdentityEngine())
    rate = 800
    g = torch.Generator().manual_seed(L)
    mix = torch.randn(1, 1, L, generator=g)
    out = m.chunk_inference({"mixture": mix, "condition": torch.zeros(1, 512)}, rate=rate)

    NL, NC, NR = rate, 3 * rate, rate
    W = NL + NC + NR
    ref = np.zeros([1, L])
    cur = 0
    while cur + W < L:
        chunk = mix[0, :, cur:cur + W].numpy()
        if cur == 0:
            ref[:, cur:cur + W - NR] = chunk[:, :-NR]
        else:
            ref[:, cur + NL:cur + W - NR] = chunk[:, NL:-NR]
        cur += NC
        if cur < L:
            chunk = mix[0, :, cur:cur + W].numpy()
            ref[:, cur + NL:cur + chunk.shape[1]] = chunk[:, NL:]
    assert out.shape == (1, L) and out.dtype == np.float64
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-7)


def test_engine_does_not_keep_the_module_alive():
    """The per-module engine registry is weak: dropping the model drops its engine (packed weights, plans, workspace)."""
    import gc
    import weakref
    from lass_b200.models import resunet as R
    m = ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    eng = m.base._get_engine(m.film)
    assert m.base._get_engine(m.film) is eng and eng.base is m.base and eng.film is m.film
    n_before = len(R._ENGINES)
    ref_m, ref_e = weakref.ref(m.base), weakref.ref(eng)
    del m, eng
    gc.collect()
    assert ref_m() is None and ref_e() is None and len(R._ENGINES) == n_before - 1


def test_engine_parameter_key_tracks_updates_without_walking_the_module(monkeypatch):
    from lass_b200 import engine as E
    m = ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    eng = E.Engine(m.base, m.film)
    monkeypatch.setattr(E, "_dev_key", lambda device: ("cpu", 0))    # no CUDA here: the device part of the key is not under test
    k0 = eng._version_key("cpu")
    assert eng._version_key("cpu") == k0
    with torch.no_grad():
        m.base.after_conv.bias.add_(1.0)                # in-place update -> version counter
    k1 = eng._version_key("cpu")
    assert k1 != k0
    m.load_state_dict(m.state_dict())                   # load_state_dict -> epoch hook + versions
    k2 = eng._version_key("cpu")
    assert k2 != k1
    m.double()                                          # _apply -> epoch bump, new storage
    assert eng._version_key("cpu") != k2
