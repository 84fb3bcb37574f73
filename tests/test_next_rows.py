"""SURVEY.md §8(f) rows that are built this round: multi-resolution STFT front end (GPU), CLAP stand-in + AudioSep shell."""
import pytest
import torch

from oracle import factory
from oracle.torchlibrosa.stft import STFT, magphase


def test_clap_standin_contract_and_cache():
    from lass_b200.models.clap_standin import RandomInitCLAPTextEncoder, synthetic_token_ids
    ids = synthetic_token_ids("a dog barking loudly")
    assert ids.shape == (512,) and int(ids[0]) == 0 and int(ids[5]) == 2 and int(ids[6]) == 1
    enc = RandomInitCLAPTextEncoder(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128)
    emb = enc.get_query_embed(modality="text", text=["a dog barking", "rain", "a dog barking"])
    assert emb.shape == (3, 512) and emb.dtype == torch.float32
    assert torch.allclose(emb.norm(dim=-1), torch.ones(3), atol=1e-5)          # L2-normalised like CLAP
    assert torch.equal(emb[0], emb[2]) and not torch.equal(emb[0], emb[1])
    assert len(enc._cache) == 2                                                # one RoBERTa pass per distinct caption
    with pytest.raises(NotImplementedError):
        enc.get_query_embed(modality="audio", audio=torch.zeros(1, 100))


def test_audiosep_shell_keeps_reference_surface():
    from lass_b200.models.audiosep import AudioSep, get_model_class
    from lass_b200.models.resunet import ResUNet30
    assert get_model_class("ResUNet30") is ResUNet30
    m = AudioSep(ss_model=torch.nn.Identity(), query_encoder=None)
    assert m.forward(torch.zeros(1)) is None                                   # reference forward is `pass`
    assert hasattr(m, "ss_model") and hasattr(m, "query_encoder") and hasattr(m, "use_text_ratio")


@pytest.mark.gpu
@pytest.mark.parametrize("win", [256, 512, 2048])
def test_multires_front_end_matches_reference_semantics(win):
    """calculate_stft_components (scripts/precompute_stfts.py:19-58): STFT + magphase at hop 160."""
    from lass_b200 import multires
    wave, _ = factory.make_inputs(3, 16000)
    stft = STFT(n_fft=win, hop_length=160, win_length=win)
    with torch.no_grad():
        re, im = stft(wave[:, 0])
        mag_ref, cos_ref, sin_ref = magphase(re, im)
    mag, cos, sin = [t.cpu() for t in multires.calculate_stft_components(wave.cuda(), win, 160, win)]
    assert mag.shape == mag_ref.shape == (3, 1, 101, win // 2 + 1)
    assert factory.max_rel_err(mag_ref, mag) <= 1e-4
    assert factory.max_rel_err(mag_ref * cos_ref, mag * cos) <= 1e-4
    assert factory.max_rel_err(mag_ref * sin_ref, mag * sin) <= 1e-4
    assert float(mag[1].abs().max()) == 0.0 and float(cos[1].abs().max()) == 0.0      # silent clip: magphase semantics
    out = multires.multires_stft(wave.cuda())
    assert sorted(out) == [256, 512, 2048]


@pytest.mark.gpu
def test_audiosep_separate_end_to_end():
    from lass_b200.models.audiosep import AudioSep
    from lass_b200.models.clap_standin import RandomInitCLAPTextEncoder
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    ss = ResUNet30(1, 1, 512).eval().cuda()
    enc = RandomInitCLAPTextEncoder(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128).cuda()
    model = AudioSep(ss_model=ss, query_encoder=enc)
    mix, _ = factory.make_inputs(2, 16000, edge_clips=False)
    out = model.separate(mix.cuda(), ["a dog barking", "rain"])
    assert out.shape == (2, 1, 16000) and bool(torch.isfinite(out).all())
