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
    import inspect
    from lass_b200.models.audiosep import AudioSep, get_loss_function, get_model_class, l1_wav
    from lass_b200.models.resunet import ResUNet30
    assert get_model_class("ResUNet30") is ResUNet30
    m = AudioSep(ss_model=torch.nn.Identity(), query_encoder=None)
    assert m.forward(torch.zeros(1)) is None                                   # reference forward is `pass`
    assert hasattr(m, "ss_model") and hasattr(m, "query_encoder") and hasattr(m, "use_text_ratio")
    # constructor keywords in the reference's order (models/audiosep.py:15-24)
    assert list(inspect.signature(AudioSep.__init__).parameters)[1:] == [
        "ss_model", "waveform_mixer", "query_encoder", "loss_function", "optimizer_type", "learning_rate", "lr_lambda_func",
        "use_text_ratio"]
    assert get_loss_function("l1_wav") is l1_wav
    with pytest.raises(NotImplementedError):
        get_loss_function("l2")
    a, b = torch.randn(3, 100), torch.randn(3, 100)
    assert torch.equal(l1_wav({"segment": a}, {"segment": b}), (a - b).abs().mean())
    m = AudioSep(ss_model=ResUNet30(1, 1, 512), query_encoder=None, optimizer_type="AdamW", learning_rate=1e-3,
                 lr_lambda_func=lambda step: 1.0)
    opt = m.configure_optimizers()
    o = opt["optimizer"]
    assert isinstance(o, torch.optim.AdamW) and o.defaults["amsgrad"] and o.defaults["weight_decay"] == 0.0
    assert opt["lr_scheduler"]["interval"] == "step" and opt["lr_scheduler"]["frequency"] == 1


def test_lr_schedulers_match_the_reference():
    """optimizers/lr_schedulers.py: same factors as the unmodified reference functions over a step sweep."""
    from lass_b200 import lr_schedulers
    from oracle import reference_loader
    if not reference_loader.reference_available():
        pytest.skip("reference tree not present")
    import importlib
    import sys
    if reference_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, reference_loader.REFERENCE_ROOT)
    ref = importlib.import_module("optimizers.lr_schedulers")
    for name in ("constant_warm_up", "linear_warm_up"):
        mine = lr_schedulers.get_lr_lambda(name, warm_up_steps=100, reduce_lr_steps=1000)
        theirs = ref.get_lr_lambda(name, warm_up_steps=100, reduce_lr_steps=1000)
        for step in list(range(0, 450, 7)) + [99, 100, 101, 199, 200, 299, 300, 999, 1000, 1001, 25000]:
            assert mine(step) == theirs(step), (name, step)
    with pytest.raises(NotImplementedError):
        lr_schedulers.get_lr_lambda("cosine", warm_up_steps=1, reduce_lr_steps=1)


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
    out = multires.multires_stft(wave.cuda())           # all three resolutions in ONE K1 launch: same tiles, same bits
    assert sorted(out) == [256, 512, 2048]
    for got, single in zip(out[win], (mag, cos, sin)):
        assert torch.equal(got.cpu(), single)


@pytest.mark.gpu
def test_multires_front_end_at_config5_size():
    """BASELINE config 5 size: 32 clips x 10 s, the three resolutions 256 / 512 / 2048 at hop 160 against the restated
    torchlibrosa STFT + magphase on four clips of the batch (first, interior, silent, full-scale sine)."""
    from lass_b200 import multires
    wave, _ = factory.make_inputs(32, 160000, seed=6)
    out = multires.multires_stft(wave.cuda())
    idx = [0, 17, 30, 31]
    for win in (256, 512, 2048):
        mag, cos, sin = [t[idx].cpu() for t in out[win]]
        assert out[win][0].shape == (32, 1, 1001, win // 2 + 1)
        stft = STFT(n_fft=win, hop_length=160, win_length=win)
        with torch.no_grad():
            re, im = stft(wave[idx, 0])
            mag_ref, cos_ref, sin_ref = magphase(re, im)
        assert factory.max_rel_err(mag_ref, mag) <= 1e-4
        assert factory.max_rel_err(mag_ref * cos_ref, mag * cos) <= 1e-4
        assert factory.max_rel_err(mag_ref * sin_ref, mag * sin) <= 1e-4
        assert float(mag[2].abs().max()) == 0.0


@pytest.mark.gpu
def test_audiosep_separate_matches_oracle_with_the_same_embedding():
    """Config 3 contract (dcase_evaluator.py:93-104): conditions from the query encoder's text tower, then the separator.
    The oracle forward is fed the SAME embeddings; bar 40 dB like every whole-forward test."""
    from helpers import build_module, snr_ok
    from lass_b200.models.audiosep import AudioSep
    from lass_b200.models.clap_standin import RandomInitCLAPTextEncoder
    from oracle import resunet_oracle as O
    ss, sd = build_module(device="cuda")
    enc = RandomInitCLAPTextEncoder(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128).cuda()
    model = AudioSep(ss_model=ss, query_encoder=enc)
    mix, _ = factory.make_inputs(3, 16000, edge_clips=False)
    text = ["a dog barking", "rain", "a dog barking"]
    out = model.separate(mix.cuda(), text)
    assert out.shape == (3, 1, 16000)
    cond = enc.get_query_embed(modality="text", text=text).cpu()
    ref = O.resunet30_forward(sd, mix, cond)
    snr_ok(ref, out.cpu(), 40.0)


@pytest.mark.gpu
def test_audiosep_training_step_and_fused_step():
    """models/audiosep.py:52-145 through the shell: training_step returns a loss connected to the parameters (autograd bridge);
    fused_training_step performs the same optimisation step in one call -- both start from the same weights and must agree
    on the loss and move the parameters the same way."""
    from functools import partial
    from lass_b200 import lr_schedulers
    from lass_b200.models.audiosep import AudioSep, get_loss_function
    from lass_b200.models.clap_standin import RandomInitCLAPTextEncoder
    from lass_b200.models.resunet import ResUNet30

    from lass_b200.data.waveform_mixers import SegmentMixer
    mixer = SegmentMixer(max_mix_num=2, lower_db=-10, higher_db=10)       # the reference's own step in front (train.py:217-221), on the GPU;
                                                                         # random.seed(batch_idx) inside the step pins its draws

    enc = RandomInitCLAPTextEncoder(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128).cuda()
    wave, _ = factory.make_inputs(4, 16000, edge_clips=False)
    batch = {"audio_text": {"text": ["a", "b", "c", "d"], "waveform": wave.cuda(), "modality": "audio_text"}}
    lam = lr_schedulers.get_lr_lambda("constant_warm_up", warm_up_steps=10, reduce_lr_steps=100)
    losses, moved = [], []
    for fused in (False, True):
        torch.manual_seed(0)
        ss = ResUNet30(1, 1, 512).cuda()
        model = AudioSep(ss_model=ss, waveform_mixer=mixer, query_encoder=enc, loss_function=get_loss_function("l1_wav"),
                         optimizer_type="AdamW", learning_rate=1e-3, lr_lambda_func=lam)
        w0 = ss.base.after_conv.weight.detach().clone()
        if fused:
            loss = model.fused_training_step(batch, 0)
        else:
            opt = model.configure_optimizers()
            loss = model.training_step(batch, 0)
            loss.backward()
            opt["optimizer"].step()
            opt["lr_scheduler"]["scheduler"].step()
        losses.append(float(loss))
        moved.append((ss.base.after_conv.weight.detach() - w0).clone())
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0])
    # first AdamW step: |delta| = lr * 0.001 (constant_warm_up plateau) for every element with a non-negligible gradient
    assert float(moved[0].abs().max()) == pytest.approx(1e-6, rel=5e-2)          # (quantised by the fp32 ulp of the weights)
    assert float((moved[0] - moved[1]).abs().max()) <= 2e-7


def test_precomputed_shard_format_matches_the_reference_writer(tmp_path):
    """The .pt shard layout of reference scripts/precompute_stfts.py:60-123 / :596-622 (list of item dicts, CPU tensors, file
    name) written by lass_b200.multires.save_batch_precomputed_data and read back; no GPU needed for the format itself."""
    from lass_b200 import multires
    wins, T, L = [256, 512, 2048], 11, 1600
    items = []
    for k in range(3):
        st = {src: {w: tuple(torch.randn(1, 1, T, w // 2 + 1) for _ in range(3)) for w in wins} for src in ("mixture", "segment")}
        items.append({"stfts": st, "target_waveform": torch.randn(1, L), "text": "caption %d" % k,
                      "mixture_component_texts": ["caption %d" % k, "other"], "stft_common_params": multires.shard_common_params(160),
                      "stft_win_lengths": wins})
    assert multires.save_batch_precomputed_data(tmp_path, 7, []) == 0 and not list(tmp_path.iterdir())
    assert multires.save_batch_precomputed_data(tmp_path, 7, items) == 3
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == ["batch_000007.pt"]
    back = multires.load_batch_precomputed_data(tmp_path / files[0])
    assert isinstance(back, list) and len(back) == 3
    for a, b in zip(items, back):
        assert sorted(b) == ["mixture_component_texts", "stft_common_params", "stft_win_lengths", "stfts", "target_waveform", "text"]
        assert b["text"] == a["text"] and b["mixture_component_texts"] == a["mixture_component_texts"]
        assert b["stft_common_params"] == {"hop_length": 160, "window": "hann", "center": True, "pad_mode": "reflect"}
        assert b["stft_win_lengths"] == wins and sorted(b["stfts"]) == ["mixture", "segment"]
        for src in ("mixture", "segment"):
            assert sorted(b["stfts"][src]) == wins
            for w in wins:
                tup = b["stfts"][src][w]
                assert isinstance(tup, tuple) and len(tup) == 3
                for x, y in zip(tup, a["stfts"][src][w]):
                    assert x.device.type == "cpu" and x.shape == (1, 1, T, w // 2 + 1) and torch.equal(x, y)
        assert torch.equal(b["target_waveform"], a["target_waveform"])


@pytest.mark.gpu
def test_shard_items_hold_the_reference_stft_components():
    """build_shard_items: item k's tuples are the [k:k+1] slices of calculate_stft_components (reference
    scripts/precompute_stfts.py:573-607), checked against the torchlibrosa restatement."""
    from lass_b200 import multires
    from oracle.torchlibrosa import stft as tl
    g = torch.Generator().manual_seed(3)
    mix, seg = 0.1 * torch.randn(3, 1, 16000, generator=g), 0.1 * torch.randn(3, 1, 16000, generator=g)
    items = multires.build_shard_items(mix.cuda(), seg.cuda(), ["a", "b", "c"], [["a", "x"], ["b"], ["c", "y", "z"]])
    assert len(items) == 3 and items[1]["text"] == "b" and items[2]["mixture_component_texts"] == ["c", "y", "z"]
    for w in (256, 512, 2048):
        ext = tl.STFT(n_fft=w, hop_length=160, win_length=w, window="hann", center=True, pad_mode="reflect", freeze_parameters=True)
        for src, wave in (("mixture", mix), ("segment", seg)):
            real, imag = ext(wave[:, 0])
            mag, cos, sin = tl.magphase(real, imag)
            for k in range(3):
                m, c, s_ = (t.cpu() for t in items[k]["stfts"][src][w])
                assert m.shape == (1, 1, 101, w // 2 + 1)
                assert float((m - mag[k:k + 1]).abs().max()) <= 1e-4 * float(mag.abs().max())
                big = mag[k:k + 1] > 1e-3 * float(mag.max())
                assert float(((c - cos[k:k + 1]).abs() * big).max()) <= 2e-3 and float(((s_ - sin[k:k + 1]).abs() * big).max()) <= 2e-3
        assert torch.equal(items[0]["target_waveform"].cpu(), seg[0])
