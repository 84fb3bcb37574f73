"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/models/resunet.py through the torchlibrosa restatement) in the build container.

    python -m oracle.make_golden

The fixtures pin (i) the travelling oracle ``oracle/resunet_oracle.py`` and (ii) the B200 path to outputs of the
reference itself; they are small (a few hundred kB) and committed.  Weights are not stored: they are regenerated
from ``oracle.factory.fill_state_dict`` (key-seeded), whose per-key checksums ARE stored so a mismatch is loud.
"""
import json
import os

import numpy as np
import torch

from . import factory
from .reference_loader import import_reference_resunet

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_mod = import_reference_resunet()
    torch.manual_seed(0)
    net = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512).eval()
    sd = factory.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)

    # weight-factory checksums (float64 sums, every key)
    sums = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
    with open(os.path.join(GOLDEN_DIR, "factory_seed0_checksums.json"), "w") as f:
        json.dump(sums, f, indent=0, sort_keys=True)

    # whole forward, reference shapes (n_fft 1024 / hop 160), three clips incl. the two edge clips
    B, L = 3, 24000
    mix, cond = factory.make_inputs(B, L, seed=1234)
    with torch.no_grad():
        wav = net({"mixture": mix, "condition": cond})["waveform"]
        # intermediates through the reference's own methods
        mag, cos, sin = net.base.wav_to_spectrogram_phase(mix)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "resunet30_fwd_b3_l24000.npz"),
                        waveform=wav.numpy(), mag=mag.numpy().astype(np.float32),
                        cos_clip0=cos[0].numpy(), sin_clip0=sin[0].numpy(),
                        meta=np.array([B, L, 1024, 160, 1234, 0]))

    # chunk_inference of the reference (RATE is hard-coded to 32000 there: models/resunet.py:661);
    # 7.2 s at "32 kHz" gives two full windows plus a truncated tail window
    Lc = 230000
    mixc, condc = factory.make_inputs(1, Lc, seed=77, edge_clips=False)
    with torch.no_grad():
        out = net.chunk_inference({"mixture": mixc, "condition": condc})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "chunk_inference_l230000.npz"),
                        waveform=out.astype(np.float32), meta=np.array([1, Lc, 77]))

    # spectral round trip of the reference's STFT / ISTFT objects (mask = identity) for both shape sets
    from torchlibrosa.stft import ISTFT, STFT
    for n_fft, hop in ((1024, 160), (2048, 320), (512, 160), (256, 160)):
        stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=True, pad_mode="reflect")
        istft = ISTFT(n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=True, pad_mode="reflect")
        wave, _ = factory.make_inputs(2, 8000, seed=5, edge_clips=False)
        with torch.no_grad():
            re, im = stft(wave[:, 0])
            back = istft(re, im, 8000)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "stft_%d_%d_l8000.npz" % (n_fft, hop)),
                            real=re.numpy(), imag=im.numpy(), roundtrip=back.numpy())
    make_train_golden(ref_mod)
    print("golden fixtures written to", GOLDEN_DIR)


def make_train_golden(ref_mod=None):
    """Training step of the UNMODIFIED reference (``.train()`` forward, reference models/audiosep.py:99-100, + l1_wav,
    losses.py:4-9, + backward): loss, per-parameter gradient norms / first entries, two updated running statistics.
    Full gradients are 100 MB, so only this summary is stored; the travelling oracle reproduces them in full."""
    from . import train_oracle
    ref_mod = ref_mod or import_reference_resunet()
    torch.manual_seed(0)
    net = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    sd = factory.fill_state_dict(net.state_dict(), seed=0)
    net.load_state_dict(sd)
    net.train()
    B, L = 2, 16000
    mix, cond = factory.make_inputs(B, L, seed=1234, edge_clips=False)
    tgt, _ = factory.make_inputs(B, L, seed=4321, edge_clips=False)
    tgt = 0.5 * tgt
    out = net({"mixture": mix, "condition": cond})["waveform"]
    loss = torch.mean(torch.abs(out.squeeze() - tgt.squeeze()))
    loss.backward()
    keys = [k for k, p in net.named_parameters() if p.requires_grad and p.grad is not None]
    assert all(not train_oracle.is_dead_key(k) for k in keys)
    grads = dict((k, p.grad) for k, p in net.named_parameters() if p.grad is not None)
    new_sd = net.state_dict()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_step_b2_l16000.npz"),
                        loss=np.float64(float(loss.detach())), keys=np.array(keys),
                        grad_norms=np.array([float(grads[k].double().norm()) for k in keys]),
                        grad_first=np.array([float(grads[k].reshape(-1)[0]) for k in keys]),
                        grad_absmax=np.array([float(grads[k].abs().max()) for k in keys]),
                        bn0_running_mean=new_sd["base.bn0.running_mean"].numpy(),
                        enc3_bn2_running_var=new_sd["base.encoder_block3.conv_block1.bn2.running_var"].numpy(),
                        waveform=out.detach().numpy(), meta=np.array([B, L, 1234, 4321]))


def make_mixer_golden():
    """``SegmentMixer`` of the UNMODIFIED reference (data/waveform_mixers.py:9-62) on a seeded batch with ``random.seed(7)``."""
    import random
    from .reference_loader import import_reference_mixers
    from .segment_mixer_oracle import make_waveforms
    ref = import_reference_mixers()
    B, L, max_mix_num, lower_db, higher_db, seed = 6, 4000, 4, -10, 10, 7
    wave = make_waveforms(B, L, seed=3)
    random.seed(seed)
    mixture, segment = ref.SegmentMixer(max_mix_num=max_mix_num, lower_db=lower_db, higher_db=higher_db)(wave)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "segment_mixer_b6_l4000.npz"), mixture=mixture.numpy(), segment=segment.numpy(),
                        meta=np.array([B, L, max_mix_num, lower_db, higher_db, seed, 3]))


if __name__ == "__main__":
    import sys
    if "--mixer-only" in sys.argv:
        make_mixer_golden()
    else:
        main()
        make_mixer_golden()
