"""K1 (STFT tcgen05 GEMM) and K5 (mask + iSTFT) against the oracle, the reference's golden vectors and fp64
torch.stft; plus size-independent properties at BASELINE config-2 size (64 x 10 s).
Tolerance (BASELINE.json north_star): STFT / iSTFT max relative error <= 1e-4 in fp32 (max|d| / max|ref| per tensor)."""
import pytest
import torch

from oracle import factory, resunet_oracle as O
from oracle.torchlibrosa.stft import ISTFT, STFT

from helpers import golden

pytestmark = pytest.mark.gpu
TOL = 1e-4
SHAPES = [(1024, 160), (2048, 320), (512, 160), (256, 160)]


def _setup(n_fft, hop):
    from lass_b200 import packing
    stft = STFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
    istft = ISTFT(n_fft=n_fft, hop_length=hop, win_length=n_fft)
    hi, lo = packing.pack_stft_basis(stft.conv_real.weight.data, stft.conv_imag.weight.data)
    window, tw = packing.istft_tables(n_fft, device="cuda")
    sd = {"base.stft.conv_real.weight": stft.conv_real.weight.data, "base.stft.conv_imag.weight": stft.conv_imag.weight.data,
          "base.istft.conv_real.weight": istft.conv_real.weight.data, "base.istft.conv_imag.weight": istft.conv_imag.weight.data,
          "base.istft.ola_window": istft.ola_window}
    return sd, hi.cuda(), lo.cuda(), window, tw


@pytest.mark.parametrize("n_fft,hop", SHAPES)
def test_stft_matches_reference_golden_and_fp64(n_fft, hop):
    from lass_b200 import ops
    sd, hi, lo, _, _ = _setup(n_fft, hop)
    g = golden("stft_%d_%d_l8000.npz" % (n_fft, hop))
    wave, _ = factory.make_inputs(2, 8000, seed=5, edge_clips=False)
    mag, cos, sin = [t.cpu() for t in ops.stft_fwd(wave[:, 0].contiguous().cuda(), hi, lo, n_fft, hop, 0)]
    re_ref, im_ref = torch.from_numpy(g["real"]), torch.from_numpy(g["imag"])
    mag_ref = torch.clamp(re_ref ** 2 + im_ref ** 2, 1e-10, float("inf")) ** 0.5
    assert factory.max_rel_err(mag_ref, mag) <= TOL
    assert factory.max_rel_err(re_ref, mag * cos) <= TOL
    assert factory.max_rel_err(im_ref, mag * sin) <= TOL
    assert float((cos ** 2 + sin ** 2 - 1).abs().max()) <= 1e-5
    spec = torch.stft(wave[:, 0].double(), n_fft, hop, n_fft, torch.hann_window(n_fft, periodic=True, dtype=torch.float64),
                      center=True, pad_mode="reflect", return_complex=True).transpose(1, 2)[:, None]
    assert factory.max_rel_err(spec.real, mag * cos) <= TOL
    assert factory.max_rel_err(spec.imag, mag * sin) <= TOL


@pytest.mark.parametrize("n_fft,hop", [(1024, 160), (2048, 320)])
def test_stft_edge_clips_and_ragged_lengths(n_fft, hop):
    """silent clip (both clamps), full-scale sine (peaky spectrum), lengths that are not multiples of hop."""
    from lass_b200 import ops
    sd, hi, lo, _, _ = _setup(n_fft, hop)
    for L in (n_fft // 2 + 1, 4 * n_fft + 37, 20000):
        wave, _ = factory.make_inputs(3, L)
        mag_ref, cos_ref, sin_ref = O.stft_mag_phase(sd, wave[:, 0], n_fft, hop)
        mag, cos, sin = [t.cpu() for t in ops.stft_fwd(wave[:, 0].contiguous().cuda(), hi, lo, n_fft, hop, 0)]
        assert mag.shape == mag_ref.shape == (3, 1, L // hop + 1, n_fft // 2 + 1)
        assert factory.max_rel_err(mag_ref, mag) <= TOL
        assert factory.max_rel_err(mag_ref * cos_ref, mag * cos) <= TOL
        assert factory.max_rel_err(mag_ref * sin_ref, mag * sin) <= TOL
        assert float((mag[1] - 1e-5).abs().max()) <= 1e-9 and float(cos[1].abs().max()) == 0.0   # silent clip


@pytest.mark.parametrize("n_fft,hop", SHAPES)
def test_mask_istft_matches_oracle(n_fft, hop):
    from lass_b200 import ops
    sd, hi, lo, window, tw = _setup(n_fft, hop)
    for L in (8000, 4 * n_fft + 37):
        wave, _ = factory.make_inputs(3, L)
        mag, cos, sin = [t.contiguous() for t in O.stft_mag_phase(sd, wave[:, 0], n_fft, hop)]
        g = torch.Generator().manual_seed(L)
        feat = torch.randn(3, 3, mag.shape[2], mag.shape[3], generator=g) * 2.0
        feat[..., -1] = 0.0            # zero-padded Nyquist column (models/resunet.py:573)
        ref = O.mask_to_wave(sd, feat, mag, cos, sin, L, n_fft, hop)[:, 0]
        out = ops.mask_istft(feat.cuda(), mag.cuda(), cos.cuda(), sin.cuda(), window, tw, n_fft, hop, L).cpu()
        assert factory.max_rel_err(ref, out) <= TOL
        # feat without the Nyquist column (the fused path's layout) gives the same result
        out2 = ops.mask_istft(feat[..., :-1].contiguous().cuda(), mag.cuda(), cos.cuda(), sin.cuda(), window, tw,
                              n_fft, hop, L).cpu()
        assert torch.equal(out, out2)


@pytest.mark.parametrize("n_fft,hop", [(1024, 160), (2048, 320), (512, 160)])
def test_full_size_round_trip_property(n_fft, hop):
    """BASELINE config 2 at full size (64 x 10 s; 512 / 160 is config 5's ISTFT, resunet_with_multistft.py:38-46): STFT ->
    identity mask -> iSTFT reproduces the input, and the
    transform is linear (size-independent properties; the oracle would take minutes at this size)."""
    from lass_b200 import ops
    sd, hi, lo, window, tw = _setup(n_fft, hop)
    B, L = 64, 160000
    g = torch.Generator().manual_seed(11)
    wave = (0.1 * torch.randn(B, L, generator=g)).cuda()
    mag, cos, sin = ops.stft_fwd(wave, hi, lo, n_fft, hop, 0)
    T, F = mag.shape[2], mag.shape[3]
    feat = torch.zeros(B, 3, T, F, device="cuda")
    feat[:, 0] = 40.0      # sigmoid -> 1
    feat[:, 1] = 40.0      # tanh -> 1  => mask phase (1, 0)
    # keep the Nyquist bin: use feat_F = F so the identity mask covers every bin
    back = ops.mask_istft(feat, mag, cos, sin, window, tw, n_fft, hop, L)
    assert float((back - wave).abs().max()) <= TOL * float(wave.abs().max())
    # linearity of the analysis: STFT(a x) = a STFT(x)
    mag2, _, _ = ops.stft_fwd(2.0 * wave[:4].contiguous(), hi, lo, n_fft, hop, 0)
    assert factory.max_rel_err(2.0 * mag[:4].cpu(), mag2.cpu()) <= TOL
