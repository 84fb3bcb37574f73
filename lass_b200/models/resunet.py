"""B200-native drop-in for the reference's ``models/resunet.py``.

Same public surface as the reference:

* ``ResUNet30(input_channels, output_channels, condition_size)`` (reference ``models/resunet.py:621-653``)
  with ``.base`` (``ResUNet30_Base``), ``.film`` (``FiLM``), ``.film_meta``, ``.forward({'mixture','condition'}) ->
  {'waveform'}`` and ``.chunk_inference``;
* identical parameter / buffer names, shapes and initialisation, so ``state_dict()`` / ``load_state_dict()``
  interoperate with reference checkpoints (332 keys, incl. the frozen DFT matrices and the dead
  ``decoder_blockN.bn2`` / ``film.decoder_blockN->beta2`` entries).

What differs is where the arithmetic runs: the sub-modules below only own parameters; ``forward`` hands the whole
computation to the sm_100a kernels behind the C ABI (``lass_b200.engine.Engine`` -> ``lass_resunet30_forward``).
There is no PyTorch / CPU fallback: without a CUDA device or without the built library the call raises.

Scope: ``input_channels == output_channels == 1`` (the reference's only shipped configuration,
``config/audiosep_base.yaml:25-28``).  ``.eval()``: the inference engine (``lass_b200/engine.py``).  ``.train()``: the
training-step engine (``lass_b200/training.py``): batch-statistics BatchNorm, hand-written backward, autograd bridge.
"""
import weakref
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .spectral import ISTFT, STFT

_ENGINES = weakref.WeakKeyDictionary()
_TRAIN_ENGINES = weakref.WeakKeyDictionary()


class _EpochMixin:
    """Bumps ``_lass_epoch`` whenever parameter OBJECTS may have been replaced (``.to()`` / ``.cuda()`` / ``.half()`` through
    ``_apply``, ``load_state_dict(assign=True)``), so the engine's cached parameter list (``Engine._version_key``) is rebuilt.
    In-place value changes are detected by the tensors' version counters and need no hook."""
    _lass_epoch = 0

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._lass_epoch += 1
        return out

    def _lass_bump_epoch(self, *_args):
        self._lass_epoch += 1


def init_layer(layer):
    """Xavier-uniform weight, zero bias (reference ``models/base.py:9-15``)."""
    nn.init.xavier_uniform_(layer.weight)
    if hasattr(layer, "bias") and layer.bias is not None:
        layer.bias.data.fill_(0.0)


def init_bn(bn):
    """Identity BatchNorm affine (reference ``models/base.py:18-21``)."""
    bn.bias.data.fill_(0.0)
    bn.weight.data.fill_(1.0)


class FiLM(_EpochMixin, nn.Module):
    """FiLM generator: one ``nn.Linear(condition_size, C)`` per (block, beta) named
    ``'encoder_block1->conv_block1->beta1'`` ... (reference ``models/resunet.py:10-81``).

    Inside ``ResUNet30.forward`` these linears are evaluated by kernel K2 as one GEMM with the BatchNorm shifts
    folded in.  Calling the module directly returns the reference's nested dict of ``(B, C, 1, 1)`` tensors, for
    callers that use ``ss_model.film(conditions=...)`` + ``ss_model.base(mixtures=, film_dict=)``
    (reference ``models/resunet.py:667-688``).
    """

    def __init__(self, film_meta, condition_size):
        super().__init__()
        self.register_load_state_dict_post_hook(lambda module, _keys: module._lass_bump_epoch())
        self.condition_size = condition_size
        self.modules, _ = self.create_film_modules(film_meta=film_meta, ancestor_names=[])

    def create_film_modules(self, film_meta, ancestor_names):
        modules = {}
        for module_name, value in film_meta.items():
            if isinstance(value, int):
                ancestor_names.append(module_name)
                unique_module_name = "->".join(ancestor_names)
                modules[module_name] = self.add_film_layer_to_module(
                    num_features=value, unique_module_name=unique_module_name)
            elif isinstance(value, dict):
                ancestor_names.append(module_name)
                modules[module_name], _ = self.create_film_modules(film_meta=value, ancestor_names=ancestor_names)
            ancestor_names.pop()
        return modules, ancestor_names

    def add_film_layer_to_module(self, num_features, unique_module_name):
        layer = nn.Linear(self.condition_size, num_features)
        init_layer(layer)
        self.add_module(name=unique_module_name, module=layer)
        return layer

    def forward(self, conditions):
        return self.calculate_film_data(conditions=conditions, modules=self.modules)

    def calculate_film_data(self, conditions, modules):
        film_data = {}
        for module_name, module in modules.items():
            if isinstance(module, nn.Module):
                film_data[module_name] = module(conditions)[:, :, None, None]
            elif isinstance(module, dict):
                film_data[module_name] = self.calculate_film_data(conditions, module)
        return film_data


class ConvBlockRes(nn.Module):
    """Parameters of one pre-activation residual block (reference ``models/resunet.py:84-165``)."""

    def __init__(self, in_channels, out_channels, kernel_size, momentum, has_film):
        super().__init__()
        padding = [kernel_size[0] // 2, kernel_size[1] // 2]
        self.bn1 = nn.BatchNorm2d(in_channels, momentum=momentum)
        self.bn2 = nn.BatchNorm2d(out_channels, momentum=momentum)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size, stride=(1, 1), dilation=(1, 1),
                               padding=padding, bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size, stride=(1, 1), dilation=(1, 1),
                               padding=padding, bias=False)
        if in_channels != out_channels:
            self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0))
            self.is_shortcut = True
        else:
            self.is_shortcut = False
        self.has_film = has_film
        init_bn(self.bn1)
        init_bn(self.bn2)
        init_layer(self.conv1)
        init_layer(self.conv2)
        if self.is_shortcut:
            init_layer(self.shortcut)


class EncoderBlockRes1B(nn.Module):
    """``conv_block1`` + average pooling (reference ``models/resunet.py:168-198``)."""

    def __init__(self, in_channels, out_channels, kernel_size, downsample, momentum, has_film):
        super().__init__()
        self.conv_block1 = ConvBlockRes(in_channels, out_channels, kernel_size, momentum, has_film)
        self.downsample = downsample


class DecoderBlockRes1B(nn.Module):
    """Transposed conv (kernel = stride) + concat + ``conv_block2`` (reference ``models/resunet.py:201-264``).
    ``bn2`` exists in the reference and in checkpoints but is never used in its forward (``:230`` vs ``:240-264``)."""

    def __init__(self, in_channels, out_channels, kernel_size, upsample, momentum, has_film):
        super().__init__()
        self.kernel_size = kernel_size
        self.stride = upsample
        self.conv1 = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=self.stride, stride=self.stride,
                                        padding=(0, 0), bias=False, dilation=(1, 1))
        self.bn1 = nn.BatchNorm2d(in_channels, momentum=momentum)
        self.conv_block2 = ConvBlockRes(out_channels * 2, out_channels, kernel_size, momentum, has_film)
        self.bn2 = nn.BatchNorm2d(in_channels, momentum=momentum)
        self.has_film = has_film
        init_bn(self.bn1)
        init_layer(self.conv1)


_ENCODERS = (  # name, cin, cout, downsample   (reference models/resunet.py:315-370)
    ("encoder_block1", 32, 32, (2, 2)), ("encoder_block2", 32, 64, (2, 2)), ("encoder_block3", 64, 128, (2, 2)),
    ("encoder_block4", 128, 256, (2, 2)), ("encoder_block5", 256, 384, (2, 2)), ("encoder_block6", 384, 384, (1, 2)),
    ("conv_block7a", 384, 384, (1, 1)),
)
_DECODERS = (  # name, cin, cout, upsample     (reference models/resunet.py:371-418)
    ("decoder_block1", 384, 384, (1, 2)), ("decoder_block2", 384, 384, (2, 2)), ("decoder_block3", 384, 256, (2, 2)),
    ("decoder_block4", 256, 128, (2, 2)), ("decoder_block5", 128, 64, (2, 2)), ("decoder_block6", 64, 32, (2, 2)),
)


class ResUNet30_Base(_EpochMixin, nn.Module):
    """STFT -> bn0 -> UNet -> mask -> ISTFT (reference ``models/resunet.py:267-595``)."""

    def __init__(self, input_channels, output_channels, window_size=1024, hop_size=160):
        super().__init__()
        self.register_load_state_dict_post_hook(lambda module, _keys: module._lass_bump_epoch())
        # reference models/resunet.py:271-282 hard-codes 1024 / 160; the kernels are parametric (BASELINE config 2
        # uses 2048 / 320), so the sizes are constructor keywords with the reference's values as defaults.
        momentum = 0.01
        self.window_size = window_size
        self.hop_size = hop_size
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.target_sources_num = 1
        self.K = 3
        self.time_downsample_ratio = 2 ** 5

        self.stft = STFT(n_fft=window_size, hop_length=hop_size, win_length=window_size, window="hann", center=True,
                         pad_mode="reflect", freeze_parameters=True)
        self.istft = ISTFT(n_fft=window_size, hop_length=hop_size, win_length=window_size, window="hann",
                           center=True, pad_mode="reflect", freeze_parameters=True)
        self.bn0 = nn.BatchNorm2d(window_size // 2 + 1, momentum=momentum)
        self.pre_conv = nn.Conv2d(input_channels, 32, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), bias=True)
        for name, cin, cout, down in _ENCODERS:
            setattr(self, name, EncoderBlockRes1B(cin, cout, (3, 3), down, momentum, True))
        for name, cin, cout, up in _DECODERS:
            setattr(self, name, DecoderBlockRes1B(cin, cout, (3, 3), up, momentum, True))
        self.after_conv = nn.Conv2d(32, output_channels * self.K, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                                    bias=True)
        init_bn(self.bn0)
        init_layer(self.pre_conv)
        init_layer(self.after_conv)

    def _get_engine(self, film):
        """Engine (packed weights + plans) for this module, with or without the FiLM generator fused in.
        Kept outside the module's attributes so that copy / pickle / state_dict never see ctypes handles."""
        from ..engine import Engine
        per_base = _ENGINES.setdefault(self, {})
        key = id(film) if film is not None else None
        eng = per_base.get(key)
        if eng is None:
            eng = per_base[key] = Engine(self, film)
        return eng

    def forward(self, mixtures, film_dict):
        """mixtures (B, 1, L) fp32 + the nested FiLM dict of ``FiLM.forward`` -> ``{'waveform': (B, 1, L)}``
        (reference ``models/resunet.py:522-595``)."""
        engine = self._get_engine(None)
        return {"waveform": engine.forward_film_dict(mixtures, film_dict)}


def get_film_meta(module):
    """Same traversal as reference ``models/resunet.py:598-618``."""
    film_meta = {}
    if hasattr(module, "has_film"):
        if module.has_film:
            film_meta["beta1"] = module.bn1.num_features
            film_meta["beta2"] = module.bn2.num_features
        else:
            film_meta["beta1"] = 0
            film_meta["beta2"] = 0
    for child_name, child_module in module.named_children():
        child_meta = get_film_meta(child_module)
        if len(child_meta) > 0:
            film_meta[child_name] = child_meta
    return film_meta


class ResUNet30(nn.Module):
    def __init__(self, input_channels, output_channels, condition_size, window_size=1024, hop_size=160):
        super().__init__()
        self.base = ResUNet30_Base(input_channels=input_channels, output_channels=output_channels,
                                   window_size=window_size, hop_size=hop_size)
        self.film_meta = get_film_meta(module=self.base)
        self.film = FiLM(film_meta=self.film_meta, condition_size=condition_size)

    def forward(self, input_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """``{'mixture': (B, 1, L) fp32, 'condition': (B, condition_size) fp32} -> {'waveform': (B, 1, L) fp32}``
        (reference ``models/resunet.py:640-653``); all work on the current CUDA stream of the inputs' device."""
        mixtures = input_dict["mixture"]
        conditions = input_dict["condition"]
        if self.training:
            # reference models/audiosep.py:99-100: the module is called in .train() with autograd on.  BatchNorm runs on batch
            # statistics (running stats updated, momentum 0.01) and the returned waveform is connected to the parameters, so
            # loss.backward() / optimizer.step() work as with the reference (lass_b200/training.py)
            from .. import training
            if not mixtures.is_cuda:
                raise RuntimeError("lass_b200 has no CPU path: move the module and its inputs to a CUDA device")
            return {"waveform": training.train_forward(self.train_engine(), mixtures, conditions)}
        engine = self.base._get_engine(self.film)
        return {"waveform": engine.forward(mixtures, conditions)}

    def train_engine(self, sync_batchnorm=None, process_group=None):
        """The training-step engine of this module (flat fp32 parameter / gradient buffers; created on first use — the
        module's parameters become views of its flat buffer, values unchanged).  ``sync_batchnorm`` (when given) switches
        BatchNorm statistics over all ranks on or off — the reference's ``sync_batchnorm: True`` Trainer flag
        (config/audiosep_base.yaml:42, train.py:176,266-283)."""
        from .. import training
        eng = _TRAIN_ENGINES.get(self)
        if eng is None or eng.device != self.base.pre_conv.weight.device:
            eng = training.TrainEngine(self)
            _TRAIN_ENGINES[self] = eng
        if sync_batchnorm is not None:
            eng.sync_batchnorm = bool(sync_batchnorm)
        if process_group is not None:
            eng.process_group = process_group
        return eng

    @torch.no_grad()
    def chunk_inference(self, input_dict, rate=None):
        """Long-audio inference (reference ``models/resunet.py:655-714``): 5 s windows (1 s left / 3 s centre / 1 s
        right context) hopping by 3 s, stitched exactly like the reference — but all windows of equal length are
        stacked on the batch axis and separated in one pass on the GPU instead of a serial loop with a device-to-host
        copy per window.  The reference hard-codes ``RATE = 32000`` (``:661``); ``rate`` overrides it.  Returns a
        ``(1, L)`` numpy array like the reference."""
        rate = 32000 if rate is None else rate
        mixtures = input_dict["mixture"]
        conditions = input_dict["condition"]
        assert mixtures.shape[0] == 1, "chunk_inference assumes batch 1 (reference models/resunet.py:677)"
        NL, NC, NR = int(1.0 * rate), int(3.0 * rate), int(1.0 * rate)
        WINDOW = NL + NC + NR
        L = mixtures.shape[2]
        # enumerate the windows the reference's loop visits: (start, is_tail)
        starts, cur = [], 0
        while cur + WINDOW < L:
            starts.append(cur)
            cur += NC
        # NB the reference re-runs the window at `cur` inside the loop; only its last instance survives in the
        # stitched output for the region beyond the full windows, so one evaluation per distinct start suffices.
        engine = self.base._get_engine(self.film)
        out = torch.zeros(1, L, dtype=torch.float32, device=mixtures.device)
        full_starts = sorted(set(starts + ([s + NC for s in starts if s + NC + WINDOW <= L])))
        if full_starts:
            batch = torch.stack([mixtures[0, :, s:s + WINDOW] for s in full_starts], dim=0).contiguous()
            cond = conditions.expand(len(full_starts), -1).contiguous()
            sep = engine.forward(batch, cond)[:, 0]            # (n, WINDOW)
            by_start = {s: sep[i] for i, s in enumerate(full_starts)}
        else:
            by_start = {}
        cur = 0
        while cur + WINDOW < L:
            chunk = by_start[cur]
            if cur == 0:
                out[0, cur:cur + WINDOW - NR] = chunk[:-NR] if NR != 0 else chunk
            else:
                out[0, cur + NL:cur + WINDOW - NR] = chunk[NL:-NR] if NR != 0 else chunk[NL:]
            cur += NC
            if cur < L:
                if cur in by_start:
                    tail_chunk = by_start[cur]
                else:
                    seg = mixtures[:, :, cur:cur + WINDOW].contiguous()
                    tail_chunk = engine.forward(seg, conditions)[0, 0]
                seg_len = tail_chunk.shape[0]
                out[0, cur + NL:cur + seg_len] = tail_chunk[NL:]
        return out.cpu().numpy().astype(np.float64)
