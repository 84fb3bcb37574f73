"""``AudioSep``-named shell around the separator + query encoder without Lightning.

The reference's ``AudioSep`` is a ``pl.LightningModule`` whose ``forward`` is a no-op (``models/audiosep.py:49-50``); its
useful contracts are

* inference: ``ss_model(input_dict)['waveform']`` with the condition from ``query_encoder.get_query_embed``
  (``dcase_evaluator.py:93-104``) -> ``separate``;
* training: ``training_step(batch_data_dict, batch_idx)`` (``models/audiosep.py:52-113``: mixer -> query embedding ->
  ``ss_model.train()`` forward -> ``l1_wav``) and ``configure_optimizers`` (``:118-145``: AdamW(amsgrad) + per-step
  ``LambdaLR``).  Both exist here with the same names and argument meaning; the trainer around them (Lightning, DDP wrapper,
  logging, checkpoint callbacks) is outside the hot path.  ``fused_training_step`` is the fast equivalent of
  ``training_step`` + ``backward`` + DDP all-reduce + ``optimizer.step`` + ``scheduler.step`` in one call on the
  training-step engine (``lass_b200/training.py``).
"""
import random

import torch
import torch.nn as nn

from .resunet import ResUNet30


def get_model_class(model_type):
    """reference ``models/audiosep.py:148-154``"""
    if model_type == "ResUNet30":
        return ResUNet30
    raise NotImplementedError


def l1(output, target):
    return torch.mean(torch.abs(output - target))


def l1_wav(output_dict, target_dict):
    """reference ``losses.py:4-9``"""
    return l1(output_dict["segment"], target_dict["segment"])


def get_loss_function(loss_type):
    """reference ``losses.py:12-17``"""
    if loss_type == "l1_wav":
        return l1_wav
    raise NotImplementedError("Error!")


class AudioSep(nn.Module):
    def __init__(self, ss_model: nn.Module = None, waveform_mixer=None, query_encoder: nn.Module = None, loss_function=None,
                 optimizer_type: str = None, learning_rate: float = None, lr_lambda_func=None, use_text_ratio: float = 1.0):
        super().__init__()
        self.ss_model = ss_model
        self.waveform_mixer = waveform_mixer
        self.query_encoder = query_encoder
        self.query_encoder_type = getattr(query_encoder, "encoder_type", None)
        self.use_text_ratio = use_text_ratio
        self.loss_function = loss_function
        self.optimizer_type = optimizer_type
        self.learning_rate = learning_rate
        self.lr_lambda_func = lr_lambda_func
        self.global_step = 0

    def forward(self, x):
        pass                                         # reference models/audiosep.py:49-50

    # ------------------------------------------------------------------ inference (dcase_evaluator.py:93-104)
    @torch.no_grad()
    def separate(self, mixture: torch.Tensor, text):
        """mixture (B, 1, L) on the separator's device, text: list of B captions -> waveform (B, 1, L)."""
        conditions = self.query_encoder.get_query_embed(modality="text", text=text)
        input_dict = {"mixture": mixture, "condition": conditions.to(mixture.device)}
        return self.ss_model(input_dict)["waveform"]

    # ------------------------------------------------------------------ training (models/audiosep.py:52-145)
    def _prepare(self, batch_data_dict, batch_idx):
        random.seed(batch_idx)                       # "[important] fix random seeds across devices" (:68-69)
        batch_audio_text_dict = batch_data_dict["audio_text"]
        batch_text = batch_audio_text_dict["text"]
        batch_audio = batch_audio_text_dict["waveform"]
        if self.waveform_mixer is None:
            raise RuntimeError("AudioSep.training_step needs a waveform_mixer (waveforms -> (mixtures, segments))")
        mixtures, segments = self.waveform_mixer(waveforms=batch_audio)
        if self.query_encoder_type != "CLAP":
            raise NotImplementedError("only CLAP-type query encoders (reference models/audiosep.py:81-87)")
        with torch.no_grad():
            conditions = self.query_encoder.get_query_embed(modality="hybird", text=batch_text, audio=segments.squeeze(1),
                                                            use_text_ratio=self.use_text_ratio)
        input_dict = {"mixture": mixtures[:, None, :].squeeze(1), "condition": conditions}
        return input_dict, segments

    def training_step(self, batch_data_dict, batch_idx):
        """Same contract as the reference: returns the loss tensor (connected to ``ss_model``'s parameters through the
        autograd bridge of the training-step engine), the caller runs ``backward`` / ``optimizer.step``."""
        input_dict, segments = self._prepare(batch_data_dict, batch_idx)
        self.ss_model.train()
        sep_segment = self.ss_model(input_dict)["waveform"].squeeze()
        return self.loss_function({"segment": sep_segment}, {"segment": segments.squeeze(1).squeeze()})

    def fused_training_step(self, batch_data_dict, batch_idx, process_group=None, sync_batchnorm=None):
        """``training_step`` + backward + gradient all-reduce (NCCL, when torch.distributed is initialised) + AdamW(amsgrad)
        + the per-step LambdaLR factor, as ONE kernel sequence without autograd.  Only ``l1_wav`` / ``AdamW`` (the reference's
        configuration, ``config/audiosep_base.yaml``).  ``sync_batchnorm=True`` = the reference Trainer's flag of the same name
        (BatchNorm statistics over all ranks).  Returns the loss of this rank as a 0-d tensor."""
        if self.loss_function is not l1_wav or self.optimizer_type != "AdamW":
            raise NotImplementedError("the fused step implements l1_wav + AdamW(amsgrad=True)")
        input_dict, segments = self._prepare(batch_data_dict, batch_idx)
        self.ss_model.train()
        scale = self.lr_lambda_func(self.global_step) if self.lr_lambda_func is not None else 1.0
        with torch.no_grad():
            loss = self.ss_model.train_engine(sync_batchnorm, process_group).training_step(input_dict["mixture"], input_dict["condition"],
                                                              segments.reshape(input_dict["mixture"].shape),
                                                              lr=self.learning_rate * scale, process_group=process_group)
        self.global_step += 1
        return loss

    def configure_optimizers(self):
        """reference ``models/audiosep.py:118-145`` (torch optimizer + scheduler objects for callers that step themselves)."""
        if self.optimizer_type != "AdamW":
            raise NotImplementedError
        optimizer = torch.optim.AdamW(params=self.ss_model.parameters(), lr=self.learning_rate, betas=(0.9, 0.999), eps=1e-08,
                                      weight_decay=0.0, amsgrad=True)
        scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, self.lr_lambda_func)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "step", "frequency": 1}}
