"""``AudioSep``-named shell around the separator + query encoder without Lightning.

The reference's ``AudioSep`` is a ``pl.LightningModule`` whose ``forward`` is a no-op (``models/audiosep.py:49-50``) and
whose useful inference contract is ``ss_model(input_dict)['waveform']`` with the condition coming from
``query_encoder.get_query_embed`` (``dcase_evaluator.py:93-104``).  This class keeps those attribute names and adds a
``separate`` convenience that performs exactly that sequence.  Training (``training_step``, optimizers) is not part of
this round.
"""
import torch
import torch.nn as nn

from .resunet import ResUNet30


def get_model_class(model_type):
    """reference ``models/audiosep.py:148-154``"""
    if model_type == "ResUNet30":
        return ResUNet30
    raise NotImplementedError


class AudioSep(nn.Module):
    def __init__(self, ss_model: nn.Module = None, query_encoder: nn.Module = None, waveform_mixer=None, loss_function=None,
                 optimizer_type: str = None, learning_rate: float = None, lr_lambda_func=None, use_text_ratio: float = 1.0):
        super().__init__()
        self.ss_model = ss_model
        self.query_encoder = query_encoder
        self.query_encoder_type = getattr(query_encoder, "encoder_type", None)
        self.waveform_mixer = waveform_mixer
        self.use_text_ratio = use_text_ratio
        self.loss_function = loss_function
        self.optimizer_type = optimizer_type
        self.learning_rate = learning_rate
        self.lr_lambda_func = lr_lambda_func

    def forward(self, x):
        pass                                         # reference models/audiosep.py:49-50

    @torch.no_grad()
    def separate(self, mixture: torch.Tensor, text):
        """mixture (B, 1, L) on the separator's device, text: list of B captions -> waveform (B, 1, L)."""
        conditions = self.query_encoder.get_query_embed(modality="text", text=text)
        input_dict = {"mixture": mixture, "condition": conditions.to(mixture.device)}
        return self.ss_model(input_dict)["waveform"]
