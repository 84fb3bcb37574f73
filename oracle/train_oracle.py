"""TEST INFRASTRUCTURE ONLY — fp32 CPU restatement of the reference's training step for the separation path.

Restates, on top of ``oracle/resunet_oracle.py`` (the functional forward, here with ``train=True``):

* ``AudioSep.training_step``        reference ``models/audiosep.py:52-113``: ``ss_model.train()`` (``:99``, BatchNorm batch
                                    statistics + running-stat update with momentum 0.01), forward (``:100``), loss (``:109``)
* ``l1_wav``                        reference ``losses.py:4-9``: mean |output - target|
* ``configure_optimizers``          reference ``models/audiosep.py:118-145``: AdamW(betas 0.9/0.999, eps 1e-8, weight_decay 0,
                                    amsgrad=True) + LambdaLR
* ``constant_warm_up``              reference ``optimizers/lr_schedulers.py:34-67`` (restated in lass_b200/lr_schedulers.py too)

Gradients come from torch autograd over the functional forward, i.e. the same aten ops in the same order as the reference
module.  Pinned to the UNMODIFIED reference (``ResUNet30(...).train()`` + ``loss.backward()``) by
``tests/test_train_oracle.py`` in the build container, and by the golden summary ``tests/golden/train_step_b2_l16000.npz``
(loss, per-parameter gradient norms / sampled entries, updated running statistics) everywhere.
"""
from typing import Dict

import torch

from . import resunet_oracle

BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


def is_trainable_key(key: str) -> bool:
    """Keys of ``ss_model.parameters()`` with ``requires_grad`` (the frozen DFT matrices and the BatchNorm buffers are not)."""
    leaf = key.rsplit(".", 1)[-1]
    if leaf in BUFFER_LEAVES or leaf == "ola_window":
        return False
    return ".stft." not in key and ".istft." not in key


# parameters that never receive a gradient: the six unused decoder bn2 and their FiLM beta2 linears
def is_dead_key(key: str) -> bool:
    if key.startswith("base.decoder_block") and ".bn2." in key and ".conv_block2." not in key:
        return True
    return key.startswith("film.decoder_block") and "->beta2." in key and "conv_block2" not in key


def l1_wav(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """reference losses.py:4-9"""
    return torch.mean(torch.abs(output - target))


def training_forward_backward(sd: Dict[str, torch.Tensor], mixture, condition, target, hop: int = 160):
    """One forward + backward of the reference's training step (no optimizer).

    sd: reference-keyed state dict (not modified).  mixture / target (B, 1, L), condition (B, 512).
    Returns (loss float, waveform (B, 1, L), grads {key: tensor} for every live trainable key, new_buffers {key: tensor}).
    """
    work = {}
    leaves = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if is_trainable_key(k):
            t.requires_grad_(True)
            leaves[k] = t
        work[k] = t
    wave = resunet_oracle.resunet30_forward_impl(work, mixture, condition, hop=hop, train=True)
    loss = l1_wav(wave.squeeze(), target.squeeze())
    loss.backward()
    grads = {k: t.grad.detach() for k, t in leaves.items() if t.grad is not None}
    buffers = {k: v.detach() for k, v in work.items() if k.rsplit(".", 1)[-1] in BUFFER_LEAVES}
    return float(loss.detach()), wave.detach(), grads, buffers


def adamw_amsgrad_step(p, g, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """In-place single-tensor AdamW(amsgrad=True) update, same op order as ``torch.optim.adamw`` (single-tensor path):
    reference models/audiosep.py:122-130.  ``step`` counts from 1."""
    p.mul_(1.0 - lr * weight_decay)
    m.lerp_(g, 1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    torch.maximum(vmax, v, out=vmax)
    denom = (vmax.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
