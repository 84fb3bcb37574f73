"""Learning-rate factor functions for ``LambdaLR`` with the reference's names and arguments
(``optimizers/lr_schedulers.py:5-101``; selected by ``config/audiosep_base.yaml`` as ``constant_warm_up``)."""
from functools import partial
from typing import Callable


def linear_warm_up(step: int, warm_up_steps: int, reduce_lr_steps: int) -> float:
    """Ramp 0 -> 1 over ``warm_up_steps``, then x0.9 every ``reduce_lr_steps`` (reference ``:5-31``)."""
    return step / warm_up_steps if step <= warm_up_steps else 0.9 ** (step // reduce_lr_steps)


def constant_warm_up(step: int, warm_up_steps: int, reduce_lr_steps: int) -> float:
    """Three plateaus 1e-3, 1e-2, 1e-1 of ``warm_up_steps`` each, then 1 (reference ``:34-67``; ``reduce_lr_steps`` is
    accepted and unused there too)."""
    if step < 0:
        return 1
    return (0.001, 0.01, 0.1, 1)[min(step // warm_up_steps, 3)]


_FUNCS = {"constant_warm_up": constant_warm_up, "linear_warm_up": linear_warm_up}


def get_lr_lambda(lr_lambda_type: str, **kwargs) -> Callable:
    """reference ``optimizers/lr_schedulers.py:70-101``"""
    if lr_lambda_type not in _FUNCS:
        raise NotImplementedError
    return partial(_FUNCS[lr_lambda_type], warm_up_steps=kwargs["warm_up_steps"], reduce_lr_steps=kwargs["reduce_lr_steps"])
