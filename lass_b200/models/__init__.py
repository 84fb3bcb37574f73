"""Drop-in mirror of the reference's ``models`` namespace for the separation hot path."""
from .resunet import ResUNet30, ResUNet30_Base, FiLM, get_film_meta  # noqa: F401


def get_model_class(model_type):
    """Same contract as reference ``models/audiosep.py:148-154``."""
    if model_type == "ResUNet30":
        return ResUNet30
    raise NotImplementedError
