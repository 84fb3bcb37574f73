"""TEST INFRASTRUCTURE ONLY — minimal stand-in package so that the unmodified reference
``models/resunet.py`` (``from torchlibrosa.stft import STFT, ISTFT, magphase``, line 6)
imports in a container where the real ``torchlibrosa==0.1.0`` wheel cannot be installed."""
__version__ = "0.1.0+oracle"
