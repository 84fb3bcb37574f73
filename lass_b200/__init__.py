"""lass_b200 — B200-native (sm_100a) implementation of the LASS/AudioSep separation hot path.

Drop-in surface (mirrors reference ``models/resunet.py``): ``lass_b200.models.resunet.ResUNet30``.
All GPU work goes through the C-ABI shared library ``lass_b200/_lib/liblass_b200.so`` (``include/lass_b200.h``),
bound with ctypes in ``lass_b200._cabi``.  There is no CPU fallback: using an op without the library or
without a CUDA device raises.
"""
__version__ = "0.1.0"
