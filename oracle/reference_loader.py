"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference from ``/root/reference``.

``/root/reference`` exists only in the build container (never on the GPU box), so this module is
used by (i) ``oracle/make_golden.py`` to generate ``tests/golden`` and (ii) the ``not gpu`` tests that
pin the travelling oracle (``oracle/resunet_oracle.py``) to the reference.  Those tests skip when
the directory is absent.

The reference's ``models/resunet.py:6`` imports ``torchlibrosa.stft``; the restatement under
``oracle/torchlibrosa`` is put first on ``sys.path``.
"""
import os
import sys

_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the byte-identical copy of the path's files that oracle/build_ref.py
# put under the git-ignored oracle/_ref (it travels with the snapshot)
_REF_COPY = os.path.join(_ORACLE_DIR, "_ref")
REFERENCE_ROOT = os.environ.get("LASS_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/models/resunet.py") else _REF_COPY)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "resunet.py"))


def import_reference_resunet():
    """Return the reference's ``models.resunet`` module (unmodified source)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if _ORACLE_DIR not in sys.path:
        sys.path.insert(0, _ORACLE_DIR)          # makes `import torchlibrosa` resolve to the restatement
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)       # `models` is a namespace package in the reference
    import importlib
    return importlib.import_module("models.resunet")


def mixers_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "data", "waveform_mixers.py"))


def import_reference_mixers():
    """Return the reference's ``data.waveform_mixers`` module (unmodified source).  Its module-level ``import pyloudnorm``
    (used only by the retired ``random_loudness_norm``) resolves to the stub under ``oracle/pyloudnorm``; the file is loaded by
    path so that the rest of the reference's ``data`` package (lightning, pyloudnorm-using datasets) is never imported."""
    if not mixers_available():
        raise RuntimeError("reference data/waveform_mixers.py not present under %s" % REFERENCE_ROOT)
    if _ORACLE_DIR not in sys.path:
        sys.path.insert(0, _ORACLE_DIR)
    import importlib.util
    import warnings
    spec = importlib.util.spec_from_file_location("_lass_reference_waveform_mixers",
                                                  os.path.join(REFERENCE_ROOT, "data", "waveform_mixers.py"))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)      # `import sre_compile` at data/waveform_mixers.py:2
        spec.loader.exec_module(mod)
    return mod
