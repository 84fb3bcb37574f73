"""The C-ABI library loads on a machine without a GPU and exports exactly the symbols that
include/lass_b200.h declares (no compute calls here)."""
import os
import re
import subprocess

import pytest

from lass_b200 import _cabi


def _declared(repo_root, header="lass_b200.h"):
    text = open(os.path.join(repo_root, "include", header)).read()
    return sorted(set(re.findall(r"LASS_API\s+[\w\s\*]+?\b(lass_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(repo_root):
    if not os.path.isfile(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    declared = _declared(repo_root)
    assert "lass_stft_fwd" in declared and "lass_mask_istft" in declared
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert sorted(_cabi.SIGNATURES) == declared, "ctypes SIGNATURES out of sync with include/lass_b200.h"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _cabi.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r"\bT (lass_\w+)", out)))
    assert exported == declared, "library exports %s, header declares %s" % (exported, declared)


def test_debug_library_is_separate(repo_root):
    """The probes / microbenchmarks are declared in their own header and exported by their own library only."""
    declared = _declared(repo_root, "lass_b200_debug.h")
    assert declared and all(n.startswith("lass_debug_umma_") for n in declared)
    assert sorted(_cabi.DEBUG_SIGNATURES) == declared
    lib = _cabi.load_debug()
    for name in declared:
        assert hasattr(lib, name)
        assert name not in _cabi.SIGNATURES
    out = subprocess.check_output(["nm", "-D", "--defined-only", _cabi.LIB_PATH]).decode()
    assert "lass_debug_umma_" not in out, "probe code leaked into the product library"


def test_version_and_error_calls_work_without_gpu():
    lib = _cabi.load()
    assert lib.lass_version() == 100
    assert lib.lass_stft_basis_rows(1024) == 4 * 256
    assert lib.lass_stft_workspace_bytes(2, 160000, 1024, 160) >= 2 * 2 * 161024 * 2
    # argument validation happens before any CUDA call
    rc = lib.lass_stft_fwd(None, 1, 100, 1024, 160, None, None, None, None, None, 0, 0, None, 0, None)
    assert rc == -1 and b"null" in lib.lass_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.LassLibraryError):
        _cabi.load()
