"""ctypes binding of the C ABI declared in ``include/lass_b200.h``.

The library is built in-tree by ``make`` / ``__graft_entry__.build()``.  Loading never falls back to anything
else: a missing library raises ``LassLibraryError`` with the build command.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "liblass_b200.so")

c_void_p, c_int, c_size_t, c_longlong = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_longlong


class LassLibraryError(RuntimeError):
    pass


class LassError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("lass_b200 error %d: %s" % (code, message))
        self.code = code


# name -> (restype, argtypes); must list EVERY symbol declared in include/lass_b200.h
# (tests/test_cabi_exports.py parses the header and checks both directions).
SIGNATURES = {
    "lass_version": (c_int, []),
    "lass_last_error": (ctypes.c_char_p, []),
    "lass_stft_basis_rows": (c_int, [c_int]),
    "lass_stft_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "lass_stft_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "lass_mask_istft": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lass_debug_umma_probe": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load (once) and return the ctypes handle of liblass_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise LassLibraryError(
                "%s not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`) "
                "at the repository root. lass_b200 has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise LassLibraryError("symbol %s missing from %s (stale build?)" % (name, LIB_PATH)) from exc
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load().lass_last_error()
        raise LassError(code, msg.decode("utf-8", "replace") if msg else "")
