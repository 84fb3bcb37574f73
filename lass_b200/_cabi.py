"""ctypes binding of the C ABI declared in ``include/lass_b200.h``.

The library is built in-tree by ``make`` / ``__graft_entry__.build()``.  Loading never falls back to anything
else: a missing library raises ``LassLibraryError`` with the build command.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# LASS_B200_LIB selects another BUILD of the same library (e.g. `make prof`); there is still no fallback of any kind.
LIB_PATH = os.environ.get("LASS_B200_LIB") or os.path.join(_HERE, "_lib", "liblass_b200.so")

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
                              c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "lass_mask_istft": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}
# liblass_b200_debug.so (include/lass_b200_debug.h): descriptor probes / microbenchmarks, not in the product library
DEBUG_LIB_PATH = os.path.join(_HERE, "_lib", "liblass_b200_debug.so")
DEBUG_SIGNATURES = {
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


_dbg = None


def load_debug():
    """The probe / microbenchmark library (tests and tools only)."""
    global _dbg
    if _dbg is None:
        with _lock:
            if _dbg is None:
                if not os.path.isfile(DEBUG_LIB_PATH):
                    raise LassLibraryError("%s not found: build it with `make`" % DEBUG_LIB_PATH)
                lib = ctypes.CDLL(DEBUG_LIB_PATH)
                for name, (restype, argtypes) in list(DEBUG_SIGNATURES.items()) + [("lass_last_error", (ctypes.c_char_p, []))]:
                    fn = getattr(lib, name)
                    fn.restype, fn.argtypes = restype, argtypes
                _dbg = lib
    return _dbg


def check(code, lib=None):
    if code != 0:
        msg = (lib or load()).lass_last_error()
        raise LassError(code, msg.decode("utf-8", "replace") if msg else "")


# ---- C structs of include/lass_b200.h (host memory) ----
class ConvSegment(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("src_cstride", c_int), ("src_coff", c_int), ("cin", c_int), ("kc", c_int),
                ("taps", c_int), ("fp16", c_int), ("weights", c_void_p)]


class ConvOut(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("cstride", c_int), ("coff", c_int), ("fp16", c_int), ("scale", c_void_p),
                ("shift", c_void_p), ("shift_bstride", c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("W", c_int), ("ncols", c_int), ("nseg", c_int),
                ("seg", ConvSegment * 2), ("bias", c_void_p), ("up_h", c_int), ("up_w", c_int), ("group_c", c_int),
                ("full_raw", ConvOut), ("full_act", ConvOut), ("pool_h", c_int), ("pool_w", c_int),
                ("pool_raw", ConvOut), ("pool_act", ConvOut), ("after_w", c_void_p), ("after_b", c_void_p),
                ("feat", c_void_p), ("resid_src", c_void_p), ("resid_in_scale", c_void_p),
                ("resid_in_shift", c_void_p), ("resid_w", c_void_p), ("resid_b", c_void_p), ("resid_T", c_int),
                ("resid_F", c_int), ("algo", c_int), ("gen_src", c_void_p), ("gen_in_scale", c_void_p),
                ("gen_in_shift", c_void_p), ("gen_w", c_void_p), ("gen_b", c_void_p), ("gen_scale", c_void_p),
                ("gen_shift", c_void_p), ("gen_shift_bstride", c_int), ("gen_T", c_int), ("gen_F", c_int)]


SIGNATURES["lass_conv_igemm"] = (c_int, [ctypes.POINTER(ConvDesc), c_void_p])
SIGNATURES["lass_debug_set_conv_flags"] = (c_int, [c_int])
SIGNATURES["lass_debug_set_conv_profile"] = (c_int, [c_void_p])
DEBUG_SIGNATURES["lass_debug_umma_bench"] = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p])
DEBUG_SIGNATURES["lass_debug_umma_bench3"] = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p])
DEBUG_SIGNATURES["lass_debug_umma_bench2"] = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p])


# ---- whole-model entry (lass_resunet30_*) ----
class _EncW(ctypes.Structure):
    _fields_ = [("conv1_w", ctypes.c_void_p), ("conv2_w", ctypes.c_void_p), ("sc_w", ctypes.c_void_p),
                ("sc_b", ctypes.c_void_p)]


class _DecW(ctypes.Structure):
    _fields_ = [("up_w", ctypes.c_void_p), ("conv1_w", ctypes.c_void_p), ("conv2_w", ctypes.c_void_p),
                ("sc_w", ctypes.c_void_p), ("sc_b", ctypes.c_void_p)]


class ResUNet30Weights(ctypes.Structure):
    """Mirror of ``lass_resunet30_weights``."""
    _fields_ = [("n_fft", ctypes.c_int), ("hop", ctypes.c_int), ("condition_size", ctypes.c_int),
                ("film_rows", ctypes.c_int),
                ("stft_basis_hi", ctypes.c_void_p), ("stft_basis_lo", ctypes.c_void_p),
                ("istft_window", ctypes.c_void_p), ("istft_twiddle", ctypes.c_void_p),
                ("bn0_scale", ctypes.c_void_p), ("bn0_shift", ctypes.c_void_p),
                ("pre_w", ctypes.c_void_p), ("pre_b", ctypes.c_void_p),
                ("film_w", ctypes.c_void_p), ("film_b", ctypes.c_void_p), ("act_scale", ctypes.c_void_p),
                ("enc", _EncW * 7), ("dec", _DecW * 6),
                ("after_w", ctypes.c_void_p), ("after_b", ctypes.c_void_p), ("dxn_mask", ctypes.c_uint)]


SIGNATURES.update({
    "lass_film": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "lass_resunet30_film_rows": (ctypes.c_int, []),
    "lass_resunet30_film_offset": (ctypes.c_int, [ctypes.c_int]),
    "lass_resunet30_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lass_resunet30_plan_create": (ctypes.c_int, [ctypes.POINTER(ResUNet30Weights), ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(ctypes.c_void_p)]),
    "lass_resunet30_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "lass_resunet30_forward_stages": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "lass_resunet30_unet_flops": (ctypes.c_double, [ctypes.c_void_p]),
    "lass_debug_time_unet_launches": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                     ctypes.c_void_p]),
    "lass_resunet30_num_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "lass_resunet30_buffer": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int * 4),
                                                ctypes.POINTER(ctypes.c_int)]),
    "lass_resunet30_plan_destroy": (None, [ctypes.c_void_p]),
})

# ---- prepared convs + training step ----
_v, _i, _ll, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double
SIGNATURES.update({
    "lass_conv_prepare": (_i, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_void_p)]),
    "lass_conv_run": (_i, [_v, _v]),
    "lass_conv_destroy": (None, [_v]),
    "lass_bn_stats": (_i, [_v, _i, _ll, _i, _i, _i, _v, _v]),
    "lass_bn0_stats": (_i, [_v, _i, _i, _i, _v, _v]),
    "lass_bn_finalize": (_i, [_v, _d, _v, _v, _v, _v, _f, _f, _i, _v, _v]),
    "lass_bn_act": (_i, [_v, _i, _i, _i, _v, _i, _i, _i, _i, _ll, _i, _v, _v, _i, _v]),
    "lass_bn_bwd_reduce": (_i, [_v, _i, _i, _v, _i, _i, _i, _i, _ll, _i, _v, _v, _i, _v, _v]),
    "lass_bn_bwd_finalize": (_i, [_v, _i, _i, _d, _v, _v, _v, _v, _v, _i, _v]),
    "lass_bn_bwd_apply": (_i, [_v, _i, _i, _v, _i, _i, _i, _v, _i, _i, _v, _i, _i, _i, _ll, _i, _v, _v, _i, _v]),
    "lass_pool_bwd": (_i, [_v, _v, _i, _i, _v, _i, _i, _i, _i, _i, _i, _v]),
    "lass_unshuffle": (_i, [_v, _i, _i, _v, _i, _i, _i, _i, _i, _i, _v]),
    "lass_channel_sum": (_i, [_v, _ll, _i, _i, _i, _v, _v]),
    "lass_wgrad": (_i, [_v, _i, _i, _i, _v, _i, _i, _i, _i, _i, _i, _i, _i, _v, _v]),
    "lass_pre_fwd": (_i, [_v, _i, _i, _i, _i, _i, _v, _v, _v, _v, _v]),
    "lass_pre_bwd": (_i, [_v, _v, _i, _i, _i, _i, _i, _v, _v, _v, _v, _v, _v, _v]),
    "lass_after_bwd": (_i, [_v, _v, _v, _v, _v, _v, _i, _ll, _v]),
    "lass_istft_bwd": (_i, [_v, _i, _i, _i, _i, _i, _v, _v, _v, _v, _v, _v, ctypes.c_size_t, _v]),
    "lass_mask_bwd": (_i, [_v, _v, _v, _v, _v, _v, _v, _i, _i, _i, _i, _i, _i, _v]),
    "lass_l1_loss": (_i, [_v, _v, _ll, _v, _v, _f, _v]),
    "lass_film_bwd": (_i, [_v, _v, _v, _v, _i, _i, _i, _v]),
    "lass_adamw_amsgrad": (_i, [_v, _v, _v, _v, _v, _ll, _f, _f, _f, _f, _f, _i, _f, _v]),
    "lass_pack_weight": (_i, [_v, _i, _i, _i, _i, _v, _i, _v, _v]),
    "lass_unpack_grad": (_i, [_v, _i, _i, _i, _i, _v, _v]),
    "lass_debug_set_istft_v1": (_i, [_i]),
    "lass_bn_stats_acc": (_i, [_v, _i, _ll, _i, _i, _i, _v, _v]),
    "lass_bn_bwd_reduce_acc": (_i, [_v, _i, _i, _v, _i, _i, _i, _i, _ll, _i, _v, _v, _i, _v, _v]),
    "lass_bn_bwd_reduce_finalize": (_i, [_v, _i, _i, _v, _i, _i, _i, _i, _ll, _i, _v, _v, _i, _v, _v, _v, _v, _v, _v, _i, _v]),
    "lass_pack_weights_multi": (_i, [_v, _i, _i, _v]),
    "lass_unpack_grads_multi": (_i, [_v, _i, _i, _v]),
    "lass_multi_chunk": (_i, []),
    "lass_pack_blocks": (_i, [_i, _i, _i]),
})
SIGNATURES["lass_syncbn_max_peers"] = (_i, [])
SIGNATURES["lass_bn_finalize_p2p"] = (_i, [_v, _v, _i, _i, _ll, _i, ctypes.c_ulonglong, _d, _v, _v, _v, _v, _f, _f, _i, _v, _v, _v])
SIGNATURES["lass_bn_bwd_finalize_p2p"] = (_i, [_v, _v, _i, _i, _ll, _ll, _i, ctypes.c_ulonglong, _i, _i, _d, _v, _v, _v, _v, _v, _i, _v, _v])
SIGNATURES["lass_bn_bwd_totals"] = (_i, [_v, _i, _i, _v, _v])
SIGNATURES["lass_bn_bwd_finalize_sync"] = (_i, [_v, _i, _i, _d, _v, _v, _v, _v, _v, _v, _i, _v])
SIGNATURES["lass_segment_mix_scratch_bytes"] = (ctypes.c_size_t, [_i])
SIGNATURES["lass_segment_mix"] = (_i, [_v, _i, _i, _i, _v, _v, _v, _v, ctypes.c_size_t, _v])
SIGNATURES["lass_wgrad_tc"] = SIGNATURES["lass_wgrad"]
SIGNATURES["lass_wgrad_tc_acc"] = SIGNATURES["lass_wgrad"]
SIGNATURES["lass_channel_sum_acc"] = SIGNATURES["lass_channel_sum"]
SIGNATURES["lass_stft_multi_fwd"] = (_i, [_v, _i, _i, _i, _i, _v, _v, _v, _v, _v, _v, _i, _i, _v, ctypes.c_size_t, _v])
DEBUG_SIGNATURES["lass_debug_umma_probe_mn"] = (_i, [_v, _i, _i, _v, _i, _i, _i, _i] + [_i] * 10 + [_v, _v])
