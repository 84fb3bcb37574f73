"""Tensor-level wrappers over the C ABI (device pointers of torch CUDA tensors, current stream).

These are the per-kernel entry points used by the module in ``lass_b200.models.resunet`` and by the parity
tests.  Every function requires CUDA tensors and raises otherwise — there is no CPU path.
"""
import ctypes

import torch

from . import _cabi


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("lass_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if not t.is_contiguous():
            raise RuntimeError("lass_b200 ops need contiguous tensors")
        if t.device.index != torch.cuda.current_device():
            raise RuntimeError("tensor on %s but the current device is cuda:%d: the kernels launch on the current device's "
                               "stream -- wrap the call in torch.cuda.device(tensor.device)" % (t.device, torch.cuda.current_device()))


def stft_fwd(wave: torch.Tensor, basis_hi: torch.Tensor, basis_lo: torch.Tensor, n_fft: int, hop: int,
             precision_mode: int = 0, workspace: torch.Tensor = None, magphase_mode: int = 0):
    """wave (B, L) fp32 -> mag, cos, sin (B, 1, T, F) fp32 (reference layout, models/base.py:83-88;
    magphase_mode = 1 gives torchlibrosa.stft.magphase semantics instead)."""
    lib = _cabi.load()
    _require_cuda(wave, basis_hi, basis_lo)
    assert wave.dtype == torch.float32 and wave.dim() == 2
    B, L = wave.shape
    T, F = L // hop + 1, n_fft // 2 + 1
    need = lib.lass_stft_workspace_bytes(B, L, n_fft, hop)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=wave.device)
    out = torch.empty(3, B, 1, T, F, dtype=torch.float32, device=wave.device)
    _cabi.check(lib.lass_stft_fwd(_ptr(wave), B, L, n_fft, hop, _ptr(basis_hi), _ptr(basis_lo), _ptr(out[0]),
                                  _ptr(out[1]), _ptr(out[2]), precision_mode, magphase_mode, _ptr(workspace),
                                  workspace.numel() * workspace.element_size(), _stream()))
    return out[0], out[1], out[2]


def stft_multi_fwd(wave: torch.Tensor, bases, n_ffts, hop: int, precision_mode: int = 0, workspace: torch.Tensor = None,
                   magphase_mode: int = 0):
    """Several STFT resolutions of the same waveforms in ONE kernel launch (lass_stft_multi_fwd): wave (B, L) fp32,
    bases = [(hi, lo)] per resolution, n_ffts = [n_fft] -> [(mag, cos, sin)] each (B, 1, T, n_fft/2 + 1) fp32."""
    lib = _cabi.load()
    n = len(n_ffts)
    assert wave.dtype == torch.float32 and wave.dim() == 2 and len(bases) == n and 1 <= n <= 3
    _require_cuda(wave, *[t for pair in bases for t in pair])
    B, L = wave.shape
    T = L // hop + 1
    need = sum(lib.lass_stft_workspace_bytes(B, L, int(f), hop) for f in n_ffts)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=wave.device)
    outs = [torch.empty(3, B, 1, T, int(f) // 2 + 1, dtype=torch.float32, device=wave.device) for f in n_ffts]
    arr_i = (ctypes.c_int * n)(*[int(f) for f in n_ffts])
    ptrs = lambda vals: (ctypes.c_void_p * n)(*vals)
    _cabi.check(lib.lass_stft_multi_fwd(_ptr(wave), B, L, hop, n, arr_i, ptrs([_ptr(b[0]) for b in bases]),
                                        ptrs([_ptr(b[1]) for b in bases]), ptrs([_ptr(o[0]) for o in outs]),
                                        ptrs([_ptr(o[1]) for o in outs]), ptrs([_ptr(o[2]) for o in outs]), precision_mode,
                                        magphase_mode, _ptr(workspace), workspace.numel() * workspace.element_size(), _stream()))
    return [(o[0], o[1], o[2]) for o in outs]


def mask_istft(feat: torch.Tensor, mag: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, window: torch.Tensor,
               twiddle: torch.Tensor, n_fft: int, hop: int, length: int, feat_F: int = None):
    """feat (B, 3, Tf, Ff) fp32 (Tf >= T rows, first ``feat_F`` bins valid) + mixture mag/cos/sin (B, 1, T, F)
    -> waveform (B, length).  Reference: models/resunet.py:436-519 + torchlibrosa ISTFT."""
    lib = _cabi.load()
    _require_cuda(feat, mag, cos, sin, window, twiddle)
    B, _, T, F = mag.shape
    assert feat.shape[0] == B and feat.shape[1] == 3 and feat.shape[2] >= T
    Ff = feat.shape[3]
    feat_F = min(Ff, F) if feat_F is None else feat_F
    out = torch.empty(B, length, dtype=torch.float32, device=mag.device)
    _cabi.check(lib.lass_mask_istft(_ptr(feat), feat.stride(0), feat.stride(1), feat.stride(2), feat_F, _ptr(mag),
                                    _ptr(cos), _ptr(sin), _ptr(window), _ptr(twiddle), B, T, F, n_fft, hop, length,
                                    _ptr(out), _stream()))
    return out


def umma_probe(A: torch.Tensor, Bm: torch.Tensor, swizzle_mode: int, a_start_bytes: int, a_sbo: int,
               a_base_offset: int, b_sbo: int):
    """One tcgen05.mma tile through caller-chosen descriptors (liblass_b200_debug.so; tests only)."""
    lib = _cabi.load_debug()
    _require_cuda(A, Bm)
    a_rows, kc = A.shape
    n = Bm.shape[0]
    out = torch.zeros(128, n, dtype=torch.float32, device=A.device)
    _cabi.check(lib.lass_debug_umma_probe(_ptr(A), a_rows, _ptr(Bm), n, kc, swizzle_mode, a_start_bytes, a_sbo,
                                          a_base_offset, b_sbo, 1 if A.dtype == torch.float16 else 0, _ptr(out),
                                          _stream()), lib)
    return out


def make_segment(src: torch.Tensor, coff: int, cin: int, weights: torch.Tensor, taps: int):
    """src (B, H, W, Cbuf) 16-bit NHWC; weights (taps, ncols, cin) of the same dtype."""
    _require_cuda(src, weights)
    assert src.dtype in (torch.bfloat16, torch.float16) and weights.dtype == src.dtype
    assert weights.shape[2] == cin and weights.shape[0] in (taps, 3)   # (3, 3*cout, cin) for the dx-in-N layout
    seg = _cabi.ConvSegment()
    seg.src = _ptr(src)
    seg.src_cstride = src.shape[3]
    seg.src_coff = coff
    seg.cin = cin
    seg.kc = 64 if cin % 64 == 0 else 32
    seg.taps = taps
    seg.fp16 = 1 if src.dtype == torch.float16 else 0
    seg.weights = _ptr(weights)
    return seg


def make_out(dst: torch.Tensor = None, coff: int = 0, scale: torch.Tensor = None, shift: torch.Tensor = None):
    """dst (B, Ho, Wo, Cbuf) 16-bit NHWC (fp16 -> saturating raw store, bf16 otherwise); optional activation
    lrelu(scale[c]*v + shift[b, c]) with shift a (B, J) table view starting at this layer's first column."""
    o = _cabi.ConvOut()
    if dst is None:
        return o
    _require_cuda(dst)
    o.ptr = _ptr(dst)
    o.cstride = dst.shape[3]
    o.coff = coff
    o.fp16 = 1 if dst.dtype == torch.float16 else 0
    if scale is not None:
        assert scale.is_cuda and shift.is_cuda and scale.dtype == torch.float32 and shift.dtype == torch.float32
        o.scale = _ptr(scale)
        o.shift = _ptr(shift)
        o.shift_bstride = shift.stride(0)
    return o


def conv_igemm(B, H, W, ncols, segments, bias=None, up=(1, 1), full_raw=None, full_act=None, pool=(1, 1),
               pool_raw=None, pool_act=None, after_w=None, after_b=None, feat=None, resid=None, algo=0, gen=None):
    """resid: optional (src (B, T, F) fp32, in_scale (F), in_shift (F), w (ncols), b (ncols)) rank-1 residual.
    gen: optional (src (B, T, F) fp32, in_scale (F), in_shift (F), w (32), b (32), scale (32), shift (B, >=32 strided))
    generated A operand of segment 0 (see lass_conv_desc.gen_src)."""
    lib = _cabi.load()
    d = _cabi.ConvDesc()
    d.B, d.H, d.W, d.ncols, d.nseg = B, H, W, ncols, len(segments)
    for i, s in enumerate(segments):
        d.seg[i] = s
    d.bias = _ptr(bias) if bias is not None else None
    d.up_h, d.up_w = up
    d.group_c = ncols // (up[0] * up[1])
    d.full_raw = full_raw if full_raw is not None else _cabi.ConvOut()
    d.full_act = full_act if full_act is not None else _cabi.ConvOut()
    d.pool_h, d.pool_w = pool
    d.pool_raw = pool_raw if pool_raw is not None else _cabi.ConvOut()
    d.pool_act = pool_act if pool_act is not None else _cabi.ConvOut()
    d.after_w = _ptr(after_w) if after_w is not None else None
    d.after_b = _ptr(after_b) if after_b is not None else None
    d.feat = _ptr(feat) if feat is not None else None
    d.algo = algo
    if resid is not None:
        src, isc, ish, rw, rb = resid
        _require_cuda(src, isc, ish, rw, rb)
        d.resid_src, d.resid_in_scale, d.resid_in_shift = _ptr(src), _ptr(isc), _ptr(ish)
        d.resid_w, d.resid_b = _ptr(rw), _ptr(rb)
        d.resid_T, d.resid_F = src.shape[1], src.shape[2]
    if gen is not None:
        src, isc, ish, gw, gb, gsc, gsh = gen
        _require_cuda(src, isc, ish, gw, gb, gsc)
        assert gsh.is_cuda and gsh.stride(1) == 1   # a row-strided view of the shift table is fine
        d.gen_src, d.gen_in_scale, d.gen_in_shift = _ptr(src), _ptr(isc), _ptr(ish)
        d.gen_w, d.gen_b, d.gen_scale, d.gen_shift = _ptr(gw), _ptr(gb), _ptr(gsc), _ptr(gsh)
        d.gen_shift_bstride = gsh.stride(0)
        d.gen_T, d.gen_F = src.shape[1], src.shape[2]
    _cabi.check(lib.lass_conv_igemm(ctypes.byref(d), _stream()))
