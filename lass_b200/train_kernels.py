"""Tensor-level wrappers over the training-step entry points of the C ABI (``include/lass_b200.h``, section
"Training step").  CUDA tensors only, current stream, no CPU path: this module is the only kernel provider the
product ``lass_b200.training.TrainEngine`` uses (the parity tests may hand the engine a pure-torch emulation of
the SAME interface, ``tests/train_emul.py``, to check the step's algebra on a CPU without a GPU).

Layouts: activations / gradients NHWC 16-bit (raw conv outputs fp16, activated tensors and gradients bf16);
per-BatchNorm parameter block ``bnp`` = 6 x C fp32: [scale | shift | mean | rstd | coefA | coefB].
"""
import ctypes

import torch

from . import _cabi

# Forward tensors are fp16 (raw conv outputs saturating; activated tensors are post-BatchNorm O(1) values) with fp16 weights:
# with bf16 activations the TRAIN-mode forward reaches only 33-34 dB against the fp32 reference (batch-statistics BatchNorm
# renormalises every conv branch to unit variance, so each rounding counts fully), with fp16 47-49 dB — same bytes, same
# tensor-core rate (tcgen05 kind::f16).  Gradients span many decades (1e-9 .. 1e-3) and stay bf16.
RAW_DTYPE = torch.float16
ACT_DTYPE = torch.float16
GRAD_DTYPE = torch.bfloat16
# weight gradients: tcgen05 kernel (csrc/wgrad_tc.cu); False selects the mma.sync kernel it replaced (csrc/wgrad.cu, kept for A/B)
WGRAD_TENSOR_CORE = True


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


CALLS = 0          # kernel-launching C-ABI calls issued through this module (bench.py reports calls per step)


def _chk(code):
    global CALLS
    CALLS += 1
    _cabi.check(code)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lass_b200 training kernels need CUDA tensors (no CPU fallback); got %s" % t.device)
        if t is not None and t.device.index != torch.cuda.current_device():
            raise RuntimeError("tensor on %s but the current device is cuda:%d (TrainEngine sets it; direct callers wrap the "
                               "call in torch.cuda.device)" % (t.device, torch.cuda.current_device()))


def empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


class ConvSpec:
    """One implicit-GEMM convolution launch (lass_conv_desc): forward convs, transposed convs and dgrad convs."""

    def __init__(self, B, H, W, ncols, segs, bias=None, up=(1, 1), full_raw=None, full_raw_coff=0, pool=(1, 1),
                 pool_raw=None, after=None):
        self.B, self.H, self.W, self.ncols = B, H, W, ncols
        self.segs = segs                    # [(src NHWC 16-bit, coff, cin, weights (taps, ncols, cin) same dtype, taps)]
        self.bias = bias                    # (ncols) fp32 or None
        self.up = up
        self.full_raw, self.full_raw_coff = full_raw, full_raw_coff    # (B, H*uh, W*uw, C) 16-bit, first channel written
        self.pool, self.pool_raw = pool, pool_raw
        self.after = after                  # (after_w (3, ncols) fp32, after_b (3) fp32, feat (B, 3, H, W) fp32) or None
        self._handle = None
        self._keep = None


def conv(spec: ConvSpec):
    """Run (and on first use prepare: tensor maps, tile configuration) one conv launch."""
    lib = _cabi.load()
    if spec._handle is None:
        d = _cabi.ConvDesc()
        d.B, d.H, d.W, d.ncols, d.nseg = spec.B, spec.H, spec.W, spec.ncols, len(spec.segs)
        for i, (src, coff, cin, w, taps) in enumerate(spec.segs):
            _need_cuda(src, w)
            assert src.dtype == w.dtype and w.shape == (taps, spec.ncols, cin), (w.shape, taps, spec.ncols, cin)
            s = d.seg[i]
            s.src, s.src_cstride, s.src_coff, s.cin = _p(src), src.shape[3], coff, cin
            s.kc = 64 if cin % 64 == 0 else 32
            s.taps, s.fp16, s.weights = taps, 1 if src.dtype == torch.float16 else 0, _p(w)
        d.bias = _p(spec.bias)
        d.up_h, d.up_w = spec.up
        d.group_c = spec.ncols // (spec.up[0] * spec.up[1])
        d.pool_h, d.pool_w = spec.pool
        if spec.full_raw is not None:
            o = d.full_raw
            o.ptr, o.cstride, o.coff = _p(spec.full_raw), spec.full_raw.shape[3], spec.full_raw_coff
            o.fp16 = 1 if spec.full_raw.dtype == torch.float16 else 0
        if spec.pool_raw is not None:
            o = d.pool_raw
            o.ptr, o.cstride, o.coff = _p(spec.pool_raw), spec.pool_raw.shape[3], 0
            o.fp16 = 1 if spec.pool_raw.dtype == torch.float16 else 0
        if spec.after is not None:
            aw, ab, feat = spec.after
            d.after_w, d.after_b, d.feat = _p(aw), _p(ab), _p(feat)
        h = ctypes.c_void_p()
        _chk(lib.lass_conv_prepare(ctypes.byref(d), ctypes.byref(h)))
        spec._handle = _ConvHandle(h)
    _chk(lib.lass_conv_run(spec._handle.h, _stream()))


class _ConvHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            _cabi.load().lass_conv_destroy(self.h)
        except Exception:
            pass


def stft(wave, basis_hi, basis_lo, n_fft, hop, workspace):
    """(B, L) fp32 -> mag, cos, sin (B, T, F) fp32 (Base.spectrogram_phase semantics)."""
    from . import ops
    mag, cos, sin = ops.stft_fwd(wave, basis_hi, basis_lo, n_fft, hop, 0, workspace, 0)
    return mag[:, 0], cos[:, 0], sin[:, 0]


def film(cond, film_w, film_b, out):
    """out (B, J) = film_b + cond @ film_w.T (all FiLM linears as one GEMM)."""
    _need_cuda(cond, film_w, film_b, out)
    B, K = cond.shape
    _chk(_cabi.load().lass_film(_p(cond), _p(film_w), _p(film_b), B, K, film_w.shape[0], _p(out), _stream()))


def bn0_stats(mag, sums):
    """mag (B, T, F) fp32; sums (2, F) float64 zeroed by the callee: per-bin sum and sum of squares over (B, T)."""
    _need_cuda(mag, sums)
    B, T, F = mag.shape
    _chk(_cabi.load().lass_bn0_stats(_p(mag), B, T, F, _p(sums), _stream()))


def bn_stats(x, coff, C, sums):
    """x (B, H, W, Cbuf) 16-bit; sums (2, C) float64 (zeroed by the callee): per-channel sum / sum of squares."""
    _need_cuda(x, sums)
    npix = x.shape[0] * x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_stats(_p(x), 1 if x.dtype == torch.float16 else 0, npix, C, x.shape[3], coff, _p(sums),
                                    _stream()))


def bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps, bnp):
    """Batch statistics -> bnp[0:4C] = scale, shift, mean, rstd; running statistics updated in place (unbiased variance)."""
    _need_cuda(sums, gamma, beta, running_mean, running_var, bnp)
    C = gamma.numel()
    _chk(_cabi.load().lass_bn_finalize(_p(sums), float(count), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                       float(momentum), float(eps), C, _p(bnp), _stream()))


def bn_stats_acc(x, coff, C, sums):
    """bn_stats ADDED into ``sums`` (2, C) float64, which the caller zeroed (one memset for all sites of a step)."""
    _need_cuda(x, sums)
    npix = x.shape[0] * x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_stats_acc(_p(x), 1 if x.dtype == torch.float16 else 0, npix, C, x.shape[3], coff, _p(sums),
                                        _stream()))


def bn_act(x, x_coff, out, out_coff, C, bnp, beta):
    """out[..., out_coff:out_coff+C] = leaky_relu(scale*x + shift + beta[b]) in out's 16-bit type;  beta (B, C) row-strided view."""
    _need_cuda(x, out, bnp, beta)
    B, pix = x.shape[0], x.shape[1] * x.shape[2]
    assert beta.stride(1) == 1
    _chk(_cabi.load().lass_bn_act(_p(x), 1 if x.dtype == torch.float16 else 0, x.shape[3], x_coff, _p(out),
                                  1 if out.dtype == torch.float16 else 0, out.shape[3], out_coff, B, pix, C, _p(bnp), _p(beta),
                                  beta.stride(0), _stream()))


def bn_bwd_reduce(dact, x, x_coff, C, bnp, beta, sums):
    """sums (B, C, 2) fp32 (zeroed by the callee): [sum g', sum g' (x - mean)], g' = dact * lrelu'(scale*x + shift + beta)."""
    _need_cuda(dact, x, bnp, beta, sums)
    B, pix = x.shape[0], x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_bwd_reduce(_p(dact), dact.shape[3], 0, _p(x), 1 if x.dtype == torch.float16 else 0,
                                         x.shape[3], x_coff, B, pix, C, _p(bnp), _p(beta), beta.stride(0), _p(sums),
                                         _stream()))


def bn_bwd_finalize(sums, count, gamma, bnp, dgamma, dbeta, dfilm):
    """-> dgamma, dbeta (C), dfilm (B, C) row-strided view or None, bnp[4C:6C] = coefA, coefB."""
    _need_cuda(sums, gamma, bnp, dgamma, dbeta)
    B, C = sums.shape[0], sums.shape[1]
    _chk(_cabi.load().lass_bn_bwd_finalize(_p(sums), B, C, float(count), _p(gamma), _p(bnp), _p(dgamma), _p(dbeta),
                                           _p(dfilm), dfilm.stride(0) if dfilm is not None else 0, _stream()))


class PeerTable:
    """Every rank's flat sums buffer and flag table as the HOST pointer arrays the ``*_p2p`` entry points take."""

    def __init__(self, sums_ptrs, flag_ptrs, rank, status):
        import ctypes
        self.world, self.rank = len(sums_ptrs), int(rank)
        if self.world > _cabi.load().lass_syncbn_max_peers():
            raise RuntimeError("at most %d ranks share BatchNorm statistics over peer memory" % _cabi.load().lass_syncbn_max_peers())
        self.sums = (ctypes.c_void_p * self.world)(*[int(p) for p in sums_ptrs])
        self.flags = (ctypes.c_void_p * self.world)(*[int(p) for p in flag_ptrs])
        self.status = status


def bn_finalize_p2p(table, sums_offset, flag_index, epoch, count_total, gamma, beta, running_mean, running_var, momentum, eps, bnp):
    """bn_finalize over the sums of ALL ranks, the exchange fused into the kernel (NVLink peer memory, no collective call)."""
    _need_cuda(gamma, beta, running_mean, running_var, bnp, table.status)
    _chk(_cabi.load().lass_bn_finalize_p2p(table.sums, table.flags, table.world, table.rank, int(sums_offset), int(flag_index),
                                           int(epoch), float(count_total), _p(gamma), _p(beta), _p(running_mean),
                                           _p(running_var), float(momentum), float(eps), gamma.numel(), _p(bnp),
                                           _p(table.status), _stream()))


def bn_bwd_finalize_p2p(table, sums_offset, totals_offset, flag_index, epoch, B, count_total, gamma, bnp, dgamma, dbeta, dfilm):
    """bn_bwd_finalize_sync with the exchange inside the kernel: this rank's per-clip sums (float element ``sums_offset``) ->
    its per-channel totals (double element ``totals_offset``) -> every rank's totals through peer memory."""
    _need_cuda(gamma, bnp, dgamma, dbeta, table.status)
    _chk(_cabi.load().lass_bn_bwd_finalize_p2p(table.sums, table.flags, table.world, table.rank, int(sums_offset),
                                               int(totals_offset), int(flag_index), int(epoch), int(B), gamma.numel(),
                                               float(count_total),
                                               _p(gamma), _p(bnp), _p(dgamma), _p(dbeta), _p(dfilm),
                                               dfilm.stride(0) if dfilm is not None else 0, _p(table.status), _stream()))


def bn_bwd_totals(sums, totals):
    """This rank's per-channel totals (C, 2) float64 of the per-clip sums (B, C, 2): the SyncBatchNorm all-reduce payload."""
    _need_cuda(sums, totals)
    _chk(_cabi.load().lass_bn_bwd_totals(_p(sums), sums.shape[0], sums.shape[1], _p(totals), _stream()))


def bn_bwd_finalize_sync(sums, count_total, totals, gamma, bnp, dgamma, dbeta, dfilm):
    """bn_bwd_finalize for SyncBatchNorm: coefA / coefB from the all-reduced ``totals`` and the global count; dgamma, dbeta,
    dfilm from this rank's ``sums``."""
    _need_cuda(sums, totals, gamma, bnp, dgamma, dbeta)
    B, C = sums.shape[0], sums.shape[1]
    _chk(_cabi.load().lass_bn_bwd_finalize_sync(_p(sums), B, C, float(count_total), _p(totals), _p(gamma), _p(bnp), _p(dgamma),
                                                _p(dbeta), _p(dfilm), dfilm.stride(0) if dfilm is not None else 0, _stream()))


def bn_bwd_reduce_acc(dact, x, x_coff, C, bnp, beta, sums):
    """bn_bwd_reduce ADDED into ``sums`` (B, C, 2) fp32, which the caller zeroed."""
    _need_cuda(dact, x, bnp, beta, sums)
    B, pix = x.shape[0], x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_bwd_reduce_acc(_p(dact), dact.shape[3], 0, _p(x), 1 if x.dtype == torch.float16 else 0,
                                             x.shape[3], x_coff, B, pix, C, _p(bnp), _p(beta), beta.stride(0), _p(sums),
                                             _stream()))


def bn_bwd_reduce_finalize(dact, x, x_coff, C, bnp, beta, sums, counter, gamma, dgamma, dbeta, dfilm):
    """A/B variant (not used by the engine: slower than reduce + finalize as two launches).  bn_bwd_reduce + bn_bwd_finalize in
    one launch; ``sums`` (B, C, 2) fp32 and ``counter`` zero on entry, left dirty."""
    _need_cuda(dact, x, bnp, beta, sums, counter, gamma, dgamma, dbeta)
    B, pix = x.shape[0], x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_bwd_reduce_finalize(_p(dact), dact.shape[3], 0, _p(x), 1 if x.dtype == torch.float16 else 0,
                                                  x.shape[3], x_coff, B, pix, C, _p(bnp), _p(beta), beta.stride(0), _p(sums),
                                                  _p(counter), _p(gamma), _p(dgamma), _p(dbeta), _p(dfilm),
                                                  dfilm.stride(0) if dfilm is not None else 0, _stream()))


def bn_bwd_apply(dact, x, x_coff, C, bnp, beta, add, add_coff, dx, dx_coff):
    """dx[..., dx_coff:+C] = bf16(scale*g' + coefA*(x - mean) + coefB (+ add[..., add_coff:+C]))."""
    _need_cuda(dact, x, bnp, beta, dx)
    B, pix = x.shape[0], x.shape[1] * x.shape[2]
    _chk(_cabi.load().lass_bn_bwd_apply(_p(dact), dact.shape[3], 0, _p(x), 1 if x.dtype == torch.float16 else 0, x.shape[3],
                                        x_coff, _p(add), add.shape[3] if add is not None else 0, add_coff, _p(dx),
                                        dx.shape[3], dx_coff, B, pix, C, _p(bnp), _p(beta), beta.stride(0), _stream()))


def pool_bwd(dpool, dskip, dskip_coff, dy, ph, pw):
    """dy (B, H, W, C) = dskip[..., coff:coff+C] (or 0) + upsample(dpool) / (ph*pw)."""
    _need_cuda(dpool, dy)
    B, H, W, C = dy.shape
    _chk(_cabi.load().lass_pool_bwd(_p(dpool), _p(dskip), dskip.shape[3] if dskip is not None else 0, dskip_coff, _p(dy),
                                    B, H, W, C, ph, pw, _stream()))


def unshuffle(src, src_coff, C, dst, uh, uw):
    """dst (B, H, W, uh*uw*C)[(dy*uw+dx)*C + c] = src (B, H*uh, W*uw, Cbuf)[h*uh+dy, w*uw+dx, src_coff + c]."""
    _need_cuda(src, dst)
    B, H, W, _ = dst.shape
    _chk(_cabi.load().lass_unshuffle(_p(src), src.shape[3], src_coff, _p(dst), B, H, W, C, uh, uw, _stream()))


def channel_sum(x, coff, C, out, acc=False):
    """out (C) fp32 = sum over pixels of x[..., coff:coff+C] (overwrites; acc: added to a pre-zeroed out, no memset inside)."""
    _need_cuda(x, out)
    npix = x.shape[0] * x.shape[1] * x.shape[2]
    lib = _cabi.load()
    _chk((lib.lass_channel_sum_acc if acc else lib.lass_channel_sum)(_p(x), npix, C, x.shape[3], coff, _p(out), _stream()))


def wgrad(dy, dy_coff, co, x, x_coff, ci, taps, dw, acc=False):
    """dw (taps, co, ci) fp32 (overwritten) = sum_p dy[p, co] * x[p + tap, ci] (zero padding; tap = ky*3+kx, centre 4).
    acc: ADDED to a pre-zeroed dw (tcgen05 kernel only; no memset inside)."""
    _need_cuda(dy, x, dw)
    B, H, W, _ = dy.shape
    assert x.shape[:3] == dy.shape[:3] and dw.numel() == taps * co * ci and dw.dtype == torch.float32
    lib = _cabi.load()
    if acc and not WGRAD_TENSOR_CORE:
        dw.zero_()
    fn = (lib.lass_wgrad_tc_acc if acc else lib.lass_wgrad_tc) if WGRAD_TENSOR_CORE else lib.lass_wgrad
    _chk(fn(_p(dy), dy.shape[3], dy_coff, co, _p(x), 1 if x.dtype == torch.float16 else 0, x.shape[3], x_coff, ci, B, H, W,
            taps, _p(dw), _stream()))


def pre_fwd(mag, bnp0, pre_w, pre_b, x0):
    """x0 (B, Tp, Fp, 32) fp16 = pre_conv(pad(bn0(mag))[..., :Fp]) (reference models/resunet.py:537-555)."""
    _need_cuda(mag, bnp0, pre_w, pre_b, x0)
    B, T, F = mag.shape
    _chk(_cabi.load().lass_pre_fwd(_p(mag), B, T, F, x0.shape[1], x0.shape[2], _p(bnp0), _p(pre_w), _p(pre_b), _p(x0),
                                   _stream()))


def pre_bwd(dx0, mag, bnp0, pre_w, dpre_w, dpre_b, dgamma0, dbeta0):
    """Gradients of pre_conv weight / bias (32 each) and bn0 weight / bias (F each; the dropped Nyquist bin gets 0)."""
    _need_cuda(dx0, mag, bnp0, pre_w, dpre_w, dpre_b, dgamma0, dbeta0)
    B, T, F = mag.shape
    _chk(_cabi.load().lass_pre_bwd(_p(dx0), _p(mag), B, T, F, dx0.shape[1], dx0.shape[2], _p(bnp0), _p(pre_w), _p(dpre_w),
                                   _p(dpre_b), _p(dgamma0), _p(dbeta0), _stream()))


def after_bwd(dfeat, y, after_w, dy, dw, db):
    """after_conv backward: dy (B, H, W, 32) bf16, dw (3, 32), db (3) from dfeat (B, 3, H, W) fp32 and y (B, H, W, 32) fp16."""
    _need_cuda(dfeat, y, after_w, dy, dw, db)
    B, H, W, _ = y.shape
    _chk(_cabi.load().lass_after_bwd(_p(dfeat), _p(y), _p(after_w), _p(dy), _p(dw), _p(db), B, H * W, _stream()))


def mask_istft(feat, mag, cos, sin, window, twiddle, n_fft, hop, length):
    from . import ops
    return ops.mask_istft(feat, mag[:, None], cos[:, None], sin[:, None], window, twiddle, n_fft, hop, length)


def istft_bwd(dwave, window, basis_hi, basis_lo, n_fft, hop, T, workspace, dre, dim):
    """Adjoint of torchlibrosa ISTFT.forward (without the c_f / n_fft factor): dre, dim (B, T, F) fp32 =
    STFT of the zero-extended dwave / window-sum."""
    _need_cuda(dwave, window, basis_hi, basis_lo, workspace, dre, dim)
    B, L = dwave.shape
    _chk(_cabi.load().lass_istft_bwd(_p(dwave), B, L, n_fft, hop, T, _p(window), _p(basis_hi), _p(basis_lo), _p(dre),
                                     _p(dim), _p(workspace), workspace.numel(), _stream()))


def mask_bwd(feat, mag, cos, sin, dre, dim, dfeat, n_fft):
    """dfeat (B, 3, Tp, Fp) fp32 from dre/dim (B, T, F): backward of feature_maps_to_wav's mask (models/resunet.py:457-505)."""
    _need_cuda(feat, mag, cos, sin, dre, dim, dfeat)
    B, T, F = mag.shape
    _chk(_cabi.load().lass_mask_bwd(_p(feat), _p(mag), _p(cos), _p(sin), _p(dre), _p(dim), _p(dfeat), B, T, F,
                                    feat.shape[2], feat.shape[3], n_fft, _stream()))


def l1_loss(wave, target, dwave, loss_sum):
    """loss_sum (1) fp32 += sum |wave - target|; dwave = sign(wave - target) / numel (reference losses.py:4-9)."""
    _need_cuda(wave, target, dwave, loss_sum)
    _chk(_cabi.load().lass_l1_loss(_p(wave), _p(target), wave.numel(), _p(loss_sum), _p(dwave), 1.0 / wave.numel(),
                                   _stream()))


def film_bwd(dbeta, cond, dw, db):
    """dw (J, K) = dbeta.T @ cond, db (J) = dbeta.sum(0)."""
    _need_cuda(dbeta, cond, dw, db)
    B, J = dbeta.shape
    _chk(_cabi.load().lass_film_bwd(_p(dbeta), _p(cond), _p(dw), _p(db), B, J, cond.shape[1], _stream()))


def adamw_amsgrad(p, g, m, v, vmax, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """Fused AdamW(amsgrad=True) over flat fp32 buffers (reference models/audiosep.py:122-130); step counts from 1."""
    _need_cuda(p, g, m, v, vmax)
    _chk(_cabi.load().lass_adamw_amsgrad(_p(p), _p(g), _p(m), _p(v), _p(vmax), p.numel(), float(lr), float(beta1),
                                         float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
                                         _stream()))


# weight kinds of pack_weight / unpack_grad
KIND_CONV, KIND_CONVT = 0, 1


def pack_weight(w, kind, fwd, dgrad):
    """fp32 parameter (torch layout) -> the 16-bit kernel layouts.
    KIND_CONV  w (co, ci, kh, kw): fwd (taps, co, ci) [dtype of `fwd`], dgrad (taps, ci, co) bf16 with flipped taps.
    KIND_CONVT w (ci, co, kh, kw): fwd (1, kh*kw*co, ci), dgrad (1, ci, kh*kw*co).  Either output may be None."""
    _need_cuda(w, fwd, dgrad)
    if kind == KIND_CONV:
        co, ci, taps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
    else:
        ci, co, taps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
    _chk(_cabi.load().lass_pack_weight(_p(w), kind, co, ci, taps, _p(fwd),
                                       1 if (fwd is not None and fwd.dtype == torch.float16) else 0, _p(dgrad), _stream()))


def unpack_grad(dw, kind, grad):
    """Packed fp32 weight gradient (taps, co, ci) [KIND_CONVT: (1, kh*kw*co, ci)] -> `grad` in the parameter's torch layout."""
    _need_cuda(dw, grad)
    if kind == KIND_CONV:
        co, ci, taps = grad.shape[0], grad.shape[1], grad.shape[2] * grad.shape[3]
    else:
        ci, co, taps = grad.shape[0], grad.shape[1], grad.shape[2] * grad.shape[3]
    _chk(_cabi.load().lass_unpack_grad(_p(dw), kind, co, ci, taps, _p(grad), _stream()))


class MultiTable:
    """Device table of a multi-tensor pack / unpack launch (8 x int64 per tensor, see include/lass_b200.h); keeps the tensors
    it points to alive."""

    def __init__(self, rows, keep, device, pack):
        lib = _cabi.load()
        chunk = lib.lass_multi_chunk()
        flat, block = [], 0
        for a, b, c, kind, co, ci, taps_word in rows:
            flat += [a, b, c, kind, co, ci, taps_word, block]
            block += lib.lass_pack_blocks(kind, co, ci) if pack else (co * ci * (taps_word & 0xffff) + chunk - 1) // chunk
        self.table = torch.tensor(flat, dtype=torch.int64).to(device)
        self.nitems, self.nblocks, self.keep = len(rows), block, keep


def _dims(w, kind):
    taps = w.shape[2] * w.shape[3]
    return (w.shape[0], w.shape[1], taps) if kind == KIND_CONV else (w.shape[1], w.shape[0], taps)


def pack_weights_table(items, device):
    """items: [(w fp32 torch layout, kind, fwd 16-bit or None, dgrad bf16 or None)] -> table for pack_weights."""
    rows = []
    for w, kind, fwd, dgrad in items:
        _need_cuda(w, fwd, dgrad)
        co, ci, taps = _dims(w, kind)
        rows.append((_p(w), _p(fwd) or 0, _p(dgrad) or 0, kind, co, ci,
                     taps | ((1 if (fwd is not None and fwd.dtype == torch.float16) else 0) << 16)))
    return MultiTable(rows, items, device, True)


def pack_weights(table):
    """pack_weight for every tensor of the table in one launch."""
    _chk(_cabi.load().lass_pack_weights_multi(_p(table.table), table.nitems, table.nblocks, _stream()))


def unpack_grads_table(items, device):
    """items: [(packed dw fp32 flat, kind, grad fp32 in the parameter's torch layout)] -> table for unpack_grads."""
    rows = []
    for dw, kind, grad in items:
        _need_cuda(dw, grad)
        co, ci, taps = _dims(grad, kind)
        rows.append((_p(dw), _p(grad), 0, kind, co, ci, taps))
    return MultiTable(rows, items, device, False)


def unpack_grads(table):
    """unpack_grad for every tensor of the table in one launch."""
    _chk(_cabi.load().lass_unpack_grads_multi(_p(table.table), table.nitems, table.nblocks, _stream()))
