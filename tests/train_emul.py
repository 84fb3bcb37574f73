"""TEST INFRASTRUCTURE ONLY — pure-torch emulation of the kernel interface in ``lass_b200/train_kernels.py``.

Every function has the signature and tensor layout of its CUDA counterpart and states the same arithmetic in torch ops.
Two uses: (i) the CPU tests hand it to ``TrainEngine(kernels=...)`` to check the ALGEBRA of the training step (the
decomposition of the backward pass into these kernels) against the autograd oracle without a GPU; (ii) the GPU tests
compare each CUDA kernel with its emulation on the same inputs.  Never imported by ``lass_b200``.
"""
import math

import torch
import torch.nn.functional as F

RAW_DTYPE = torch.float16
ACT_DTYPE = torch.float16
GRAD_DTYPE = torch.bfloat16
KIND_CONV, KIND_CONVT = 0, 1
SLOPE = 0.01


def set_exact(exact: bool):
    """exact = all 16-bit storage replaced by fp32 (isolates algebra errors from rounding)."""
    global RAW_DTYPE, ACT_DTYPE, GRAD_DTYPE
    if exact:
        RAW_DTYPE = ACT_DTYPE = GRAD_DTYPE = torch.float32
    else:
        RAW_DTYPE, ACT_DTYPE, GRAD_DTYPE = torch.float16, torch.float16, torch.bfloat16


def empty(shape, dtype, device):
    return torch.zeros(shape, dtype=dtype, device=device)


def _store(dst, coff, val):
    """Saturating store into a channel slice (fp16 raw stores saturate like the kernel's)."""
    C = val.shape[-1]
    if dst.dtype == torch.float16:
        val = val.clamp(-65504.0, 65504.0)
    dst[..., coff:coff + C] = val.to(dst.dtype)


class ConvSpec:
    def __init__(self, B, H, W, ncols, segs, bias=None, up=(1, 1), full_raw=None, full_raw_coff=0, pool=(1, 1),
                 pool_raw=None, after=None):
        self.B, self.H, self.W, self.ncols = B, H, W, ncols
        self.segs, self.bias, self.up = segs, bias, up
        self.full_raw, self.full_raw_coff = full_raw, full_raw_coff
        self.pool, self.pool_raw, self.after = pool, pool_raw, after


def conv(spec):
    acc = None
    for src, coff, cin, w, taps in spec.segs:
        x = src[..., coff:coff + cin].float().permute(0, 3, 1, 2)          # NCHW
        wt = w.float()                                                       # (taps, ncols, cin)
        if taps == 9:
            k = wt.reshape(3, 3, spec.ncols, cin).permute(2, 3, 0, 1)
            y = F.conv2d(x, k, padding=1)
        else:
            y = F.conv2d(x, wt[0][:, :, None, None])
        acc = y if acc is None else acc + y
    if spec.bias is not None:
        acc = acc + spec.bias.float()[None, :, None, None]
    uh, uw = spec.up
    gc = spec.ncols // (uh * uw)
    if uh * uw > 1:
        B, _, H, W = acc.shape
        acc = acc.reshape(B, uh, uw, gc, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, gc, H * uh, W * uw)
    out = acc.permute(0, 2, 3, 1)
    if spec.full_raw is not None:
        _store(spec.full_raw, spec.full_raw_coff, out)
    if spec.pool_raw is not None:
        _store(spec.pool_raw, 0, F.avg_pool2d(acc, spec.pool).permute(0, 2, 3, 1))
    if spec.after is not None:
        aw, ab, feat = spec.after
        feat.copy_(torch.einsum("bchw,kc->bkhw", acc, aw.float()) + ab.float()[None, :, None, None])


def stft(wave, basis_hi, basis_lo, n_fft, hop, workspace):
    conv_real, conv_imag = basis_hi, basis_lo      # the emulation is handed the reference's frozen conv weights
    x = F.pad(wave[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")
    real = F.conv1d(x, conv_real, stride=hop).transpose(1, 2)
    imag = F.conv1d(x, conv_imag, stride=hop).transpose(1, 2)
    mag = torch.clamp(real ** 2 + imag ** 2, 1e-10, math.inf) ** 0.5
    return mag.contiguous(), (real / mag).contiguous(), (imag / mag).contiguous()


def film(cond, film_w, film_b, out):
    out.copy_(F.linear(cond, film_w, film_b))


def bn0_stats(mag, sums):
    m = mag.double()
    sums[0] = m.sum(dim=(0, 1))
    sums[1] = (m * m).sum(dim=(0, 1))


def bn_stats(x, coff, C, sums):
    v = x[..., coff:coff + C].double().reshape(-1, C)
    sums[0] = v.sum(0)
    sums[1] = (v * v).sum(0)


def bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps, bnp):
    C = gamma.numel()
    mean = sums[0] / count
    var = (sums[1] / count - mean * mean).clamp_min(0.0)
    rstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * rstd
    bnp[0 * C:1 * C] = scale.float()
    bnp[1 * C:2 * C] = (beta.double() - mean * scale).float()
    bnp[2 * C:3 * C] = mean.float()
    bnp[3 * C:4 * C] = rstd.float()
    running_mean.mul_(1.0 - momentum).add_(momentum * mean.float())
    running_var.mul_(1.0 - momentum).add_(momentum * (var * (count / max(count - 1.0, 1.0))).float())


def _pre(x, x_coff, C, bnp, beta):
    xv = x[..., x_coff:x_coff + C].float()
    pre = bnp[0:C] * xv + bnp[C:2 * C] + beta[:, None, None, :C]
    return xv, pre


def bn_act(x, x_coff, out, out_coff, C, bnp, beta):
    _, pre = _pre(x, x_coff, C, bnp, beta)
    _store(out, out_coff, F.leaky_relu(pre, SLOPE))


def bn_bwd_reduce(dact, x, x_coff, C, bnp, beta, sums):
    xv, pre = _pre(x, x_coff, C, bnp, beta)
    g = dact[..., :C].float() * torch.where(pre > 0, 1.0, SLOPE)
    sums[:, :, 0] = g.sum(dim=(1, 2))
    sums[:, :, 1] = (g * (xv - bnp[2 * C:3 * C])).sum(dim=(1, 2))


def bn_bwd_finalize(sums, count, gamma, bnp, dgamma, dbeta, dfilm):
    B, C = sums.shape[0], sums.shape[1]
    t1 = sums[:, :, 0].double().sum(0)
    t2 = sums[:, :, 1].double().sum(0)
    rstd = bnp[3 * C:4 * C].double()
    scale = bnp[0:C].double()
    dbeta.copy_(t1.float())
    dgamma.copy_((rstd * t2).float())
    if dfilm is not None:
        dfilm.copy_(sums[:, :, 0])
    bnp[4 * C:5 * C] = (-scale * rstd * rstd * t2 / count).float()
    bnp[5 * C:6 * C] = (-scale * t1 / count).float()


def bn_bwd_totals(sums, totals):
    totals.copy_(sums.double().sum(0))


def bn_bwd_finalize_sync(sums, count_total, totals, gamma, bnp, dgamma, dbeta, dfilm):
    B, C = sums.shape[0], sums.shape[1]
    rstd = bnp[3 * C:4 * C].double()
    scale = bnp[0:C].double()
    dbeta.copy_(sums[:, :, 0].double().sum(0).float())
    dgamma.copy_((rstd * sums[:, :, 1].double().sum(0)).float())
    if dfilm is not None:
        dfilm.copy_(sums[:, :, 0])
    bnp[4 * C:5 * C] = (-scale * rstd * rstd * totals[:, 1] / count_total).float()
    bnp[5 * C:6 * C] = (-scale * totals[:, 0] / count_total).float()


def bn_bwd_apply(dact, x, x_coff, C, bnp, beta, add, add_coff, dx, dx_coff):
    xv, pre = _pre(x, x_coff, C, bnp, beta)
    g = dact[..., :C].float() * torch.where(pre > 0, 1.0, SLOPE)
    r = bnp[0:C] * g + bnp[4 * C:5 * C] * (xv - bnp[2 * C:3 * C]) + bnp[5 * C:6 * C]
    if add is not None:
        r = r + add[..., add_coff:add_coff + C].float()
    dx[..., dx_coff:dx_coff + C] = r.to(dx.dtype)


def pool_bwd(dpool, dskip, dskip_coff, dy, ph, pw):
    B, H, W, C = dy.shape
    up = dpool.float().repeat_interleave(ph, dim=1).repeat_interleave(pw, dim=2) / (ph * pw)
    if dskip is not None:
        up = up + dskip[..., dskip_coff:dskip_coff + C].float()
    dy.copy_(up.to(dy.dtype))


def unshuffle(src, src_coff, C, dst, uh, uw):
    B, H, W, _ = dst.shape
    s = src[..., src_coff:src_coff + C].reshape(B, H, uh, W, uw, C).permute(0, 1, 3, 2, 4, 5)
    dst.copy_(s.reshape(B, H, W, uh * uw * C))


def channel_sum(x, coff, C, out, acc=False):
    v = x[..., coff:coff + C].float().reshape(-1, C).sum(0)
    out.add_(v) if acc else out.copy_(v)


def wgrad(dy, dy_coff, co, x, x_coff, ci, taps, dw, acc=False):
    if acc:                                          # added to the pre-zeroed buffer
        tmp = torch.zeros_like(dw)
        wgrad(dy, dy_coff, co, x, x_coff, ci, taps, tmp)
        dw.add_(tmp)
        return
    g = dy[..., dy_coff:dy_coff + co].float()
    xv = x[..., x_coff:x_coff + ci].float()
    if x.dtype == torch.float16 and dy.dtype == torch.bfloat16:
        xv = xv.to(torch.bfloat16).float()          # the kernel converts fp16 sources to bf16 operands
    B, H, W, _ = g.shape
    dwv = dw.view(taps, co, ci)
    if taps == 1:
        dwv[0] = torch.einsum("bhwo,bhwi->oi", g, xv)
        return
    xp = F.pad(xv, (0, 0, 1, 1, 1, 1))
    for ky in range(3):
        for kx in range(3):
            dwv[ky * 3 + kx] = torch.einsum("bhwo,bhwi->oi", g, xp[:, ky:ky + H, kx:kx + W, :])


def pre_fwd(mag, bnp0, pre_w, pre_b, x0):
    B, T, Fq = mag.shape
    Tp, Fp = x0.shape[1], x0.shape[2]
    xbn = torch.zeros(B, Tp, Fp)
    xbn[:, :T] = (bnp0[0:Fq] * mag + bnp0[Fq:2 * Fq])[:, :, :Fp]
    _store(x0, 0, xbn[..., None] * pre_w + pre_b)


def pre_bwd(dx0, mag, bnp0, pre_w, dpre_w, dpre_b, dgamma0, dbeta0):
    B, T, Fq = mag.shape
    Tp, Fp = dx0.shape[1], dx0.shape[2]
    d = dx0.float()
    xbn = torch.zeros(B, Tp, Fp)
    xbn[:, :T] = (bnp0[0:Fq] * mag + bnp0[Fq:2 * Fq])[:, :, :Fp]
    dpre_b.copy_(d.sum(dim=(0, 1, 2)))
    dpre_w.copy_((d * xbn[..., None]).sum(dim=(0, 1, 2)))
    dxbn = (d[:, :T] * pre_w).sum(-1)                                        # (B, T, Fp)
    xhat = ((mag - bnp0[2 * Fq:3 * Fq]) * bnp0[3 * Fq:4 * Fq])[:, :, :Fp]
    dgamma0.zero_()
    dbeta0.zero_()
    dgamma0[:Fp] = (dxbn * xhat).sum(dim=(0, 1))
    dbeta0[:Fp] = dxbn.sum(dim=(0, 1))


def after_bwd(dfeat, y, after_w, dy, dw, db):
    df = dfeat.float()                                                       # (B, 3, H, W)
    dy.copy_(torch.einsum("bkhw,kc->bhwc", df, after_w.float()).to(dy.dtype))
    dw.copy_(torch.einsum("bkhw,bhwc->kc", df, y.float()))
    db.copy_(df.sum(dim=(0, 2, 3)))


def _window_sum(window, n_fft, hop, T):
    P = (T - 1) * hop + n_fft
    w2 = (window.double() ** 2)[None, :, None].repeat(1, 1, T)
    return F.fold(w2, (1, P), (1, n_fft), stride=(1, hop)).reshape(-1).clamp_min(1e-11)


def mask_istft(feat, mag, cos, sin, window, twiddle, n_fft, hop, length):
    B, T, Fq = mag.shape
    f = torch.zeros(B, 3, T, Fq)
    Ff = min(feat.shape[3], Fq)
    f[:, :, :, :Ff] = feat[:, :, :T, :Ff]
    mm = torch.sigmoid(f[:, 0])
    a, b = torch.tanh(f[:, 1]), torch.tanh(f[:, 2])
    r = torch.clamp((a * a + b * b) ** 0.5, 1e-10, math.inf)
    mc, ms = a / r, b / r
    oc = cos * mc - sin * ms
    osn = sin * mc + cos * ms
    om = F.relu(mag * mm)
    spec = torch.complex((om * oc).double(), (om * osn).double())            # (B, T, F)
    frames = torch.fft.irfft(spec, n=n_fft, dim=-1) * window.double()        # (B, T, n_fft)
    P = (T - 1) * hop + n_fft
    y = F.fold(frames.transpose(1, 2), (1, P), (1, n_fft), stride=(1, hop)).reshape(B, P)
    y = y / _window_sum(window, n_fft, hop, T)
    return y[:, n_fft // 2:n_fft // 2 + length].float().contiguous()


def istft_bwd(dwave, window, basis_hi, basis_lo, n_fft, hop, T, workspace, dre, dim):
    B, L = dwave.shape
    P = (T - 1) * hop + n_fft
    v = torch.zeros(B, P, dtype=torch.float64)
    v[:, n_fft // 2:n_fft // 2 + L] = dwave.double()
    v = v / _window_sum(window, n_fft, hop, T)
    frames = v.unfold(1, n_fft, hop) * window.double()                       # (B, T, n_fft)
    spec = torch.fft.rfft(frames, dim=-1)                                    # sum x w exp(-i...)
    dre.copy_(spec.real.float())
    dim.copy_(spec.imag.float())


def mask_bwd(feat, mag, cos, sin, dre, dim, dfeat, n_fft):
    B, T, Fq = mag.shape
    Fp = dfeat.shape[3]
    cf = torch.full((Fq,), 2.0 / n_fft)
    cf[0] = cf[-1] = 1.0 / n_fft
    gre, gim = (dre * cf)[:, :, :Fp], (dim * cf)[:, :, :Fp]
    x = feat[:, :, :T, :Fp]
    mg, cs, sn = mag[:, :, :Fp], cos[:, :, :Fp], sin[:, :, :Fp]
    m = torch.sigmoid(x[:, 0])
    a, b = torch.tanh(x[:, 1]), torch.tanh(x[:, 2])
    r = (a * a + b * b) ** 0.5
    rc = torch.clamp(r, 1e-10, math.inf)
    mc, ms = a / rc, b / rc
    cy = cs * mc - sn * ms
    sy = sn * mc + cs * ms
    absy = mg * m
    dabs = gre * cy + gim * sy
    dcy, dsy = gre * absy, gim * absy
    dmc = dcy * cs + dsy * sn
    dms = -dcy * sn + dsy * cs
    inv3 = 1.0 / (rc * rc * rc)
    ok = r > 1e-10
    da = torch.where(ok, b * inv3 * (dmc * b - dms * a), dmc / 1e-10)
    db = torch.where(ok, a * inv3 * (dms * a - dmc * b), dms / 1e-10)
    dfeat.zero_()
    dfeat[:, 0, :T] = dabs * mg * m * (1.0 - m)
    dfeat[:, 1, :T] = da * (1.0 - a * a)
    dfeat[:, 2, :T] = db * (1.0 - b * b)


def l1_loss(wave, target, dwave, loss_sum):
    d = wave - target
    loss_sum += d.abs().double().sum().float()
    dwave.copy_(torch.sign(d) / wave.numel())


def film_bwd(dbeta, cond, dw, db):
    dw.copy_(dbeta.t() @ cond)
    db.copy_(dbeta.sum(0))


def adamw_amsgrad(p, g, m, v, vmax, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    from oracle import train_oracle
    train_oracle.adamw_amsgrad_step(p, g * grad_scale if grad_scale != 1.0 else g, m, v, vmax, step, lr, beta1, beta2, eps,
                                    weight_decay)


def pack_weight(w, kind, fwd, dgrad):
    if kind == KIND_CONV:
        co, ci, kh, kw = w.shape
        taps = kh * kw
        wt = w.detach().reshape(co, ci, taps)
        if fwd is not None:
            fwd.copy_(wt.permute(2, 0, 1).to(fwd.dtype))
        if dgrad is not None:
            dgrad.copy_(wt.flip(2).permute(2, 1, 0).to(dgrad.dtype))
    else:
        ci, co, kh, kw = w.shape
        wt = w.detach().reshape(ci, co, kh * kw)
        if fwd is not None:
            fwd.copy_(wt.permute(2, 1, 0).reshape(1, kh * kw * co, ci).to(fwd.dtype))
        if dgrad is not None:
            dgrad.copy_(wt.permute(0, 2, 1).reshape(1, ci, kh * kw * co).to(dgrad.dtype))


def unpack_grad(dw, kind, grad):
    if kind == KIND_CONV:
        co, ci, kh, kw = grad.shape
        grad.copy_(dw.view(kh * kw, co, ci).permute(1, 2, 0).reshape(co, ci, kh, kw))
    else:
        ci, co, kh, kw = grad.shape
        grad.copy_(dw.view(kh * kw, co, ci).permute(2, 1, 0).reshape(ci, co, kh, kw))


# ---- fused / multi-tensor forms of the interface (same semantics as the separate calls) ----
def bn_stats_acc(x, coff, C, sums):
    v = x[..., coff:coff + C].double().reshape(-1, C)
    sums[0] += v.sum(0)
    sums[1] += (v * v).sum(0)


def bn_bwd_reduce_acc(dact, x, x_coff, C, bnp, beta, sums):
    xv, pre = _pre(x, x_coff, C, bnp, beta)
    g = dact[..., :C].float() * torch.where(pre > 0, 1.0, SLOPE)
    sums[:, :, 0] += g.sum(dim=(1, 2))
    sums[:, :, 1] += (g * (xv - bnp[2 * C:3 * C])).sum(dim=(1, 2))


def bn_bwd_reduce_finalize(dact, x, x_coff, C, bnp, beta, sums, counter, gamma, dgamma, dbeta, dfilm):
    bn_bwd_reduce(dact, x, x_coff, C, bnp, beta, sums)
    bn_bwd_finalize(sums, x.shape[0] * x.shape[1] * x.shape[2], gamma, bnp, dgamma, dbeta, dfilm)


def pack_weights_table(items, device):
    return items


def pack_weights(table):
    for w, kind, fwd, dgrad in table:
        pack_weight(w, kind, fwd, dgrad)


def unpack_grads_table(items, device):
    return items


def unpack_grads(table):
    for dw, kind, grad in table:
        unpack_grad(dw, kind, grad)
