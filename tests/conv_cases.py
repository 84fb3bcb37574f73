"""Case table + runner for the tcgen05 implicit-GEMM conv kernel (K3 / K4): many small cases against torch fp32
convolutions on the same 16-bit operands (tests/test_gpu_conv.py parametrises over CASES).
Stand-alone bring-up use: python tests/conv_cases.py <tag> [case ...] -> gpurun_out/conv_probe_p<tag>.json"""
import json
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200 import _cabi, ops, packing  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def nhwc(x):  # (B,C,H,W) fp32 -> (B,H,W,C)
    return x.permute(0, 2, 3, 1).contiguous()


def lrelu_affine(v, scale, shift):  # v (B,C,H,W); scale (C); shift (B,C)
    return F.leaky_relu(v * scale[None, :, None, None] + shift[:, :, None, None], 0.01)


def run_case(name, B, H, W, cin, cout, src_dtype=torch.bfloat16, shortcut_cin=0, up=(1, 1), pool=(1, 1),
             want_raw=True, want_act=True, want_pool=False, after=False, bias=False, out_cstride_mult=1, out_coff=0,
             src_extra=0, seed=0, resid=False, algo=0, gen=False, flags=0, shrink=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    taps = 9 if up == (1, 1) else 1
    # source buffer with optional extra channels (reads a channel slice)
    cbuf = cin + src_extra
    src_full = r(B, cbuf, H, W).to(src_dtype)
    src = nhwc(src_full.float()).to(src_dtype).to(dev)
    coff = src_extra
    x = src_full.float()[:, coff:coff + cin].to(dev)
    nup = up[0] * up[1]
    if taps == 9:
        w = (r(cout, cin, 3, 3) / (3.0 * cin ** 0.5)).to(src_dtype)
        wp = packing.pack_conv_weight(w.float(), src_dtype).to(dev)
        ref = F.conv2d(x, w.float().to(dev), None, padding=1)
    else:
        w = (r(cin, cout, up[0], up[1]) / cin ** 0.5).to(src_dtype)
        wp = packing.pack_convT_weight(w.float(), src_dtype).to(dev)
        ref = F.conv_transpose2d(x, w.float().to(dev), None, stride=up)
    gen_arg = None
    if gen:
        # generated A operand: x (B, cin = 32, H, W) = bf16(lrelu(gsc * (gw * m + gb) + gsh[b])) of a 1-channel map m with
        # T < H valid rows (zero rows after the input affine) and F = W + 1 columns; the conv then reads no activation tensor
        assert cin == 32 and taps == 9 and src_dtype == torch.bfloat16
        T, Fm = H - 3, W + 1
        g_src = r(B, T, Fm).to(dev)
        g_isc, g_ish = (0.5 + torch.rand(Fm, generator=g)).to(dev), (0.1 * r(Fm)).to(dev)
        g_w, g_b = r(cin).to(dev), (0.1 * r(cin)).to(dev)
        g_sc = (0.5 + torch.rand(cin, generator=g)).to(dev)
        g_sh = (0.2 * r(B, cin + 16)).to(dev)[:, 8:8 + cin]
        m = torch.zeros(B, H, W, device=dev)
        m[:, :T] = (g_src * g_isc + g_ish)[:, :, :W]
        pre = g_w[None, :, None, None] * m[:, None] + g_b[None, :, None, None]
        x = lrelu_affine(pre, g_sc, g_sh).to(torch.bfloat16).float()
        ref = F.conv2d(x, w.float().to(dev), None, padding=1)
        gen_arg = (g_src, g_isc, g_ish, g_w, g_b, g_sc, g_sh)
    segs = [ops.make_segment(src, coff, cin, wp, taps)]
    keep = [src, wp]
    if shortcut_cin:
        sc_src_full = r(B, shortcut_cin, H, W).to(torch.float16)
        sc_src = nhwc(sc_src_full.float()).to(torch.float16).to(dev)
        wsc = (r(cout, shortcut_cin, 1, 1) / shortcut_cin ** 0.5).to(torch.float16)
        wscp = packing.pack_conv_weight(wsc.float(), torch.float16).to(dev)
        segs.append(ops.make_segment(sc_src, 0, shortcut_cin, wscp, 1))
        ref = ref + F.conv2d(sc_src_full.float().to(dev), wsc.float().to(dev), None)
        keep += [sc_src, wscp]
    resid_arg = None
    if resid:
        # rank-1 residual regenerated from a 1-channel map with T < H (zero rows after the input affine) and F = W + 1
        T, Fm = H - 5, W + 1
        r_src = r(B, T, Fm).to(dev)
        r_isc, r_ish = (0.5 + torch.rand(Fm, generator=g)).to(dev), (0.1 * r(Fm)).to(dev)
        r_w, r_b = r(cout).to(dev), (0.1 * r(cout)).to(dev)
        xmap = torch.zeros(B, H, W, device=dev)
        xmap[:, :T] = (r_src * r_isc + r_ish)[:, :, :W]
        ref = ref + r_w[None, :, None, None] * xmap[:, None] + r_b[None, :, None, None]
        resid_arg = (r_src, r_isc, r_ish, r_w, r_b)
        keep += list(resid_arg)
    bias_t = None
    if bias:
        bias_t = (0.1 * r(cout * nup)).to(dev)
        # bias is per GEMM column; for conv (nup == 1) per cout
        ref = ref + bias_t[:cout][None, :, None, None] if nup == 1 else ref
    Ho, Wo = H * up[0], W * up[1]
    ccap = cout * out_cstride_mult
    scale = (0.5 + torch.rand(cout, generator=g)).to(dev)
    shift = (0.2 * r(B, cout + 16)).to(dev)[:, 8:8 + cout]  # non-contiguous rows: exercises shift_bstride
    shift = shift.contiguous() if False else shift
    outs = {}
    kw = {}
    if want_raw:
        outs["raw"] = torch.full((B, Ho, Wo, ccap), 7.0, dtype=torch.float16, device=dev)
        kw["full_raw"] = ops.make_out(outs["raw"], out_coff)
    if want_act:
        outs["act"] = torch.full((B, Ho, Wo, ccap), 7.0, dtype=torch.bfloat16, device=dev)
        kw["full_act"] = ops.make_out(outs["act"], out_coff, scale, shift)
    if want_pool:
        Hp, Wp = H // pool[0], W // pool[1]
        outs["praw"] = torch.full((B, Hp, Wp, cout), 7.0, dtype=torch.float16, device=dev)
        outs["pact"] = torch.full((B, Hp, Wp, cout), 7.0, dtype=torch.bfloat16, device=dev)
        kw["pool_raw"] = ops.make_out(outs["praw"], 0)
        kw["pool_act"] = ops.make_out(outs["pact"], 0, scale, shift)
        kw["pool"] = pool
    if after:
        aw = (r(3, cout) / cout ** 0.5).to(dev)
        ab = (0.1 * r(3)).to(dev)
        outs["feat"] = torch.full((B, 3, H, W), 7.0, dtype=torch.float32, device=dev)
        kw.update(after_w=aw, after_b=ab, feat=outs["feat"])
    # shift tensor must be addressable with a row stride: pass the strided view directly
    # the cases are small grids: without `shrink` they keep the tile the layer gets at the bench size (flag 262144); the
    # "shrink_" cases run the narrower small-grid tiles the same shapes get at small batches
    flags = flags | (0 if shrink else 262144)
    if flags:
        _cabi.check(_cabi.load().lass_debug_set_conv_flags(flags))   # e.g. 4096: no CTA pairs (the single-CTA streamed path)
    try:
        ops.conv_igemm(B, H, W, cout * nup, segs, bias=bias_t, up=up, resid=resid_arg, algo=algo, gen=gen_arg, **kw)
        torch.cuda.synchronize()
    finally:
        if flags:
            _cabi.check(_cabi.load().lass_debug_set_conv_flags(0))
    res = {}
    refn = nhwc(ref)
    sl = slice(out_coff, out_coff + cout)
    scale_ref = float(ref.abs().max())
    if want_raw:
        res["raw"] = float((outs["raw"][..., sl].float() - refn).abs().max()) / scale_ref
        if out_cstride_mult > 1:
            other = torch.ones(ccap, dtype=torch.bool)
            other[sl] = False
            res["raw_untouched"] = bool((outs["raw"][..., other.to(dev)] == 7.0).all())
    if want_act:
        a_ref = nhwc(lrelu_affine(ref, scale, shift))
        res["act"] = float((outs["act"][..., sl].float() - a_ref).abs().max()) / float(a_ref.abs().max())
    if want_pool:
        p_ref = F.avg_pool2d(ref, kernel_size=pool)
        res["praw"] = float((outs["praw"].float() - nhwc(p_ref)).abs().max()) / scale_ref
        pa_ref = nhwc(lrelu_affine(p_ref, scale, shift))
        res["pact"] = float((outs["pact"].float() - pa_ref).abs().max()) / float(pa_ref.abs().max())
    if after:
        f_ref = F.conv2d(ref, aw[:, :, None, None], ab)
        res["feat"] = float((outs["feat"] - f_ref).abs().max()) / float(f_ref.abs().max())
    return res


CASES = {
    "c32_32": dict(B=2, H=32, W=16, cin=32, cout=32),
    "c32_32_mt1": dict(B=2, H=48, W=24, cin=32, cout=32),
    "c32_64": dict(B=1, H=32, W=16, cin=32, cout=64),
    "c64_64": dict(B=2, H=64, W=32, cin=64, cout=64),
    "c64_32": dict(B=1, H=32, W=32, cin=64, cout=32),
    "c128_128": dict(B=2, H=32, W=16, cin=128, cout=128),
    "c256_384": dict(B=1, H=32, W=8, cin=256, cout=384),
    "c768_384": dict(B=1, H=32, W=16, cin=768, cout=384, want_raw=False),
    "c256_256": dict(B=1, H=32, W=16, cin=256, cout=256, want_raw=False),
    "pool32_wide": dict(B=2, H=64, W=24, cin=32, cout=32, shortcut_cin=32, want_pool=True, pool=(2, 2)),
    "resid_pool": dict(B=2, H=64, W=16, cin=32, cout=32, resid=True, want_pool=True, pool=(2, 2), out_cstride_mult=2, out_coff=32),
    "sc_pool": dict(B=2, H=32, W=16, cin=64, cout=64, shortcut_cin=32, bias=True, want_pool=True, pool=(2, 2)),
    "sc_pool12": dict(B=1, H=32, W=16, cin=384, cout=384, shortcut_cin=384, bias=True, want_pool=True, pool=(1, 2)),
    "sc_big": dict(B=1, H=32, W=16, cin=128, cout=128, shortcut_cin=256, bias=True),
    "convT22": dict(B=2, H=16, W=16, cin=64, cout=32, up=(2, 2), out_cstride_mult=2, out_coff=0),
    "convT22_big": dict(B=1, H=32, W=16, cin=384, cout=256, up=(2, 2), out_cstride_mult=2, out_coff=0),
    "convT12": dict(B=1, H=32, W=8, cin=384, cout=384, up=(1, 2), out_cstride_mult=2, out_coff=0),
    "slice_out": dict(B=1, H=32, W=16, cin=32, cout=32, out_cstride_mult=2, out_coff=32, src_extra=32),
    "after": dict(B=2, H=32, W=16, cin=32, cout=32, shortcut_cin=64, bias=True, after=True, want_raw=False, want_act=False),
    "partial": dict(B=1, H=40, W=12, cin=64, cout=64),
    "fp16src": dict(B=1, H=32, W=16, cin=64, cout=64, src_dtype=torch.float16),
    "gen_a": dict(B=3, H=64, W=24, cin=32, cout=32, gen=True, want_raw=False),
    "gen_a_mt1": dict(B=2, H=48, W=12, cin=32, cout=32, gen=True, want_raw=False),
}

# Streamed weights with N >= 128 run as CTA pairs (tcgen05 cta_group::2) when the pixel tiles pair up: longer item
# sequences per pair, several N tiles, pooled / sliced / transposed outputs -- and the single-CTA streamed path (debug flag 4096)
CASES["pair_c128_128_long"] = dict(B=6, H=128, W=64, cin=128, cout=128, want_raw=False)
CASES["pair_c256_256_long"] = dict(B=5, H=64, W=64, cin=256, cout=256, want_raw=False)
CASES["pair_c512_256"] = dict(B=2, H=32, W=32, cin=512, cout=256, shortcut_cin=512, bias=True)
CASES["pair_c384_384_sc"] = dict(B=4, H=32, W=16, cin=384, cout=384, shortcut_cin=256, bias=True, want_pool=True, pool=(2, 2))
CASES["pair_slice"] = dict(B=2, H=32, W=32, cin=256, cout=128, out_cstride_mult=2, out_coff=128)
for _name in ("c128_128", "c768_384", "c256_256", "sc_pool12", "sc_big", "pair_c384_384_sc", "pair_c512_256"):
    CASES["nopair_" + _name] = dict(CASES[_name], flags=4096)
# streamed 3x3 weights move in row stages (three taps per TMA box / handshake) where three such stages fit; flag 8192: one tap
# per stage (pairs and single CTAs)
for _name in ("c128_128", "c256_256", "sc_big", "pair_c384_384_sc", "pair_c512_256"):
    CASES["tapstage_" + _name] = dict(CASES[_name], flags=8192)
    CASES["nopair_tapstage_" + _name] = dict(CASES[_name], flags=4096 | 8192)
# resident weights in CTA pairs (experimental, debug flag 32768)
for _name in ("c32_32", "c64_64", "c64_32", "pool32_wide", "resid_pool", "sc_pool", "after", "fp16src", "slice_out", "c32_64"):
    CASES["respair_" + _name] = dict(CASES[_name], flags=32768)
CASES["respair_c128_64_long"] = dict(B=4, H=64, W=64, cin=128, cout=64, want_raw=False, flags=32768)
CASES["respair_c64_32_long"] = dict(B=4, H=128, W=64, cin=64, cout=32, want_raw=False, flags=32768)

# small-grid tiles (one m-tile, N down to 32) for streamed-weight layers: plain, shortcut segment, pooled / sliced outputs, fp16
for _name in ("c128_128", "c256_384", "c768_384", "c256_256", "sc_pool12", "sc_big", "pair_c384_384_sc", "pair_c512_256", "pair_slice"):
    CASES["shrink_" + _name] = dict(CASES[_name], shrink=True)
CASES["shrink_c384_384_b1"] = dict(B=1, H=32, W=8, cin=384, cout=384, shortcut_cin=384, bias=True, shrink=True)
CASES["shrink_c128_256_pool"] = dict(B=1, H=64, W=32, cin=128, cout=256, shortcut_cin=128, bias=True, want_pool=True, pool=(2, 2), shrink=True)

if __name__ == "__main__":
    pitch = int(sys.argv[1])   # kept for the log name; the halo pitch is fixed at 10 pixels
    names = sys.argv[2:] or list(CASES)
    report = {}
    for n in names:
        try:
            report[n] = run_case(n, **CASES[n])
        except Exception as exc:
            report[n] = {"error": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}
            print(n, report[n], flush=True)
            if "CUDA" in repr(exc) or "cuda" in repr(exc):
                break
        print(n, json.dumps(report[n]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conv_probe_p%d.json" % pitch), "w") as f:
        json.dump(report, f, indent=1)
