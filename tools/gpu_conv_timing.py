"""Timing experiments on single conv launches configured exactly like the ResUNet30 plan's (B200): which role of the
kernel bounds a layer?  Usage: python tools/gpu_conv_timing.py [B] [name-substring ...]
Debug flags (lass_debug_set_conv_flags): 1 epilogue idle, 2 no MMA, 4 no A loads, 16 no stores, 64 MMA issuers only,
128 single MMA issuer, 256 generic (unspecialised) epilogue."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LASS_B200_LIB", os.path.join(ROOT, "lass_b200", "_lib", "liblass_b200_prof.so"))   # `make prof`
from lass_b200 import _cabi, ops, packing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
FILTER = [a for a in sys.argv[2:]]
FLAGS = tuple(int(x) for x in os.environ.get("LASS_TIMING_FLAGS", "0,1,2,256").split(","))
ROUNDS = int(os.environ.get("LASS_TIMING_ROUNDS", "3"))
dev = "cuda"
# kind: c1 = first conv of a block (one activated output), enc2 = encoder conv2 (+1x1 shortcut or rank-1 residual; raw + act
# skip into the concat buffers, pooled raw + act), dec2 = decoder conv2 (+ shortcut over the raw concat), up = transposed
# conv into the concat buffers, last = decoder_block6 conv2 + after_conv
LAYERS = {
    "enc0.c1 32->32 @1024x512": dict(kind="c1", H=1024, W=512, cin=32, cout=32),
    "enc0.c1gen 1->32->32 @1024x512": dict(kind="c1", H=1024, W=512, cin=32, cout=32, gen=True),
    "enc0.c2 32->32+resid @1024x512": dict(kind="enc2", H=1024, W=512, cin=32, cout=32, sc=0),
    "enc1.c1 32->64 @512x256": dict(kind="c1", H=512, W=256, cin=32, cout=64),
    "enc1.c2 64->64+sc32 @512x256": dict(kind="enc2", H=512, W=256, cin=64, cout=64, sc=32),
    "enc2.c1 64->128 @256x128": dict(kind="c1", H=256, W=128, cin=64, cout=128),
    "enc2.c2 128->128+sc64 @256x128": dict(kind="enc2", H=256, W=128, cin=128, cout=128, sc=64),
    "enc3.c1 128->256 @128x64": dict(kind="c1", H=128, W=64, cin=128, cout=256),
    "enc3.c2 256->256+sc128 @128x64": dict(kind="enc2", H=128, W=64, cin=256, cout=256, sc=128),
    "enc4.c1 256->384 @64x32": dict(kind="c1", H=64, W=32, cin=256, cout=384),
    "enc4.c2 384->384+sc256 @64x32": dict(kind="enc2", H=64, W=32, cin=384, cout=384, sc=256),
    "dec1.c1 768->384 @64x32": dict(kind="c1", H=64, W=32, cin=768, cout=384),
    "dec1.c2 384->384+sc768 @64x32": dict(kind="dec2", H=64, W=32, cin=384, cout=384, sc=768),
    "dec2.c1 512->256 @128x64": dict(kind="c1", H=128, W=64, cin=512, cout=256),
    "dec2.c2 256->256+sc512 @128x64": dict(kind="dec2", H=128, W=64, cin=256, cout=256, sc=512),
    "dec3.up 256->128x4 @64x32": dict(kind="up", H=64, W=32, cin=256, cout=128),
    "dec3.c1 256->128 @256x128": dict(kind="c1", H=256, W=128, cin=256, cout=128),
    "dec3.c2 128->128+sc256 @256x128": dict(kind="dec2", H=256, W=128, cin=128, cout=128, sc=256),
    "dec4.up 128->64x4 @256x128": dict(kind="up", H=256, W=128, cin=128, cout=64),
    "dec4.c1 128->64 @512x256": dict(kind="c1", H=512, W=256, cin=128, cout=64),
    "dec4.c2 64->64+sc128 @512x256": dict(kind="dec2", H=512, W=256, cin=64, cout=64, sc=128),
    "dec5.up 64->32x4 @512x256": dict(kind="up", H=512, W=256, cin=64, cout=32),
    "dec5.c1 64->32 @1024x512": dict(kind="c1", H=1024, W=512, cin=64, cout=32),
    "dec5.c2 32->32+sc64+after @1024x512": dict(kind="last", H=1024, W=512, cin=32, cout=32, sc=64),
}
NAMES = ["prod_wait_a_empty", "prod_wait_b_empty", "prod_total", "mma_wait_acc_empty", "mma_wait_a_full",
         "mma_wait_b_full", "mma_total", "epi_wait_acc_full", "epi_total"]


def build_layer(kind, H, W, cin, cout, sc=0, gen=False):
    """-> (ncols, segments, kwargs, flops, keep-alive list)"""
    keep = []
    src = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(B, cout, device=dev) * 0.1
    keep += [src, scale, shift]
    kw = {}
    if kind == "up":
        w = packing.pack_convT_weight(torch.randn(cin, cout, 2, 2, device=dev) / cin ** 0.5, torch.bfloat16)
        segs = [ops.make_segment(src, 0, cin, w, 1)]
        raw = torch.empty(B, 2 * H, 2 * W, 2 * cout, dtype=torch.float16, device=dev)
        act = torch.empty(B, 2 * H, 2 * W, 2 * cout, dtype=torch.bfloat16, device=dev)
        kw.update(up=(2, 2), full_raw=ops.make_out(raw, 0), full_act=ops.make_out(act, 0, scale, shift))
        keep += [w, raw, act]
        return 4 * cout, segs, kw, 2.0 * B * H * W * 4 * cout * cin, keep
    w = packing.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5), torch.bfloat16)
    segs = [ops.make_segment(src, 0, cin, w, 9)]
    keep.append(w)
    if sc:
        rawin = torch.randn(B, H, W, sc, device=dev).to(torch.float16)
        wsc = packing.pack_conv_weight(torch.randn(cout, sc, 1, 1, device=dev) / sc ** 0.5, torch.float16)
        segs.append(ops.make_segment(rawin, 0, sc, wsc, 1))
        kw["bias"] = torch.randn(cout, device=dev) * 0.1
        keep += [rawin, wsc, kw["bias"]]
    flops = 2.0 * B * H * W * cout * (9 * cin + sc)
    if gen:   # encoder_block1 conv1: operand generated from the magnitude (bn0 + pre_conv + BN + FiLM + lrelu)
        T, F = H - 23, W + 1
        g = (torch.randn(B, T, F, device=dev), torch.rand(F, device=dev) + 0.5, torch.randn(F, device=dev) * 0.1,
             torch.randn(cin, device=dev), torch.randn(cin, device=dev) * 0.1, torch.rand(cin, device=dev) + 0.5,
             torch.randn(B, cin, device=dev) * 0.1)
        kw["gen"] = g
        keep += list(g)
    if kind in ("c1", "dec2"):
        act = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=dev)
        kw["full_act"] = ops.make_out(act, 0, scale, shift)
        keep.append(act)
    elif kind == "enc2":
        raw = torch.empty(B, H, W, 2 * cout, dtype=torch.float16, device=dev)
        act = torch.empty(B, H, W, 2 * cout, dtype=torch.bfloat16, device=dev)
        pr = torch.empty(B, H // 2, W // 2, cout, dtype=torch.float16, device=dev)
        pa = torch.empty(B, H // 2, W // 2, cout, dtype=torch.bfloat16, device=dev)
        kw.update(full_raw=ops.make_out(raw, cout), full_act=ops.make_out(act, cout, scale, shift), pool=(2, 2),
                  pool_raw=ops.make_out(pr, 0), pool_act=ops.make_out(pa, 0, scale, shift))
        keep += [raw, act, pr, pa]
        if not sc:   # encoder_block1: rank-1 identity residual regenerated from the magnitude
            T, F = H - 23, W + 1
            r = (torch.randn(B, T, F, device=dev), torch.rand(F, device=dev) + 0.5, torch.randn(F, device=dev) * 0.1,
                 torch.randn(cout, device=dev), torch.randn(cout, device=dev) * 0.1)
            kw["resid"] = r
            keep += list(r)
    elif kind == "last":
        aw, ab = torch.randn(3, cout, device=dev) * 0.2, torch.randn(3, device=dev) * 0.1
        feat = torch.empty(B, 3, H, W, device=dev)
        kw.update(after_w=aw, after_b=ab, feat=feat)
        keep += [aw, ab, feat]
    return cout, segs, kw, flops, keep


def bench_layer(cfg):
    ncols, segs, kw, flops, keep = build_layer(**cfg)
    H, W = cfg["H"], cfg["W"]
    lib = _cabi.load()
    res = {}
    prof = torch.zeros(296 * 16, dtype=torch.int64, device=dev)
    profiling = not os.environ.get("LASS_NO_PROFILE_RUN")
    # the flag settings are interleaved over ROUNDS rounds (best round kept): clock drift under the power cap would otherwise
    # bias an A/B comparison towards whichever setting ran first
    for rnd in range(ROUNDS):
        for flags in FLAGS:
            lib.lass_debug_set_conv_flags(flags)
            for _ in range(2 if rnd == 0 else 1):
                ops.conv_igemm(B, H, W, ncols, segs, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.conv_igemm(B, H, W, ncols, segs, **kw)
            e1.record()
            torch.cuda.synchronize()
            t = round(e0.elapsed_time(e1) / 3, 4)
            res["flags%d" % flags] = min(t, res.get("flags%d" % flags, 1e9))
    for flags in FLAGS:
        lib.lass_debug_set_conv_flags(flags)
        if profiling and flags in (0, 1, 64):
            prof.zero_()
            _cabi.check(lib.lass_debug_set_conv_profile(prof.data_ptr()))
            ops.conv_igemm(B, H, W, ncols, segs, **kw)
            torch.cuda.synchronize()
            lib.lass_debug_set_conv_profile(None)
            pr = prof.view(296, 16).cpu().double()
            pr = pr[pr[:, 9] > 0]
            if pr.shape[0]:
                it = pr[:, 9].mean().item()
                res["profile_flags%d" % flags] = {n: round(pr[:, i].mean().item() / it) for i, n in enumerate(NAMES)}
                res["items_per_cta"] = it
    lib.lass_debug_set_conv_flags(0)
    res["tflops"] = round(flops / (res["flags0"] * 1e-3) / 1e12, 1)
    return res


if __name__ == "__main__":
    out = {}
    for name, cfg in LAYERS.items():
        if FILTER and not any(f in name for f in FILTER):
            continue
        out[name] = bench_layer(cfg)
        r = out[name]
        print("%-38s" % name, " ".join("%s=%.3f" % (k[5:], v) for k, v in r.items() if k.startswith("flags")), "TF=%.0f" % r["tflops"],
              flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_timing_b%d.json" % B), "w"), indent=1)
