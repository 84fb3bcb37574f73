"""Timing experiments on single conv layers (B200): which role of the kernel bounds a layer?
Usage: python tools/gpu_conv_timing.py [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LASS_B200_LIB", os.path.join(ROOT, "lass_b200", "_lib", "liblass_b200_prof.so"))   # `make prof`
from lass_b200 import _cabi, ops, packing  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
FLAGS = (0, 1, 2, 256)
dev = "cuda"
LAYERS = {
    "dec5.up 64->32x4 @512x256": (512, 256, 64, 32, 0, 2, False, (2, 2)),
    "dec4.up 128->64x4 @256x128": (256, 128, 128, 64, 0, 2, False, (2, 2)),
    "enc0.c2 32->32+id @1024x512 4out": (1024, 512, 32, 32, 32, 2, True),
    # name: (H, W, cin, cout, shortcut_cin, n_outputs(act only=1), pool)
    "enc0.c1 32->32 @1024x512": (1024, 512, 32, 32, 0, 1, False),
    "dec5.c1 64->32 @1024x512": (1024, 512, 64, 32, 0, 1, False),
    "enc1.c2 64->64+sc32 @512x256 4out": (512, 256, 64, 64, 32, 2, True),
    "dec4.c1 128->64 @512x256": (512, 256, 128, 64, 0, 1, False),
    "enc2.c2 128->128+sc64 @256x128 4out": (256, 128, 128, 128, 64, 2, True),
    "dec3.c1 256->128 @256x128": (256, 128, 256, 128, 0, 1, False),
    "dec2.c1 512->256 @128x64": (128, 64, 512, 256, 0, 1, False),
}


def bench_layer(H, W, cin, cout, sc, nout, pool, up=(1, 1), algo=0):
    src = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    nup = up[0] * up[1]
    if nup == 1:
        w = (packing.pack_conv_weight_dxn if algo == 1 else packing.pack_conv_weight)(
            torch.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5), torch.bfloat16)
        segs = [ops.make_segment(src, 0, cin, w, 9)]
    else:
        w = packing.pack_convT_weight(torch.randn(cin, cout, up[0], up[1], device=dev) / cin ** 0.5, torch.bfloat16)
        segs = [ops.make_segment(src, 0, cin, w, 1)]
    keep = [src, w]
    if sc:
        raw = torch.randn(B, H, W, sc, device=dev).to(torch.float16)
        wsc = packing.pack_conv_weight(torch.randn(cout, sc, 1, 1, device=dev) / sc ** 0.5, torch.float16)
        segs.append(ops.make_segment(raw, 0, sc, wsc, 1))
        keep += [raw, wsc]
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(B, cout, device=dev) * 0.1
    cbuf = cout * (2 if nup > 1 else 1)
    act = torch.empty(B, H * up[0], W * up[1], cbuf, dtype=torch.bfloat16, device=dev)
    kw = dict(full_act=ops.make_out(act, 0, scale, shift), up=up, algo=algo)
    if nout > 1:
        rawo = torch.empty(B, H * up[0], W * up[1], cbuf, dtype=torch.float16, device=dev)
        kw["full_raw"] = ops.make_out(rawo, 0)
        keep.append(rawo)
    if pool:
        pr = torch.empty(B, H // 2, W // 2, cout, dtype=torch.float16, device=dev)
        pa = torch.empty(B, H // 2, W // 2, cout, dtype=torch.bfloat16, device=dev)
        kw.update(pool=(2, 2), pool_raw=ops.make_out(pr, 0), pool_act=ops.make_out(pa, 0, scale, shift))
        keep += [pr, pa]
    res = {}
    # role profile (cycles per item, averaged over CTAs)
    prof = torch.zeros(296 * 16, dtype=torch.int64, device=dev)
    lib = _cabi.load()
    if not os.environ.get("LASS_NO_PROFILE_RUN"):
        _cabi.check(lib.lass_debug_set_conv_profile(prof.data_ptr()))
        ops.conv_igemm(B, H, W, cout * nup, segs, **kw)
        torch.cuda.synchronize()
        lib.lass_debug_set_conv_profile(None)
    pr = prof.view(296, 16).cpu().double()
    pr = pr[pr[:, 9] > 0]
    if pr.shape[0] == 0:
        pr = torch.ones(1, 16, dtype=torch.float64)
    items = pr[:, 9].mean().item()
    names = ["prod_wait_a_empty", "prod_wait_b_empty", "prod_total", "mma_wait_acc_empty", "mma_wait_a_full",
             "mma_wait_b_full", "mma_total", "epi_wait_acc_full", "epi_total"]
    res["profile_cyc_per_item"] = {n: round(pr[:, i].mean().item() / items) for i, n in enumerate(names)}
    res["items_per_cta"] = items
    res["ctas"] = int(pr.shape[0])
    for flags in FLAGS:
        _cabi.load().lass_debug_set_conv_flags(flags)
        for _ in range(2):
            ops.conv_igemm(B, H, W, cout * nup, segs, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.conv_igemm(B, H, W, cout * nup, segs, **kw)
        e1.record()
        torch.cuda.synchronize()
        res["flags%d" % flags] = round(e0.elapsed_time(e1) / 3, 4)
        if flags in (1, 64) and not os.environ.get("LASS_NO_PROFILE_RUN"):
            prof.zero_()
            lib.lass_debug_set_conv_profile(prof.data_ptr())
            ops.conv_igemm(B, H, W, cout * nup, segs, **kw)
            torch.cuda.synchronize()
            lib.lass_debug_set_conv_profile(None)
            pr = prof.view(296, 16).cpu().double()
            pr = pr[pr[:, 9] > 0]
            if pr.shape[0]:
                it = pr[:, 9].mean().item()
                res["profile_flags%d" % flags] = {n: round(pr[:, i].mean().item() / it) for i, n in enumerate(names)}
    _cabi.load().lass_debug_set_conv_flags(0)
    flops = 2.0 * B * H * W * cout * nup * ((9 if nup == 1 else 1) * cin + sc)
    res["tflops_normal"] = round(flops / (res["flags0"] * 1e-3) / 1e12, 1)
    return res


out = {}
if len(sys.argv) > 2 and sys.argv[2] == "dxn":
    DXN = {k: v for k, v in LAYERS.items() if "up " not in k and v[3] <= 64}
    DXN["dec4.c2 64->64+sc128 @512x256"] = (512, 256, 64, 64, 128, 1, False)
    DXN["dec5.c2 32->32+sc64 @1024x512"] = (1024, 512, 32, 32, 64, 1, False)
    DXN["enc1.c1 32->64 @512x256"] = (512, 256, 32, 64, 0, 1, False)
    LAYERS = {}
    for k, v in DXN.items():
        LAYERS[k + " [K]"] = v
        LAYERS[k + " [dxN]"] = tuple(v) + ((1, 1), 1)
for name, cfg in LAYERS.items():
    out[name] = bench_layer(*cfg)
    print(name, json.dumps(out[name]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_timing_b%d.json" % B), "w"), indent=1)
