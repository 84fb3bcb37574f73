"""tcgen05.mma throughput with an unrolled issue loop (B200): cycles per MMA for M=128, K=16 vs N / swizzle / accumulators."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200 import _cabi
lib = _cabi.load_debug()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
res = {}
iters = 8192
for grid in (148,):
    for n in (32, 64, 96, 128, 256):
        for name, kc, swz, start, sbo in (("sw128", 64, 2, 0, 1024), ("sw64", 32, 4, 0, 512), ("sw64_dense16shift", 32, 4, 16 * 64, 512),
                                          ("sw128_pitch10", 64, 2, 0, 1280), ("sw128_pitch10_shift11", 64, 2, 11 * 128, 1280),
                                          ("sw64_pitch10", 32, 4, 0, 640), ("sw64_pitch10_shift11", 32, 4, 11 * 64, 640),
                                          ("sw64_pitch10_shift1", 32, 4, 64, 640), ("sw64_pitch12", 32, 4, 0, 768),
                                          ("sw64_pitch16", 32, 4, 0, 1024), ("sw128_pitch12", 64, 2, 0, 1536), ("sw128_pitch16", 64, 2, 0, 2048),
                                          ("sw64_shift1", 32, 4, 64, 512), ("sw128_shift1", 64, 2, 128, 1024)):
            for nacc in (1, 2):
                if nacc * n > 512:
                    continue
                _cabi.check(lib.lass_debug_umma_bench2(n, kc, swz, start, sbo, iters, nacc, grid, out.data_ptr(), None))
                torch.cuda.synchronize()
                res["g%d_n%d_%s_acc%d" % (grid, n, name, nacc)] = round(out[:grid].double().mean().item() / iters, 1)
for k, v in res.items():
    print(k, v)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "umma_bench2.json"), "w"), indent=1)
