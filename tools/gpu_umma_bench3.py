"""Issue rate of the conv kernel's steady-state MMA code in isolation (B200): cycles per tcgen05.mma."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lass_b200 import _cabi
lib = _cabi.load_debug()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
res = {}
iters = 512
for (mt, bn, ks) in ((2, 32, 2), (1, 64, 4), (1, 256, 4)):
    for mode in (0, 2, 4, 8, 14):
        _cabi.check(lib.lass_debug_umma_bench3(mt, bn, ks, mode, iters, 148, out.data_ptr(), None))
        torch.cuda.synchronize()
        res["mt%d_n%d_ks%d_mode%d" % (mt, bn, ks, mode)] = round(out.double().mean().item() / (iters * 9 * mt * ks), 1)
for k, v in res.items():
    print(k, v)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "umma_bench3.json"), "w"), indent=1)
