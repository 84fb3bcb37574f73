"""How much the training gradient of this randomly initialised network moves when ONLY the batch size changes: B clips against
the same clips duplicated (2B) on one GPU.  BatchNorm statistics of a duplicated batch are exactly those of the batch, and the
mean loss too, so in exact arithmetic the two gradients are identical; on the GPU the launches pick other tile / split-K
configurations (another fp32 summation order), 16-bit roundings flip, and the train-mode network amplifies it.  This is the noise
floor against which tools/gpu_syncbn_check.py's split-batch comparison has to be read."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from lass_b200 import training  # noqa: E402
from oracle import factory  # noqa: E402

dev = torch.device("cuda", 0)
B, L = 2, 32000
mix, cond = factory.make_inputs(B, L, seed=1234, edge_clips=False)
tgt, _ = factory.make_inputs(B, L, seed=4321, edge_clips=False)
tgt = 0.5 * tgt
out = {}
runs = {}
for name, rep in (("B2", 1), ("B2_again", 1), ("B4_duplicated", 2)):
    model, _ = helpers.build_module(device=dev)
    model.train()
    eng = training.TrainEngine(model)
    with torch.no_grad():
        eng.training_step(mix.repeat(rep, 1, 1).to(dev), cond.repeat(rep, 1).to(dev), tgt.repeat(rep, 1, 1).to(dev), lr=1e-3)
    torch.cuda.synchronize()
    runs[name] = (eng.G[:eng.live_end].clone(), eng._last.wave[:B].clone(), eng)
names = ["after.w", "dec5.cb2.conv2.weight", "dec5.cb2.conv1.weight", "dec5.up", "dec4.cb2.conv1.weight", "dec2.cb2.conv1.weight", "dec0.up"]
eng = runs["B2"][2]
for other in ("B2_again", "B4_duplicated"):
    a, b = runs["B2"][0], runs[other][0]
    groups = {}
    for nm in names:
        off, prm = eng.index[nm]
        groups[nm] = round(float(torch.nn.functional.cosine_similarity(a[off:off + prm.numel()], b[off:off + prm.numel()], dim=0)), 5)
    out[other] = {"grad_cosine": float(torch.nn.functional.cosine_similarity(a, b, dim=0)), "grad_norm_ratio": float(b.norm() / a.norm()),
                  "wave_snr_db_per_clip": [round(float(v), 2) for v in factory.snr_db(runs["B2"][1].cpu(), runs[other][1].cpu())],
                  "groups": groups}
print(json.dumps(out))
