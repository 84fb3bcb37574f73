"""Step-size sweep of the directional-derivative check of tests/test_gpu_training.py (own-gradient and fp32-autograd
directions, per parameter group) -- how the test's step and band were chosen."""
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np
import helpers
from oracle import factory, train_oracle
from lass_b200 import training
B, L = 2, 16000
model, sd = helpers.build_module()
mix, cond = factory.make_inputs(B, L, seed=1234, edge_clips=False)
tgt, _ = factory.make_inputs(B, L, seed=4321, edge_clips=False)
tgt = 0.5 * tgt
o_loss, o_wave, o_grads, o_buf = train_oracle.training_forward_backward(sd, mix, cond, tgt)
model = model.cuda().train()
mix, cond, tgt = mix.cuda(), cond.cuda(), tgt.cuda()
eng = training.TrainEngine(model)
names = {id(p): n for n, p in model.named_parameters()}
def grad_once():
    with torch.no_grad():
        w = eng.forward(mix, cond)
        eng.backward(torch.sign(w - tgt) / w.numel())
    return eng.G.clone()
G = grad_once(); G2 = grad_once()
print("determinism: rel diff of two gradient evaluations %.3e" % float((G - G2).norm() / G.norm()))
O = torch.zeros_like(G)
for name, (off, p) in eng.index.items():
    if name.startswith("dead."): continue
    O[off:off + p.numel()] = o_grads[names[id(p)]].reshape(-1).cuda()
P0 = eng.P.clone()
def loss_at(P):
    with torch.no_grad():
        eng.P.copy_(P); eng.refresh_weights()
        return float(torch.mean(torch.abs(eng.forward(mix, cond) - tgt)).double())
base = loss_at(P0)
groups = {"all": (0, eng.live_end), "dec": (0, eng.bucket_a_end), "enc": (eng.bucket_a_end, eng.film_w_off), "film": (eng.film_w_off, eng.live_end)}
for gname, (lo, hi) in groups.items():
    for dname, D in (("own", G), ("oracle", O)):
        d = torch.zeros_like(G); d[lo:hi] = D[lo:hi]; nrm = float(d.double().norm()); d = d / nrm
        pred = float((G.double() * d.double()).sum())
        out = []
        for rel in (5e-4, 1e-3, 2e-3, 4e-3):
            eta = rel * base / nrm
            fd = (loss_at(P0 + eta * d) - loss_at(P0 - eta * d)) / (2 * eta)
            out.append("%.4f" % (fd / pred))
        cosv = float((G[lo:hi].double() * O[lo:hi].double()).sum() / (G[lo:hi].double().norm() * O[lo:hi].double().norm()))
        print("%s dir=%s: pred %.4e  fd/pred at rel 5e-4..4e-3: %s   cos(G,O)=%.4f |G|/|O|=%.3f" % (gname, dname, pred, " ".join(out), cosv, float(G[lo:hi].norm() / O[lo:hi].norm())), flush=True)
