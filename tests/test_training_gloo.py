"""The training step's only collective on CPU: world_size-2 gloo run of TrainEngine.training_step over the pure-torch kernel
emulation (tests/train_emul.py).  Each rank trains on its own clip (DDP, reference train.py:266-283); afterwards both ranks
must hold IDENTICAL parameters, the flat gradient must be the SUM of the per-rank gradients (dead parameters excluded on
every rank alike), and the optimizer must have applied their mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
import train_emul
from oracle import factory
from lass_b200 import training

L = 6000


def _data():
    mix, cond = factory.make_inputs(2, L, seed=1234, edge_clips=False)
    tgt, _ = factory.make_inputs(2, L, seed=4321, edge_clips=False)
    return mix, cond, 0.5 * tgt


def _engine():
    train_emul.set_exact(True)
    model, _ = helpers.build_module()
    model.train()
    return model, training.TrainEngine(model, kernels=train_emul)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mix, cond, tgt = _data()
    model, eng = _engine()
    with torch.no_grad():
        loss = eng.training_step(mix[rank:rank + 1], cond[rank:rank + 1], tgt[rank:rank + 1], lr=1e-3)
    ret["P%d" % rank] = eng.P.clone()
    ret["G%d" % rank] = eng.G.clone()
    ret["loss%d" % rank] = float(loss)
    ret["live_end"] = int(eng.live_end)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_step_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    n = ret["live_end"]
    assert torch.equal(ret["P0"], ret["P1"]), "ranks diverged after the step"
    assert torch.equal(ret["G0"][:n], ret["G1"][:n])
    assert float(ret["G0"][n:].abs().max()) == 0.0 if ret["G0"].numel() > n else True      # dead parameters: no gradient
    # single-process reference of the collective: per-clip gradients summed, AdamW on their mean
    mix, cond, tgt = _data()
    g_sum, p0 = None, None
    for r in range(2):
        model, eng = _engine()
        p0 = eng.P.clone()
        with torch.no_grad():
            wave = eng.forward(mix[r:r + 1], cond[r:r + 1])
            loss = float(torch.mean(torch.abs(wave - tgt[r:r + 1])))
            assert abs(loss - ret["loss%d" % r]) <= 1e-6 * loss
            eng.backward(torch.sign(wave - tgt[r:r + 1]) / wave.numel())
        g_sum = eng.G.clone() if g_sum is None else g_sum + eng.G
    scale = float(g_sum[:n].abs().max())
    # (the workers run 2 threads, this process all of them: fp32 summation order differs and the network amplifies it)
    assert float((ret["G0"][:n] - g_sum[:n]).abs().max()) <= 3e-3 * scale
    p = torch.nn.Parameter(p0[:n].clone())
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    p.grad = 0.5 * ret["G0"][:n]                      # the optimizer must see the MEAN of the all-reduced sum (1 / world)
    opt.step()
    assert torch.allclose(ret["P0"][:n], p.detach(), rtol=1e-5, atol=1e-7)
    assert torch.equal(ret["P0"][n:], p0[n:])


# ------------------------------------------------------------------------------------------------- SyncBatchNorm
def _sync_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mix, cond, tgt = _data()
    train_emul.set_exact(True)
    model, _ = helpers.build_module()
    model.train()
    eng = training.TrainEngine(model, kernels=train_emul, sync_batchnorm=True)
    with torch.no_grad():
        loss = eng.training_step(mix[rank:rank + 1], cond[rank:rank + 1], tgt[rank:rank + 1], lr=1e-3)
    ret["P%d" % rank] = eng.P.clone()
    ret["G%d" % rank] = eng.G.clone()
    ret["loss%d" % rank] = float(loss)
    ret["rm%d" % rank] = model.base.encoder_block3.conv_block1.bn2.running_mean.clone()
    ret["rv%d" % rank] = model.base.encoder_block3.conv_block1.bn2.running_var.clone()
    ret["rv0_%d" % rank] = model.base.bn0.running_var.clone()
    ret["live_end"] = int(eng.live_end)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sync_batchnorm_equals_one_rank_with_the_whole_batch():
    """``sync_batchnorm: True`` (reference config/audiosep_base.yaml:42 -> torch.nn.SyncBatchNorm under DDP): two ranks with one
    clip each must reproduce ONE rank training on both clips -- same running statistics, gradient sum = 2 x the whole-batch
    gradient (each rank's loss is the mean over ITS clip), same parameters after the step."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_sync_worker, args=(2, port, ret), nprocs=2, join=True)
    n = ret["live_end"]
    assert torch.equal(ret["P0"], ret["P1"]) and torch.equal(ret["G0"][:n], ret["G1"][:n])
    for key in ("rm", "rv", "rv0_"):
        assert torch.equal(ret[key + "0"], ret[key + "1"])           # every rank tracks the GLOBAL statistics
    mix, cond, tgt = _data()
    model, eng = _engine()                                            # statistics per rank = over everything on one rank
    with torch.no_grad():
        loss = eng.training_step(mix, cond, tgt, lr=1e-3)
    assert abs(float(loss) - 0.5 * (ret["loss0"] + ret["loss1"])) <= 1e-5 * float(loss)
    bn = model.base.encoder_block3.conv_block1.bn2
    assert torch.allclose(ret["rm0"], bn.running_mean, rtol=1e-4, atol=1e-6)
    assert torch.allclose(ret["rv0"], bn.running_var, rtol=1e-4, atol=1e-7)          # unbiased with the GLOBAL count
    assert torch.allclose(ret["rv0_0"], model.base.bn0.running_var, rtol=1e-4, atol=1e-7)
    scale = float(eng.G[:n].abs().max())
    assert float((0.5 * ret["G0"][:n] - eng.G[:n]).abs().max()) <= 3e-3 * scale
    cos = torch.nn.functional.cosine_similarity(0.5 * ret["G0"][:n], eng.G[:n], dim=0)
    assert float(cos) > 0.9999
