"""Algebra of the training step (lass_b200/training.py: forward in train mode + hand-derived backward as a sequence of
kernel calls) checked on the CPU: the engine runs over the pure-torch emulation of the kernel interface
(tests/train_emul.py) and is compared with the autograd oracle (oracle/train_oracle.py, pinned to the unmodified
reference).  The CUDA kernels themselves are compared with the same emulation in tests/test_gpu_training.py."""
import numpy as np
import pytest
import torch

import helpers
import train_emul
from oracle import factory, train_oracle
from lass_b200 import training


def _setup(B=2, L=16000, exact=True):
    train_emul.set_exact(exact)
    model, sd = helpers.build_module()
    model.train()
    mix, cond = factory.make_inputs(B, L, seed=1234, edge_clips=False)
    tgt, _ = factory.make_inputs(B, L, seed=4321, edge_clips=False)
    return model, sd, mix, cond, 0.5 * tgt


def _key_of(model):
    return {id(p): n for n, p in model.named_parameters()}


def _compare(model, eng, o_grads, rel, floor, rel_l2):
    """Per parameter tensor: max|d| <= rel * max|ref| + floor * (largest gradient entry of the model), and
    ||d||_2 <= rel_l2 * ||ref||_2 + the same floor.  The max-norm bound is loose on purpose: a pre-activation within
    rounding of zero flips LeakyReLU's slope between two fp32 evaluations and moves ONE channel of a small tensor by a
    per cent or two (seen in the reference's own fp32-vs-fp64 comparison); the L2 bound is the sharp one."""
    names = _key_of(model)
    got = {names[id(p)]: g for p, g in eng.grads().items()}
    assert sorted(got) == sorted(o_grads)
    gmax = max(float(v.abs().max()) for v in o_grads.values())
    worst = 0.0
    for kname, ref in o_grads.items():
        d = float((got[kname] - ref).abs().max())
        bound = rel * float(ref.abs().max()) + floor * gmax
        worst = max(worst, d / bound)
        assert d <= bound, (kname, d, float(ref.abs().max()))
        l2 = float((got[kname] - ref).double().norm())
        assert l2 <= rel_l2 * float(ref.double().norm()) + floor * gmax * ref.numel() ** 0.5, (kname, l2, float(ref.norm()))
    return worst


def test_training_step_algebra_exact_storage():
    model, sd, mix, cond, tgt = _setup(exact=True)
    o_loss, o_wave, o_grads, o_buf = train_oracle.training_forward_backward(sd, mix, cond, tgt)
    with torch.no_grad():
        eng = training.TrainEngine(model, kernels=train_emul)
        wave = eng.forward(mix, cond)
        assert float((wave - o_wave).abs().max()) <= 2e-5 * float(o_wave.abs().max())
        loss = torch.mean(torch.abs(wave.squeeze() - tgt.squeeze()))
        assert abs(float(loss) - o_loss) <= 1e-6 * o_loss
        dwave = torch.sign(wave - tgt) / wave.numel()
        eng.backward(dwave)
    _compare(model, eng, o_grads, 3e-2, 1e-5, 5e-3)
    new_sd = model.state_dict()
    for kname, v in o_buf.items():
        assert torch.allclose(new_sd[kname].float(), v.float(), rtol=1e-4, atol=1e-6), kname


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm()) / max(float(b.double().norm()), 1e-30)


def test_training_step_16bit_storage_model():
    """fp16 forward tensors / bf16 gradient tensors as on the GPU.

    What can be asked of the gradients: this randomly initialised 26-conv network is CHAOTIC in its gradient — rounding the
    reference's own conv weights to fp16 (a 2^-12 relative perturbation that moves its output by -51 dB) moves its exact
    fp32 autograd gradients by ~22 % (median relative L2 over the parameter tensors, measured below on the oracle itself):
    every LeakyReLU unit that crosses zero switches its slope 100x, and a parameter gradient is a sum of millions of such
    gated terms of random sign.  No 16-bit implementation can agree with the fp32 gradients more closely than the reference
    agrees with itself under that perturbation, so the bound is stated relative to it; the algebra is pinned exactly by the
    fp32-storage test above and the GPU tests add a self-consistent directional-derivative check."""
    model, sd, mix, cond, tgt = _setup(exact=False)
    o_loss, o_wave, o_grads, _ = train_oracle.training_forward_backward(sd, mix, cond, tgt)
    sd_r = {k: (v.to(torch.float16).float() if (k.endswith("weight") and "conv" in k and "stft" not in k) else v)
            for k, v in sd.items()}
    _, r_wave, r_grads, _ = train_oracle.training_forward_backward(sd_r, mix, cond, tgt)
    own = float(np.median([_rel_l2(r_grads[k], o_grads[k]) for k in o_grads]))
    assert float(factory.snr_db(o_wave, r_wave).min()) >= 45.0 and own >= 0.05      # the premise of the docstring
    with torch.no_grad():
        eng = training.TrainEngine(model, kernels=train_emul)
        wave = eng.forward(mix, cond)
        snr = factory.snr_db(o_wave, wave)
        assert float(snr.min()) >= 40.0, snr
        eng.backward(torch.sign(wave - tgt) / wave.numel())
    names = _key_of(model)
    got = {names[id(p)]: g for p, g in eng.grads().items()}
    ours = float(np.median([_rel_l2(got[k], o_grads[k]) for k in o_grads]))
    assert ours <= 2.0 * own, (ours, own)
    cos = [float((got[k].double() * o_grads[k].double()).sum() / (got[k].double().norm() * o_grads[k].double().norm() + 1e-300))
           for k in o_grads if float(o_grads[k].abs().max()) > 1e-7]
    assert float(np.median(cos)) >= 0.9, float(np.median(cos))
    train_emul.set_exact(False)


def test_autograd_bridge_and_fused_step():
    model, sd, mix, cond, tgt = _setup(B=2, L=8000, exact=True)
    eng = training.TrainEngine(model, kernels=train_emul)
    wave = training.train_forward(eng, mix, cond)
    loss = torch.mean(torch.abs(wave.squeeze() - tgt[:, :, :8000].squeeze()))
    loss.backward()
    g_bridge = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert all(not train_oracle.is_dead_key(n) for n in g_bridge)
    # the fused step computes the same gradients and applies AdamW-amsgrad to the flat buffer
    model2, _, _, _, _ = _setup(B=2, L=8000, exact=True)
    eng2 = training.TrainEngine(model2, kernels=train_emul)
    before = {n: p.detach().clone() for n, p in model2.named_parameters()}
    with torch.no_grad():
        loss2 = eng2.training_step(mix, cond, tgt[:, :, :8000], lr=1e-3)
    assert abs(float(loss2) - float(loss.detach())) <= 1e-6 * float(loss.detach())
    opt_p = {n: torch.nn.Parameter(v.clone()) for n, v in before.items() if n in g_bridge}
    opt = torch.optim.AdamW(list(opt_p.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    for n, p in opt_p.items():
        p.grad = g_bridge[n]
    opt.step()
    after = dict(model2.named_parameters())
    for n, p in opt_p.items():
        assert torch.allclose(after[n].detach(), p.detach(), rtol=1e-5, atol=1e-7), n
    for n in before:
        if n not in g_bridge and after[n].requires_grad:
            assert torch.equal(after[n].detach(), before[n]), n      # dead parameters are untouched
