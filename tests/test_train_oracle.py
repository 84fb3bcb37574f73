"""The training-step oracle (oracle/train_oracle.py) against the UNMODIFIED reference in ``.train()`` mode
(reference models/audiosep.py:99-111 + losses.py:4-9), and against torch.optim.AdamW(amsgrad=True)."""
import os

import numpy as np
import pytest
import torch

from oracle import factory, reference_loader, train_oracle
import helpers


def _inputs(B=2, L=16000):
    mix, cond = factory.make_inputs(B, L, seed=1234, edge_clips=False)
    tgt, _ = factory.make_inputs(B, L, seed=4321, edge_clips=False)
    return mix, cond, 0.5 * tgt


@pytest.mark.skipif(not reference_loader.reference_available(), reason="/root/reference only exists in the build container")
def test_train_oracle_matches_unmodified_reference():
    ref_mod = reference_loader.import_reference_resunet()
    torch.manual_seed(0)
    model = ref_mod.ResUNet30(input_channels=1, output_channels=1, condition_size=512)
    sd = factory.fill_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    model.train()                                                       # reference models/audiosep.py:99
    mix, cond, tgt = _inputs()
    out = model({"mixture": mix, "condition": cond})["waveform"]        # :100
    loss = torch.mean(torch.abs(out.squeeze() - tgt.squeeze()))         # losses.py:4-9
    loss.backward()
    o_loss, o_wave, o_grads, o_buf = train_oracle.training_forward_backward(sd, mix, cond, tgt)
    assert abs(o_loss - float(loss.detach())) <= 1e-7
    assert float((o_wave - out.detach()).abs().max()) <= 1e-6
    n_live = 0
    # fp32 autograd of this network is only reproducible to ~1e-3 of a tensor's max (thread count alone moves it by that
    # much; fp32 vs fp64: 2e-3), and some tensors' gradients are pure rounding noise (shortcut biases ahead of a batch-stat
    # BatchNorm: ~1e-10 against 1e-3 elsewhere) -- hence a per-tensor relative bound plus a global absolute floor
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.grad is None:
            assert train_oracle.is_dead_key(name), name
            assert name not in o_grads
            continue
        assert not train_oracle.is_dead_key(name), name
        n_live += 1
        d = float((o_grads[name] - p.grad).abs().max())
        assert d <= 1e-2 * float(p.grad.abs().max()) + 1e-6 * gmax, (name, d, float(p.grad.abs().max()))
    assert n_live == len(o_grads)
    new_sd = model.state_dict()
    for k, v in o_buf.items():
        assert torch.allclose(v.float(), new_sd[k].float(), rtol=1e-6, atol=1e-7), k


def test_train_oracle_matches_golden_summary():
    g = helpers.golden("train_step_b2_l16000.npz")
    from lass_b200.models.resunet import ResUNet30
    torch.manual_seed(0)
    sd = factory.fill_state_dict(ResUNet30(1, 1, 512).state_dict(), seed=0)
    mix, cond, tgt = _inputs()
    loss, wave, grads, buf = train_oracle.training_forward_backward(sd, mix, cond, tgt)
    assert abs(loss - float(g["loss"])) <= 1e-6 * float(g["loss"])
    keys = [str(k) for k in g["keys"]]
    assert sorted(keys) == sorted(grads)
    norms = np.array([float(grads[k].double().norm()) for k in keys])
    # fp32 autograd of this network reproduces to ~1e-3 of a tensor's max only (see the test above)
    gmax = float(g["grad_absmax"].max())
    assert np.all(np.abs(norms - g["grad_norms"]) <= 5e-3 * g["grad_norms"] + 1e-5 * gmax)
    first = np.array([float(grads[k].reshape(-1)[0]) for k in keys])
    assert np.all(np.abs(first - g["grad_first"]) <= 1e-2 * g["grad_absmax"] + 1e-6 * gmax)
    assert np.allclose(buf["base.bn0.running_mean"].numpy(), g["bn0_running_mean"], rtol=1e-5, atol=1e-7)
    assert np.allclose(buf["base.encoder_block3.conv_block1.bn2.running_var"].numpy(), g["enc3_bn2_running_var"], rtol=1e-5, atol=1e-7)


def test_adamw_amsgrad_restatement_matches_torch():
    torch.manual_seed(3)
    p0 = torch.randn(1000)
    p_t = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_t], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True, foreach=False)
    p, m, v, vmax = p0.clone(), torch.zeros(1000), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 6):
        g = torch.randn(1000) * (0.1 if step != 3 else 10.0)
        p_t.grad = g.clone()
        opt.step()
        train_oracle.adamw_amsgrad_step(p, g, m, v, vmax, step, 1e-3)
        assert torch.equal(p, p_t.detach()), step
