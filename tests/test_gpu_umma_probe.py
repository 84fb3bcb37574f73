"""Pins the tcgen05 shared-memory descriptor rules the conv kernel (lass_b200/csrc/conv.cu) relies on, on real
hardware: swizzle is applied on absolute shared-memory address bits, so descriptor start addresses shifted by whole
rows and 8-row-group strides that are not multiples of the swizzle atom are legal (base_offset = 0)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kc,swz", [(64, 2), (32, 4)])
def test_descriptor_rules(dtype, kc, swz):
    from lass_b200 import ops
    torch.manual_seed(0)
    rowb, atom = kc * 2, 8 * kc * 2
    A = torch.randn(256, kc, device="cuda").to(dtype)
    idx = torch.arange(128, device="cuda")
    for n in (32, 64, 128, 256):
        Bm = torch.randn(n, kc, device="cuda").to(dtype)
        full = A.float() @ Bm.float().t()
        tol = 1e-4 * float(full.abs().max())
        assert float((ops.umma_probe(A, Bm, swz, 0, atom, 0, atom) - full[:128]).abs().max()) <= tol
    Bm = torch.randn(64, kc, device="cuda").to(dtype)
    full = A.float() @ Bm.float().t()
    tol = 1e-4 * float(full.abs().max())
    for dx in (1, 2, 7):        # row-shifted start
        out = ops.umma_probe(A, Bm, swz, rowb * dx, atom, 0, atom)
        assert float((out - full[dx:dx + 128]).abs().max()) <= tol
    rows10 = (idx // 8) * 10 + idx % 8     # dense halo pitch (10 pixels per image row), tap offsets dy*10 + dx
    for shift in (0, 1, 2, 11, 22):
        out = ops.umma_probe(A, Bm, swz, rowb * shift, 10 * rowb, 0, atom)
        assert float((out - full[rows10 + shift]).abs().max()) <= tol
