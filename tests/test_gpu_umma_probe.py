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


# ---- MN-major operands (rows = contraction index): the layouts of the weight-gradient kernel (csrc/wgrad_tc.cu) ----
def _probe_mn(A, Bm, a_swz, b_swz, n, ksteps, a_start, a_lbo, a_sbo, a_kstep, b_start, b_lbo, b_sbo, b_kstep):
    import ctypes
    from lass_b200 import _cabi
    out = torch.zeros(128, n, device="cuda")
    lib = _cabi.load_debug()
    _cabi.check(lib.lass_debug_umma_probe_mn(
        A.data_ptr(), A.shape[0], a_swz, Bm.data_ptr(), Bm.shape[0], b_swz, n, ksteps, a_start, a_lbo, a_sbo, a_kstep,
        b_start, b_lbo, b_sbo, b_kstep, 1 if A.dtype == torch.float16 else 0, 1 if Bm.dtype == torch.float16 else 0,
        out.data_ptr(), torch.cuda.current_stream().cuda_stream), lib)
    torch.cuda.synchronize()
    return out.cpu()


def _expect_mn(A, Bm, n, ksteps, a_start, a_lbo, a_sbo, a_kstep, b_start, b_lbo, b_sbo, b_kstep):
    """D[m, j] = sum_k A_op[k, m] * B_op[k, j] with the canonical MN-major addressing in units of tile rows."""
    wa, wb = A.shape[1], Bm.shape[1]
    ra, rb = wa * 2, wb * 2
    Af, Bf = A.float().cpu(), Bm.float().cpu()
    D = torch.zeros(128, n, dtype=torch.float64)
    for ks in range(ksteps):
        for k in range(16):
            arow = (a_start + ks * a_kstep) // ra + (k // 8) * (a_sbo // ra) + k % 8
            brow = (b_start + ks * b_kstep) // rb + (k // 8) * (b_sbo // rb) + k % 8
            a = torch.stack([Af[arow + (m // wa) * (a_lbo // ra), m % wa] for m in range(128)]).double()
            b = torch.stack([Bf[brow + (j // wb) * (b_lbo // rb), j % wb] for j in range(n)]).double()
            D += a[:, None] * b[None, :]
    return D.float()


# Both operands share ONE 16-bit format: a descriptor with a_format != b_format (fp16 x bf16) raised cudaErrorIllegalInstruction
# on B200 and poisons the context, so that case cannot be kept as a test; csrc/wgrad_tc.cu converts fp16 tiles to bf16 instead.
MN_CASES = [
    # name, dtype A, dtype B, a_swz, b_swz, a_rows, b_rows, n, ksteps, a_start, a_lbo, a_sbo, a_kstep, b_start, b_lbo, b_sbo, b_kstep
    ("sw128 overlapped atoms (2 dx taps x 64 ch), shifted start", torch.bfloat16, torch.bfloat16, 2, 2, 64, 32, 64, 2,
     3 * 128, 128, 10 * 128, 20 * 128, 0, 0, 8 * 128, 16 * 128),
    ("sw64 four overlapped atoms (4 dx taps x 32 ch)", torch.bfloat16, torch.bfloat16, 4, 4, 64, 32, 32, 2, 5 * 64, 64,
     10 * 64, 20 * 64, 0, 0, 8 * 64, 16 * 64),
    ("sw64 fp16, N = 64 over two separate atoms", torch.float16, torch.float16, 4, 4, 64, 96, 64, 1, 11 * 64, 64, 10 * 64, 0,
     0, 32 * 64, 8 * 64, 0),
    ("sw128 atoms in separate regions, N = 128", torch.bfloat16, torch.bfloat16, 2, 2, 128, 128, 128, 1, 0, 64 * 128, 8 * 128,
     0, 0, 64 * 128, 8 * 128, 0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", MN_CASES, ids=[c[0] for c in MN_CASES])
def test_mn_major_descriptors(case):
    name, dta, dtb, a_swz, b_swz, a_rows, b_rows, n, ksteps = case[:9]
    args = case[9:]
    g = torch.Generator().manual_seed(5)
    A = torch.randint(-4, 5, (a_rows, 64 if a_swz == 2 else 32), generator=g).to(dta).cuda()
    Bm = torch.randint(-4, 5, (b_rows, 64 if b_swz == 2 else 32), generator=g).to(dtb).cuda()
    got = _probe_mn(A, Bm, a_swz, b_swz, n, ksteps, *args)
    want = _expect_mn(A, Bm, n, ksteps, *args)
    assert torch.equal(got, want), "%s: max |d| = %g" % (name, float((got - want).abs().max()))
