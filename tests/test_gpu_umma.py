"""Pins the tcgen05 / TMEM / swizzled-operand conventions of csrc/umma.cuh on hardware."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split", [0, 1])
def test_umma_tile_gemm(split):
    from pita_b200 import _native as N
    lib = N.load()
    gen = torch.Generator().manual_seed(split)
    A = torch.randn(128, 32, generator=gen).cuda()
    B = torch.randn(32, 32, generator=gen).cuda()
    D = torch.zeros(128, 32, device="cuda")
    N.check(lib.pita_umma_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), split, N.stream_ptr()), "pita_umma_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ B.abs().double().T).max().item()
    tol = 2e-6 if split else 2e-3
    assert err <= tol * scale, (err, scale)
    if not split:
        assert err > 1e-7 * scale  # really went through TF32
