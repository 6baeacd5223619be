"""GPU: the tcgen05 3xTF32 GEMM building block against a torch fp64 product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 224, 32), (128, 224, 64), (256, 448, 96), (100, 224, 40), (1024, 224 * 3, 784),
                                   (130, 100, 50), (784, 500, 1024), (1, 1, 1), (129, 225, 33)])
def test_gemm_nt_3xtf32(M, N, K):
    from brancher_b200 import _cuda as cu
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = cu.gemm_nt_3xtf32(A, B)
    torch.cuda.synchronize()
    assert cu.last_variant() == "tcgen05"
    ref = A.double() @ B.double().T
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ B.abs().double().T).max().item()
    # fp32-equivalent: error relative to sum |a||b| of the order 2^-22 (single TF32 would be ~2^-11)
    assert err <= 2e-6 * scale, (err, scale, err / scale)
    fp32 = (A @ B.T).double()
    print("M,N,K=%s: err/scale=%.2e (torch fp32: %.2e)" % ((M, N, K), err / scale, (fp32 - ref).abs().max().item() / scale))


@pytest.mark.parametrize("M,N,K", [(128, 128, 8), (128, 128, 16), (256, 256, 64), (200, 136, 300), (784, 512, 4096), (33, 40, 7)])
def test_gemm_mn_major_tf32(M, N, K, monkeypatch):
    """Both operands read MN-major (D = At^T Bt for row-major At [K][M], Bt [K][N]: the K5 weight-gradient GEMM): tf32 through the
    128-byte swizzle with 32-byte atoms (TMA SWIZZLE_128B_ATOM_32B, UMMA layout SWIZZLE_128B_BASE32B), operands split into TF32
    pairs inside the stage.  Ragged tiles in M, N and K; same accuracy bound as the K-major kernel."""
    from brancher_b200 import _cuda as cu
    monkeypatch.setenv("BRN_GEMM_BN", "-300")
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = cu.gemm_nt_3xtf32(A, B)          # the entry transposes A and B into [K][M] / [K][N] and runs the MN-major kernel on those
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ B.abs().double().T).max().item()
    assert err <= 2e-6 * scale, (err, scale, err / scale)
