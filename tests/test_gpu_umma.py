"""tcgen05 3xTF32 GEMM (TMA + UMMA + TMEM) against an fp64 product: fp32-level accuracy on the tensor cores."""
import ctypes as C

import pytest
import torch

from humor_b200 import _ext

pytestmark = pytest.mark.gpu


def umma(A, B, bias=None):
    L = _ext.lib()
    M, K = A.shape
    N = B.shape[0]
    ldc = (N + 3) // 4 * 4
    Cm = torch.full((M, ldc), float('nan'), device='cuda')
    ws = torch.empty(L.humor_umma_gemm_workspace_bytes(M, N, A.stride(0), B.stride(0)) // 4, device='cuda')
    rc = L.humor_umma_gemm(_ext.ptr(A), A.stride(0), _ext.ptr(B), B.stride(0), _ext.ptr(bias), _ext.ptr(Cm), ldc, M, N, K,
                           _ext.ptr(ws), ws.numel() * 4, _ext.stream_ptr())
    _ext.check(rc, 'humor_umma_gemm')
    torch.cuda.synchronize()
    return Cm[:, :N]


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (128, 128, 256), (300, 200, 96), (1024, 1024, 1024), (15104, 96, 1024), (77, 352, 1024)])
def test_umma_gemm_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ B.double().t() + bias.double()
    out = umma(A, B, bias)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err                       # 3xTF32: fp32-level accuracy (single-pass TF32 would be ~5e-4)
    fp32 = float(((A @ B.t() + bias).double() - ref).abs().max() / ref.abs().max())
    assert err < 8 * fp32 + 1e-7


def test_umma_strided_operands():
    """operands embedded in wider buffers (leading dimension > K), as the rollout tape stores them."""
    g = torch.Generator().manual_seed(1)
    Abuf = torch.randn(200, 416, generator=g).cuda()
    Bbuf = torch.randn(1024, 352, generator=g).cuda()
    out = umma(Abuf[:, :352], Bbuf)
    ref = Abuf[:, :352].double() @ Bbuf.double().t()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 1e-5
