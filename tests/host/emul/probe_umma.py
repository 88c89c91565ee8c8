"""Child process: humor_umma_gemm (3xTF32 tcgen05 GEMM) through the library's own dispatch on the emulation - shapes that take
the single-CTA tiles and shapes that take the split-K path over a 4-CTA cluster (DSMEM reduction), against fp64.  Prints JSON."""
import ctypes as C
import json
import sys

import numpy as np
import torch

root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

L = cpu_backend.install(lib)
out = []
for M, N, K in [(130, 200, 288), (40, 70, 64)]:
    rng = np.random.RandomState(M + N + K)
    A = torch.tensor(rng.randn(M, K).astype(np.float32))
    B = torch.tensor((rng.randn(N, K) * 0.1).astype(np.float32))
    bias = torch.tensor(rng.randn(N).astype(np.float32))
    ldc = ((N + 3) // 4) * 4
    Cm = torch.full((M, ldc), float('nan'))
    ws = torch.empty(L.humor_umma_gemm_workspace_bytes(M, N, K, K) // 4)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = L.humor_umma_gemm(p(A), K, p(B), K, p(bias), p(Cm), ldc, M, N, K, p(ws), ws.numel() * 4, None)
    ref = A.double() @ B.double().T + bias.double()
    err = float((Cm[:, :N].double() - ref).abs().max() / ref.abs().max())
    out.append({'shape': [M, N, K], 'rc': rc, 'rel_err': err, 'finite': bool(torch.isfinite(Cm[:, :N]).all())})
# the same product from 4-byte operand elements (fp16 hi + scaled fp16 lo planes, umma_gemm16.cuh): K % 64 == 0
out16 = []
for M, N, K in [(130, 200, 576), (40, 70, 64), (1030, 70, 64)]:
    rng = np.random.RandomState(M + N + K + 1)
    A = torch.tensor((rng.randn(M, K) * np.exp(rng.randn(M, K))).astype(np.float32))      # several decades of magnitude
    B = torch.tensor((rng.randn(N, K) * 0.05).astype(np.float32))
    B[0, :8] = torch.tensor([1e-6, -3e-7, 5e-5, 6.2e-5, 1e-9, 0.0, -2.5e-4, 7e-8])         # around and below fp16's normal range
    bias = torch.tensor(rng.randn(N).astype(np.float32))
    ldc = ((N + 3) // 4) * 4
    Cm = torch.full((M, ldc), float('nan'))
    ws = torch.empty(L.humor_umma_gemm16_workspace_bytes(M, N, K, K) // 4 + 1)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = L.humor_umma_gemm16(p(A), K, p(B), K, p(bias), p(Cm), ldc, M, N, K, p(ws), ws.numel() * 4, None)
    ref = A.double() @ B.double().T + bias.double()
    scale = (A.double().abs() @ B.double().abs().T + bias.double().abs())                   # sum |a||b|: the natural error scale
    err = float(((Cm[:, :N].double() - ref).abs() / scale).max())
    out16.append({'shape': [M, N, K], 'rc': rc, 'rel_err': err, 'finite': bool(torch.isfinite(Cm[:, :N]).all()),
                  'row0_err': float((Cm[:, 0].double() - ref[:, 0]).abs().max())})
print(json.dumps({'cases': out, 'cases16': out16}))
