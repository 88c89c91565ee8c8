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
for M, N, K in [(130, 200, 288), (256, 216, 576), (40, 70, 64), (300, 130, 96)]:
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
print(json.dumps({'cases': out}))
