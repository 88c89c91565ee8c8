#!/usr/bin/env python
"""Hardware probe for the operand-ingest hypothesis: the 3xTF32 GEMM (8 bytes per operand element) against
the fp16 hi/lo GEMM (4 bytes per element, csrc/umma_gemm16.cuh) at the shapes of the step - a decoder-chain layer (M = 256,
split-K clusters), the batched prior (M = 15 104, 128-wide tiles) - accuracy against fp64 and in-situ kernel time from
torch.profiler (CUPTI; the conversion kernels of the utility entry points are reported separately, not mixed in).
One JSON line per shape.   python tools/umma16_probe.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import _ext  # noqa: E402

L = _ext.lib()
p = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(kind, A, B, bias, Cm, M, N, K):
    if kind == 'tf32x3':
        ws = torch.empty(L.humor_umma_gemm_workspace_bytes(M, N, K, K) // 4, device='cuda')
        return lambda: _ext.check(L.humor_umma_gemm(p(A), K, p(B), K, p(bias), p(Cm), Cm.stride(0), M, N, K, p(ws), ws.numel() * 4, st()), kind)
    ws = torch.empty(L.humor_umma_gemm16_workspace_bytes(M, N, K, K) // 4 + 1, device='cuda')
    return lambda: _ext.check(L.humor_umma_gemm16(p(A), K, p(B), K, p(bias), p(Cm), Cm.stride(0), M, N, K, p(ws), ws.numel() * 4, st()), kind)


def kernel_us(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            a = agg.setdefault(e.name.split('(')[0][:60], [0.0, 0])
            a[0] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
            a[1] += 1
    return {k: round(v[0] / v[1], 3) for k, v in agg.items()}


for name, (M, N, K) in [('decoder layer 2 (split-K clusters)', (256, 1024, 1088)), ('decoder layer 1', (256, 1024, 448)),
                        ('batched prior layer', (15104, 1024, 1024)), ('LBS blend slab', (512, 20672, 256))]:
    g = torch.Generator(device='cpu').manual_seed(M + N)
    A = (torch.randn(M, K, generator=g) * 0.7).cuda()
    B = (torch.randn(N, K, generator=g) * 0.03).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = (A.double() @ B.double().T + bias.double())
    scale = float(ref.abs().max())
    rec = {'shape': [M, N, K], 'what': name}
    for kind in ('tf32x3', 'fp16x2'):
        Cm = torch.full((M, N), float('nan'), device='cuda')
        fn = run(kind, A, B, bias, Cm, M, N, K)
        fn()
        torch.cuda.synchronize()
        rec[kind] = {'rel_err_vs_fp64': float((Cm.double() - ref).abs().max()) / scale, 'kernel_us': kernel_us(fn)}
    print(json.dumps(rec), flush=True)
