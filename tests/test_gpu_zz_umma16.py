"""GPU parity of the fp16 hi/lo tcgen05 GEMM (csrc/umma_gemm16.cuh, utility entry point humor_umma_gemm16) against fp64: single-CTA
tiles, split-K clusters, 128-wide tiles, ragged M / N, operands spread over decades.  The same cases pass on the CPU emulation
(tests/test_emul_product.py::test_umma_gemm_single_cta_and_split_k_cluster)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from humor_b200 import _ext

# First hardware run: round 2, call r02a (profiles/r02a_gpu_tests_ungated.txt): green on the B200.
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,N,K', [(130, 200, 576), (256, 1024, 1088), (40, 70, 64), (1100, 130, 128), (2048, 1024, 1024)])
def test_umma_gemm16_matches_fp64(M, N, K):
    L = _ext.lib()
    rng = np.random.RandomState(M + N + K)
    A = torch.tensor((rng.randn(M, K) * np.exp(rng.randn(M, K))).astype(np.float32)).cuda()
    B = torch.tensor((rng.randn(N, K) * 0.05).astype(np.float32)).cuda()
    bias = torch.tensor(rng.randn(N).astype(np.float32)).cuda()
    ldc = ((N + 3) // 4) * 4
    Cm = torch.full((M, ldc), float('nan'), device='cuda')
    ws = torch.empty(L.humor_umma_gemm16_workspace_bytes(M, N, K, K) // 4 + 1, device='cuda')
    p = lambda t: C.c_void_p(t.data_ptr())
    _ext.check(L.humor_umma_gemm16(p(A), K, p(B), K, p(bias), p(Cm), ldc, M, N, K, p(ws), ws.numel() * 4, _ext.stream_ptr()), 'humor_umma_gemm16')
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T + bias.double()
    scale = A.double().abs() @ B.double().abs().T + bias.double().abs()
    assert torch.isfinite(Cm[:, :N]).all()
    assert float(((Cm[:, :N].double() - ref).abs() / scale).max()) < 2e-6          # emulation: 4e-7; the tensor core truncates its accumulator
    assert torch.isnan(Cm[:, N:]).all()


@pytest.mark.parametrize('name', ['stage3_rgb', 'stage3_rgb_phase1', 'stage3_amass'])
def test_closure_tensor16_matches_reference_golden(name):
    """precision 'tensor16': the forward decoder chain on fp16 hi/lo operand planes (umma_gemm16 + chain16_pack_kernel), everything
    else as 'tensor' - against the fixtures of the unmodified reference, with the tolerances of the 'tensor' mode."""
    from humor_b200 import synth  # noqa: F401
    from tests import util_stage3 as U
    from tests.golden_util import load_case, check_against_golden
    g, prob, c = load_case(name)
    mo = U.build_product(c['B'], c['T'], c['W'], c['optim_floor'], prob)
    mo.set_precision('tensor16')
    loss, grads, aux = U.closure_product(mo, prob, c['nsteps'], c['scale'])
    check_against_golden(g, loss, aux['stats'], grads, loss_tol=1e-5, stat_tol=1e-4, grad_tol=1e-2)
    assert np.abs(aux['cam_pred']['verts3d'].detach().cpu().numpy() - g['cam_verts3d']).max() < 1e-4
    assert np.abs(aux['roll']['trans'].detach().cpu().numpy() - g['rollout_trans']).max() < 2e-5
    pm = aux['roll']['cond_prior'][0].detach().cpu().numpy()
    assert np.abs(pm - g['cond_prior_mean']).max() / np.abs(g['cond_prior_mean']).max() < 1e-5
    # and the mode really changes the kernels: same closure in 'tensor' differs in the last bits
    mo.set_precision('tensor')
    loss_t, _, _ = U.closure_product(mo, prob, c['nsteps'], c['scale'])
    assert loss_t != loss and abs(loss_t - loss) <= 2e-6 * abs(loss)
