"""lbs_skin_group_kernel (lane = frame, vertex groups; humor_b200/csrc/lbs_skin_group.cuh) executed on the CPU through
the SIMT shim against a dense numpy evaluation of the skinning sum  out = sum_j W[v,j] (A[n,j] [p;1]) + trans."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from humor_b200 import synth
from humor_b200.body_model import pack_smplh

HERE = os.path.dirname(os.path.abspath(__file__))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def H():
    so = os.path.join(HERE, 'host', 'lbs_skin_host.so')
    src = os.path.join(HERE, 'host', 'lbs_skin_host.cpp')
    subprocess.check_call(['g++', '-O1', '-std=c++20', '-pthread', '-shared', '-fPIC', '-I' + os.path.join(HERE, 'host', 'shim'),
                           '-DHB_HOST_SHIM', src, '-o', so])
    return ctypes.CDLL(so)


@pytest.fixture(scope='module')
def packed():
    asset = synth.make_smplh_asset()
    return asset, pack_smplh(asset, 16)


def test_group_tables_reproduce_the_weights(packed):
    asset, p = packed
    W = np.zeros_like(asset['weights'])
    gs, gj, gw = p['g_start'], p['g_joint'], p['g_w']
    assert gs[0] == 0 and gs[-1] == len(gj) and p['num_groups'] == (6890 + 7) // 8
    for g in range(p['num_groups']):
        js = gj[gs[g]:gs[g + 1]] // 12
        assert np.all(np.diff(js) > 0)                          # sorted, unique joints
        for e in range(gs[g], gs[g + 1]):
            nv = min(8, 6890 - 8 * g)
            W[8 * g:8 * g + nv, gj[e] // 12] = gw[e, :nv]
            assert not gw[e, nv:].any()
    assert np.array_equal(W, asset['weights'])


@pytest.mark.parametrize('nframes,gpb', [(40, 48), (32, 862), (7, 100)])
def test_skin_group_kernel_matches_dense_sum(H, packed, nframes, gpb):
    asset, p = packed
    V, v3_ld = 6890, p['v3_ld']
    rng = np.random.RandomState(nframes)
    # v3_ld = 20 672 < 862 groups * 24 floats: the last group of a row reads 16 floats into the next row (or, for the last
    # row, into the slack the real slab buffer has after it); those belong to vertices >= 6890, carry zero weights and are
    # never stored.  NaN everywhere the GEMM does not write makes any use of them visible.
    flat = np.full(nframes * v3_ld + 32, np.nan, np.float32)
    vposed = flat[:nframes * v3_ld].reshape(nframes, v3_ld)
    assert v3_ld < p['num_groups'] * 24 and v3_ld % 4 == 0
    vposed[:, :3 * V] = rng.randn(nframes, 3 * V).astype(np.float32)
    A = rng.randn(nframes, 52, 3, 4).astype(np.float32)
    trans = rng.randn(nframes, 3).astype(np.float32)
    out = np.full((nframes, V, 3), np.nan, np.float32)
    H.h_lbs_skin_group(V, p['num_groups'], P(p['g_start']), P(p['g_joint']), P(p['g_w']), nframes, v3_ld, P(vposed), P(A),
                       P(trans), P(out), gpb)
    W = asset['weights'].astype(np.float64)
    vp = vposed[:, :3 * V].reshape(nframes, V, 3).astype(np.float64)
    T = np.einsum('vj,njrc->nvrc', W, A.astype(np.float64))     # blended transforms (n,V,3,4)
    ref = np.einsum('nvrc,nvc->nvr', T[..., :3], vp) + T[..., 3] + trans[:, None].astype(np.float64)
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
