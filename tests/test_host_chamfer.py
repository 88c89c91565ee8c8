"""The chamfer KERNELS (humor_b200/csrc/chamfer.cu) executed on the CPU through the SIMT shim
(tests/host/shim/cuda_runtime.h: one std::thread per CUDA thread, real barriers) and compared bit for bit with the
golden vectors of the compiled reference.  This checks the kernels' tiling, tails, ownership and barrier structure
without a GPU; the -m gpu tests repeat the comparison on the device."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import chamfer as oc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, 'golden', 'chamfer_*.npz')))
P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def H():
    so = os.path.join(HERE, 'host', 'chamfer_host.so')
    src = os.path.join(HERE, 'host', 'chamfer_host.cpp')
    subprocess.check_call(['g++', '-O1', '-std=c++20', '-ffp-contract=off', '-pthread', '-shared', '-fPIC',
                           '-I' + os.path.join(HERE, 'host', 'shim'), '-DHB_HOST_SHIM', src, '-o', so])
    return ctypes.CDLL(so)


def run_nn(H, q, p):
    b, nq, np_ = q.shape[0], q.shape[1], p.shape[1]
    dist = np.full((b, nq), -1.0, np.float32)
    idx = np.full((b, nq), -1, np.int32)
    H.h_chamfer_nn(b, nq, P(q), np_, P(p), P(dist), P(idx))
    return dist, idx


def run_bwd(H, a, c, gd1, i1, gd2, i2, want1=True, want2=True):
    b, n, m = a.shape[0], a.shape[1], c.shape[1]
    g1 = np.full((b, n, 3), np.nan, np.float32) if want1 else None
    g2 = np.full((b, m, 3), np.nan, np.float32) if want2 else None
    H.h_chamfer_bwd(b, n, P(a), m, P(c), P(gd1), P(i1), P(gd2), P(i2), P(g1), P(g2))
    return g1, g2


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_kernels_match_reference_golden(H, path):
    g = np.load(path)
    a, c = np.ascontiguousarray(g['xyz1']), np.ascontiguousarray(g['xyz2'])
    d1, i1 = run_nn(H, a, c)
    d2, i2 = run_nn(H, c, a)
    assert np.array_equal(i1, g['idx1']) and np.array_equal(i2, g['idx2'])
    assert np.array_equal(bits(d1), bits(g['dist1'])) and np.array_equal(bits(d2), bits(g['dist2']))
    gd1, gd2 = np.ascontiguousarray(g['grad_dist1']), np.ascontiguousarray(g['grad_dist2'])
    g1, g2 = run_bwd(H, a, c, gd1, i1, gd2, i2)
    assert np.array_equal(bits(g1), bits(g['grad_xyz1'])) and np.array_equal(bits(g2), bits(g['grad_xyz2']))
    # one-way (points3d_loss): no dist2 gradient, only the predicted cloud's gradient requested
    _, h2 = run_bwd(H, a, c, gd1, i1, None, None, want1=False)
    assert np.array_equal(bits(h2), bits(g['grad_xyz2_oneway']))
    h1, h2 = run_bwd(H, a, c, gd1, i1, None, None)
    assert np.array_equal(bits(h1), bits(g['grad_xyz1_oneway'])) and np.array_equal(bits(h2), bits(g['grad_xyz2_oneway']))


@pytest.mark.parametrize('b,n,m', [(1, 1, 1), (2, 1023, 1025), (1, 1025, 3), (3, 5, 2049), (1, 2050, 1024)])
def test_kernels_match_port_on_ragged_sizes(H, b, n, m):
    rng = np.random.default_rng(n * 7 + m)
    a = rng.normal(size=(b, n, 3)).astype(np.float32)
    c = rng.normal(size=(b, m, 3)).astype(np.float32)
    d1, i1 = run_nn(H, a, c)
    e1, j1 = oc.nn_search(a, c)
    assert np.array_equal(i1, j1) and np.array_equal(bits(d1), bits(e1))
    gd1 = rng.normal(size=(b, n)).astype(np.float32)
    g1, g2 = run_bwd(H, a, c, gd1, i1, None, None)
    o1, o2 = oc.chamfer_backward(a, c, gd1, None, j1, None)
    assert np.array_equal(bits(g1), bits(o1)) and np.array_equal(bits(g2), bits(o2))


def test_nan_and_empty_semantics(H):
    a = np.zeros((1, 3, 3), np.float32)
    c = np.array([[[np.nan, 0, 0], [1, 0, 0], [0.5, 0, 0]]], np.float32)
    d, i = run_nn(H, a, c)                         # NaN at k == 0 sticks (`k == 0 || d < best`)
    assert np.isnan(d).all() and not i.any()
    c2 = np.array([[[2, 0, 0], [np.nan, 0, 0], [1, 0, 0]]], np.float32)
    d, i = run_nn(H, a, c2)                        # later NaNs never win
    assert np.array_equal(d, np.full((1, 3), 1.0, np.float32)) and (i == 2).all()
    e, j = oc.nn_search(a, c2)
    assert np.array_equal(d, e) and np.array_equal(i, j)
    d, i = run_nn(H, a, np.zeros((1, 0, 3), np.float32))
    assert not d.any() and not i.any()
