"""The tcgen05 kernels of the dense LBS forward executed on the CPU: SIMT shim + functional emulation of mbarriers, TMA
(SWIZZLE_128B), tcgen05.mma kind::tf32 and TMEM (tests/host/shim/tc_emul.h), same kernel source as the GPU build.

* lbs_fuseg_kernel (blend GEMM + group skinning, the default dense forward) in its two operand forms: 3xTF32 planes and fp16
  hi/lo planes; tile order, slot schedule, descriptor offsets, partial row/column tiles, operand planes past the matrix.
* umma_gemm16_kernel with its GroupNorm epilogue and fp16 output planes."""
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
    so = os.path.join(HERE, 'host', 'tc_host.so')
    src = os.path.join(HERE, 'host', 'tc_host.cpp')
    subprocess.check_call(['g++', '-O2', '-std=c++20', '-pthread', '-shared', '-fPIC', '-I' + os.path.join(HERE, 'host', 'shim'),
                           '-DHB_HOST_SHIM', src, '-o', so])
    L = ctypes.CDLL(so)
    return L


def split_rn(x):
    """x = hi + lo with hi = x rounded to nearest-even on tf32's 10 mantissa bits (what the product's plane kernels write)."""
    u = x.view(np.uint32).astype(np.uint64)
    hi = ((u + 0x0fff + ((u >> 13) & 1)) & 0xffffe000).astype(np.uint32).view(np.float32)
    return np.ascontiguousarray(hi), np.ascontiguousarray(x - hi)


# ---- skin form 3: fused blend + lane = frame group skinning (csrc/lbs_fuseg.cuh)
def _fuseg_problem(N, seed):
    asset = synth.make_smplh_asset()
    p = pack_smplh(asset, 16)
    V, K = 6890, 224
    rng = np.random.RandomState(seed)
    feat = np.zeros((N, K), np.float32)
    feat[:, :16] = rng.randn(N, 16).astype(np.float32) * 0.7            # betas
    feat[:, 16:205] = (rng.randn(N, 189) * 0.3).astype(np.float32)      # R - I at moderate poses
    feat[:, 205] = 1.0                                                  # the pose kernels write 1 here ...
    bt = np.zeros((p['v3_ld'], K), np.float32)
    bt[:, :208] = p['blend_t']
    bt[:3 * V, 205] = p['v_template']                                   # ... and column 205 of the planes carries the template
    A = rng.randn(N, 52, 3, 4).astype(np.float32)
    trans = rng.randn(N, 3).astype(np.float32)
    vp = feat[:, :208].astype(np.float64) @ p['blend'][:, :3 * V].astype(np.float64) + p['v_template'].astype(np.float64)
    W = asset['weights'].astype(np.float64)
    T = np.einsum('vj,njrc->nvrc', W, A.astype(np.float64))
    ref = np.einsum('nvrc,nvc->nvr', T[..., :3], vp.reshape(N, V, 3)) + T[..., 3] + trans[:, None].astype(np.float64)
    return p, feat, bt, A, trans, ref


def _run_fuseg(H, p, feat, bt, A, trans, grid, f16=False, fold=False, fpb=0):
    """The kernel's contract (csrc/umma_launch.cuh LbsFusegArgs): A arrives with its rotation part times the accumulator scale and,
    with fold, the root translation inside its translation column (trans = NULL) - what the pose kernels write for this pass."""
    N, K, V = feat.shape[0], 224, 6890
    fh, fl = split_rn(feat)
    out = np.full((N, V, 3), np.nan, np.float32)
    ntma = ctypes.c_longlong(0)
    H.h_lbs_fuseg.restype = ctypes.c_longlong
    A = A.copy()
    if f16:
        A[..., :3] *= np.float32(2.0 ** -10)
    if fold:
        A[..., 3] += trans[:, None, :]
    rec = np.ascontiguousarray(p['ft_rec'])
    assert rec.ctypes.data % 16 == 0
    tabs = (N, V, p['num_groups'], P(p['ft_tab']), P(rec), rec.shape[1], P(A), None if fold else P(trans), P(out), grid)
    if f16 and fpb:   # ... with one shape per fpb frames: pose columns only (K = 192), template + shape blend per sequence
        nseq = -(-N // fpb)
        assert all((feat[s * fpb:(s + 1) * fpb, :16] == feat[s * fpb, :16]).all() for s in range(nseq))
        fp = np.zeros((N, 192), np.float32)
        fp[:, :189] = feat[:, 16:205]
        bp = np.zeros((bt.shape[0], 192), np.float32)
        bp[:, :189] = bt[:, 16:205] * np.float32(1024)
        vs = np.zeros((nseq, bt.shape[0]), np.float32)
        vs[:] = (feat[::fpb, :16].astype(np.float64) @ bt[:, :16].astype(np.float64).T + bt[:, 205].astype(np.float64)) * 1024.0
        f_h, b_h = fp.astype(np.float16), bp.astype(np.float16)
        f_l, b_l = (fp - f_h.astype(np.float32)).astype(np.float16), (bp - b_h.astype(np.float32)).astype(np.float16)
        keep = [np.ascontiguousarray(x) for x in (f_h, b_h, f_l, b_l, vs)]
        nmma = H.h_lbs_fuseg(P(fh), P(fl), K, P(fh), P(fl), K, bt.shape[0], 0, *tabs, ctypes.byref(ntma), P(keep[0]), P(keep[1]), 192, 3,
                             P(keep[2]), P(keep[3]), P(keep[4]), bt.shape[0], fpb)
    elif f16:   # blend form 5: every column as fp16 hi + unscaled lo planes (K padded to 256), three products, no tf32 k-blocks
        fp = np.zeros((N, 256), np.float32)
        fp[:, :K] = feat
        bp = np.zeros((bt.shape[0], 256), np.float32)
        bp[:, :K] = bt * np.float32(1024)
        f_h, b_h = fp.astype(np.float16), bp.astype(np.float16)
        f_l, b_l = (fp - f_h.astype(np.float32)).astype(np.float16), (bp - b_h.astype(np.float32)).astype(np.float16)
        keep = [np.ascontiguousarray(x) for x in (f_h, b_h, f_l, b_l)]
        nmma = H.h_lbs_fuseg(P(fh), P(fl), K, P(fh), P(fl), K, bt.shape[0], 0, *tabs, ctypes.byref(ntma), P(keep[0]), P(keep[1]), 256, 4,
                             P(keep[2]), P(keep[3]), None, 0, 1)
    else:
        bh, bl = split_rn(bt)
        nmma = H.h_lbs_fuseg(P(fh), P(fl), K, P(bh), P(bl), K, bt.shape[0], K, *tabs, ctypes.byref(ntma), None, None, 0, 0, None, None, None, 0, 1)
    return out, nmma, ntma.value


def test_fuseg_slot_schedule_is_consistent():
    """body_model.fuseg_tables: every slotted entry's joint is resident in its slot whether the CTA arrived incrementally or
    started fresh at that tile; incremental loads never take a slot the previous tile reads."""
    p = pack_smplh(synth.make_smplh_asset(), 16)
    tab, gs, gj, gsl = p['ft_tab'], p['g_start'], p['g_joint'], p['g_slot']
    nct, ng = tab.shape[0], p['num_groups']
    assert nct == (ng + 7) // 8 and len(gsl) == len(gj)
    state = {}                                                   # slot -> joint*12, as left by an incremental walk from tile 0
    for c in range(nct):
        fresh = {int(e) >> 16: int(e) & 0xffff for e in tab[c, 4:4 + tab[c, 0]]}
        inc = {int(e) >> 16: int(e) & 0xffff for e in tab[c, 17:17 + tab[c, 1]]}
        assert len(fresh) == tab[c, 0] <= 13 and all(0 <= s < 13 for s in fresh)
        prev_used = set(state) if c else set()
        assert not (set(inc) & prev_used)                        # a new joint never overwrites a slot tile c-1 may be reading
        kept = {s: j for s, j in state.items() if s in fresh and fresh[s] == j}
        state = {**kept, **inc}
        assert state == fresh                                    # incremental arrival == fresh start
        for g in range(8 * c, min(8 * c + 8, ng)):
            for e in range(gs[g], gs[g + 1]):
                if gsl[e] >= 0:
                    assert gsl[e] % (128 * 48) == 0 and fresh[gsl[e] // (128 * 48)] == gj[e]
    assert (gsl < 0).mean() < 0.05                               # SMPL-like locality: few joints are left to global loads
    # the per-tile records the kernel's producer bulk-copies: the same entries, tile by tile
    rec, gw = p['ft_rec'], p['g_w']
    assert rec.shape[0] == nct and rec.shape[1] % 16 == 0 and rec.shape[1] <= 64 + 48 * 96
    r32 = np.ascontiguousarray(rec).view(np.int32)
    for c in range(nct):
        e0 = gs[min(8 * c, ng)]
        offs = r32[c, :9]
        assert [int(o) for o in offs] == [int(gs[min(8 * c + i, ng)] - e0) for i in range(9)]
        ne = int(offs[8])
        assert tab[c, 2] == 64 + 48 * ne
        body = r32[c, 16:16 + 12 * ne].reshape(ne, 12)
        assert (body[:, 0] == gsl[e0:e0 + ne]).all() and (body[:, 1] == gj[e0:e0 + ne]).all()
        assert (body[:, 4:].view(np.float32) == gw[e0:e0 + ne]).all()


@pytest.mark.parametrize('N,grid,fold', [(200, 3, False), (40, 7, True), (140, 2, True)])
def test_fuseg_kernel_matches_fp64(H, N, grid, fold):
    """The whole mesh (108 column tiles, the last one 64 operand rows past the planes) x 1-2 row tiles (ragged), CTA chunks that
    start in the middle of a row and cross into the next one (fresh slot loads + full drain); blend form 1 (3xTF32)."""
    p, feat, bt, A, trans, ref = _fuseg_problem(N, N + grid)
    out, nmma, ntma = _run_fuseg(H, p, feat, bt, A, trans, grid, fold=fold)
    ntiles = ((N + 127) // 128) * 108
    assert nmma == ntiles * 7 * 4 * 3
    assert np.isfinite(out).all()                                # every vertex of every frame written
    err = np.abs(out - ref).max()
    scale = max(1.0, np.abs(ref).max())
    assert err < 4e-6 * scale, err
    # operand entries (2 TMA boxes each) + transform slots: at most the fresh list per tile, at least the incremental one
    ent = ntiles * 14 * 2
    tab = p['ft_tab']
    nrt = (N + 127) // 128
    assert ent + nrt * int(tab[:, 1].sum()) <= ntma <= ent + nrt * int(tab[:, 1].sum()) + (grid + nrt) * 13


def test_fuseg_kernel_on_a_mesh_without_locality(H):
    """Nothing in the slot schedule assumes SMPL: with skinning weights scattered over 20 joints (up to 2 influences per
    vertex, every 64-vertex tile touching all 20) a third of the group entries find no slot and the epilogue reads their transforms
    from global memory - same result.  A mesh whose tiles need more entries than a record buffer holds (8 scattered influences)
    gets no records: the dispatcher then runs skin form 1."""
    asset = dict(synth.make_smplh_asset())
    V = 6890

    def scattered(kmax, seed, pool=52):
        rng = np.random.RandomState(seed)
        W = np.zeros((V, 52))
        for v in range(V):
            js = rng.choice(pool, size=rng.randint(1, kmax + 1), replace=False)
            W[v, js] = rng.rand(len(js)) + 0.05
        return (W / W.sum(1, keepdims=True)).astype(np.float32)
    asset['weights'] = scattered(8, 5)
    assert pack_smplh(asset, 16)['ft_rec'] is None
    asset['weights'] = scattered(2, 3, pool=20)
    rng = np.random.RandomState(4)
    p = pack_smplh(asset, 16)
    assert p['wk'] == 2 and (p['g_slot'] < 0).mean() > 0.3 and p['ft_tab'][:, 0].max() == 13 and p['ft_rec'] is not None
    N, K = 70, 224
    feat = np.zeros((N, K), np.float32)
    feat[:, :205] = (rng.randn(N, 205) * 0.3).astype(np.float32)
    feat[:, 205] = 1.0
    bt = np.zeros((p['v3_ld'], K), np.float32)
    bt[:, :208] = p['blend_t']
    bt[:3 * V, 205] = p['v_template']
    A = rng.randn(N, 52, 3, 4).astype(np.float32)
    trans = rng.randn(N, 3).astype(np.float32)
    out, nmma, _ = _run_fuseg(H, p, feat, bt, A, trans, 2)
    vp = feat[:, :208].astype(np.float64) @ p['blend'][:, :3 * V].astype(np.float64) + p['v_template'].astype(np.float64)
    T = np.einsum('vj,njrc->nvrc', asset['weights'].astype(np.float64), A.astype(np.float64))
    ref = np.einsum('nvrc,nvc->nvr', T[..., :3], vp.reshape(N, V, 3)) + T[..., 3] + trans[:, None].astype(np.float64)
    assert nmma == 108 * 84 and np.isfinite(out).all()
    assert np.abs(out - ref).max() < 4e-6 * max(1.0, np.abs(ref).max())


# ---- fp16 hi/lo GEMM with its forward epilogues (csrc/umma_gemm16.cuh)
def split16_np(x):
    h = x.astype(np.float16)
    l = ((x - h.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    return np.ascontiguousarray(h), np.ascontiguousarray(l)


@pytest.mark.parametrize('ks,bn', [(1, 64), (4, 64), (1, 128)])          # 128-wide tiles: the batched prior's shape class
@pytest.mark.parametrize('gsize', [64, 32])
def test_umma_gemm16_groupnorm_epilogue_and_fp16_planes(H, ks, bn, gsize):
    """One decoder-style layer from 4-byte operand elements: Linear + bias + GroupNorm + ReLU in the epilogue (x-hat and 1/sigma to
    the tape as the reverse pass expects them), result as fp32 AND as the fp16 hi/lo planes of the next layer; single CTA per tile
    and split-K over a 4-CTA cluster; ragged M and N."""
    M, N, K = 150, 192 - 8, 576
    rng = np.random.RandomState(ks + gsize)
    A = (rng.randn(M, K) * 0.8).astype(np.float32)
    W = (rng.randn(N, K) * 0.04).astype(np.float32)
    # parameter vectors padded to the tile width: the epilogue normalises whole groups of its (padded) 64-column tile
    bias, gamma, beta = (np.ascontiguousarray(rng.randn(256).astype(np.float32) * s) for s in (0.1, 1.0, 0.2))
    gamma = np.ascontiguousarray(gamma + 1.0)
    Ah, Al = split16_np(A)
    Wh, Wl = split16_np(W)
    ldc, ld16, ldxh = 256, 256, 256                            # room for the 128-wide tiles' padded columns (x-hat is written per whole group)
    C = np.full((M, ldc), np.nan, np.float32)
    Ch = np.zeros((M, ld16), np.float16)
    Cl = np.zeros((M, ld16), np.float16)
    xhat = np.full((M, ldxh), np.nan, np.float32)
    rstd = np.full((M, 16), np.nan, np.float32)
    H.h_umma_gemm16.restype = ctypes.c_longlong
    nmma = H.h_umma_gemm16(P(Ah), P(Al), K, P(Wh), P(Wl), K, M, N, K, P(C), ldc, P(Ch), P(Cl), ld16, 1, P(bias), P(gamma), P(beta),
                           P(xhat), ldxh, P(rstd), gsize, ks, bn)
    assert nmma == 2 * -(-N // bn) * (K // 64) * 4 * 3          # row tiles x column tiles x k-blocks x 4 K-steps x (h.h, l.h, h.l)
    y = A.astype(np.float64) @ W.astype(np.float64).T + bias[:N]
    # N = 184 is not a multiple of the group size: the kernel normalises whole groups of its padded tile; compare complete groups
    ng = N // gsize
    yg = y[:, :ng * gsize].reshape(M, ng, gsize)
    mean, var = yg.mean(-1, keepdims=True), yg.var(-1, keepdims=True)
    xh_ref = ((yg - mean) / np.sqrt(var + 1e-5)).reshape(M, ng * gsize)
    out_ref = np.maximum(xh_ref * gamma[:ng * gsize] + beta[:ng * gsize], 0.0)
    nc = ng * gsize
    assert np.abs(C[:, :nc] - out_ref).max() < 5e-6 and np.abs(xhat[:, :nc] - xh_ref).max() < 5e-6
    assert np.abs(rstd[:, :ng] - 1.0 / np.sqrt(var[..., 0] + 1e-5)).max() < 1e-5 * (1.0 / np.sqrt(var.min() + 1e-5))
    rec = Ch.astype(np.float32) + Cl.astype(np.float32) / np.float32(2048)
    assert np.abs(rec[:, :nc] - C[:, :nc]).max() <= 3e-7 * max(1.0, np.abs(C[:, :nc]).max())   # the planes carry the fp32 result to 2^-22
    assert np.isnan(C[:, N:]).all() and not Ch[:, N:].any()    # padding columns untouched


def test_fuseg_kernel_fp16_pose_columns_with_shape_rows(H):
    """One shape per 40 frames (HuMoR: per sub-sequence): the GEMM carries the 189 pose columns (three fp16 k-blocks = 36 MMAs per
    tile), the epilogue adds the sequence's shaped template from the rows the producer staged - a ragged last tile whose frames
    span four sequences, transforms with the translation folded in."""
    N, grid, fpb = 140, 3, 40
    p, feat, bt, A, trans, _ = _fuseg_problem(N, 77)
    feat[:, :16] = np.repeat(feat[::fpb, :16], fpb, axis=0)[:N]       # one shape per sequence
    V = 6890
    vp = feat[:, :208].astype(np.float64) @ p['blend'][:, :3 * V].astype(np.float64) + p['v_template'].astype(np.float64)
    T = np.einsum('vj,njrc->nvrc', synth.make_smplh_asset()['weights'].astype(np.float64), A.astype(np.float64))
    ref = np.einsum('nvrc,nvc->nvr', T[..., :3], vp.reshape(N, V, 3)) + T[..., 3] + trans[:, None].astype(np.float64)
    out, nmma, ntma = _run_fuseg(H, p, feat, bt, A, trans, grid, f16=True, fold=True, fpb=fpb)
    assert nmma == 2 * 108 * 3 * 4 * 3
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() < 4e-6 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


def test_fuseg_kernel_fp16_three_products(H):
    """blend form 5: all 208 feature columns as fp16 hi + unscaled lo planes, h.h + l.h + h.l per 64-wide k-block into the one
    accumulator - the accuracy class of three TF32 passes (1e-6) from 4-byte operand elements: 8 ring entries and 48 MMAs per tile."""
    N, grid = 140, 2
    p, feat, bt, A, trans, ref = _fuseg_problem(N, N + grid)
    out, nmma, ntma = _run_fuseg(H, p, feat, bt, A, trans, grid, f16=True)
    ntiles = 2 * 108
    assert nmma == ntiles * 4 * 4 * 3
    assert np.isfinite(out).all()
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(out - ref).max() < 4e-6 * scale, np.abs(out - ref).max()


# ---- persistent 128x128-tile 3xTF32 GEMM with two epilogue groups (csrc/umma_gemm.cuh: umma_gemm3p_kernel)
@pytest.mark.parametrize('K,grid', [(96, 3), (352, 2), (1024, 5)])        # 1, 3 and 8 promotion chunks; 12 tiles over 2-5 CTAs
def test_persistent_gemm_groupnorm_forward_reverse_and_bias(H, K, grid):
    """The batched prior's shape class: ragged M (5 row tiles + 37 rows) x 256 columns.  CTAs take several tiles each, so both
    epilogue groups, the re-use of their TMEM buffer pairs across tiles (odd and even chunk counts) and the operand ring running
    across tile boundaries are exercised; forward GroupNorm + ReLU with hi/lo output planes and the x-hat / 1/sigma tape, the
    reverse of it, and the plain bias epilogue."""
    M, N = 5 * 128 + 37, 256
    rng = np.random.RandomState(K + grid)
    A = (rng.randn(M, K) * 0.8).astype(np.float32)
    W = (rng.randn(N, K) * (1.5 / np.sqrt(K))).astype(np.float32)
    bias, gamma, beta = (np.ascontiguousarray(rng.randn(N).astype(np.float32) * s) for s in (0.1, 1.0, 0.2))
    gamma = np.ascontiguousarray(gamma + 1.0)
    (Ah, Al), (Wh, Wl) = split_rn(A), split_rn(W)
    H.h_umma_gemm3p.restype = ctypes.c_longlong
    ntiles = 6 * 2
    # forward: Linear + bias + GroupNorm(64) + ReLU
    Chi, Clo = np.full((M, N), np.nan, np.float32), np.full((M, N), np.nan, np.float32)
    xhat, rstd = np.full((M, N), np.nan, np.float32), np.full((M, 16), np.nan, np.float32)
    nmma = H.h_umma_gemm3p(P(Ah), P(Al), K, P(Wh), P(Wl), K, M, N, K, None, P(Chi), P(Clo), N, 1, P(bias), P(gamma), P(beta), P(xhat), N,
                           P(rstd), 64, N, grid)
    assert nmma == ntiles * (K // 32) * 4 * 3
    y = A.astype(np.float64) @ W.astype(np.float64).T + bias
    yg = y.reshape(M, N // 64, 64)
    mean, var = yg.mean(-1, keepdims=True), yg.var(-1, keepdims=True)
    xh_ref = ((yg - mean) / np.sqrt(var + 1e-5)).reshape(M, N)
    out_ref = np.maximum(xh_ref * gamma + beta, 0.0)
    out = Chi.astype(np.float64) + Clo
    assert np.abs(out - out_ref).max() < 3e-6 * max(1.0, np.abs(out_ref).max()) and np.abs(xhat - xh_ref).max() < 3e-6 * np.abs(xh_ref).max()
    assert np.abs(rstd[:, :N // 64] - 1.0 / np.sqrt(var[..., 0] + 1e-5)).max() < 1e-5 * (1.0 / np.sqrt(var.min() + 1e-5))
    assert (Chi.view(np.uint32) & 0x1fff).max() == 0                        # hi plane: tf32-representable
    # reverse: d y (N = Cch) -> d of the GroupNorm input, through ReLU mask and the saved x-hat / 1/sigma
    G = (rng.randn(M, K) * 0.5).astype(np.float32)                        # upstream gradient, as the K-wide operand
    (Gh, Gl) = split_rn(G)
    Dhi, Dlo = np.full((M, N), np.nan, np.float32), np.full((M, N), np.nan, np.float32)
    H.h_umma_gemm3p(P(Gh), P(Gl), K, P(Wh), P(Wl), K, M, N, K, None, P(Dhi), P(Dlo), N, 2, None, P(gamma), P(beta), P(xhat), N, P(rstd), 64, N,
                    grid)
    dy = G.astype(np.float64) @ W.astype(np.float64).T
    u = np.where(xh_ref * gamma + beta > 0.0, dy * gamma, 0.0).reshape(M, N // 64, 64)
    xg = xh_ref.reshape(M, N // 64, 64)
    d_ref = ((u - u.mean(-1, keepdims=True) - xg * (u * xg).mean(-1, keepdims=True)) / np.sqrt(var + 1e-5)).reshape(M, N)
    d = Dhi.astype(np.float64) + Dlo
    # entries whose ReLU argument is within rounding of zero may fall on the other side of the mask
    close = np.abs(xh_ref * gamma + beta) < 1e-5
    bad_rows = close.reshape(M, N // 64, 64).any(-1).repeat(64, -1).reshape(M, N)
    assert np.abs(d - d_ref)[~bad_rows].max() < 2e-5 * max(1.0, np.abs(d_ref).max())
    # bias epilogue, fp32 output, ragged N
    Nr = 256 - 24
    Cb = np.full((M, N), np.nan, np.float32)
    H.h_umma_gemm3p(P(Ah), P(Al), K, P(Wh), P(Wl), K, M, Nr, K, P(Cb), None, None, N, 0, P(bias), None, None, None, 0, None, 64, 0, grid)
    assert np.abs(Cb[:, :Nr] - y[:, :Nr]).max() < 4e-6 * max(1.0, np.abs(y).max()) and np.isnan(Cb[:, Nr:]).all()
