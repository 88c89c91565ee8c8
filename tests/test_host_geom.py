"""The device geometry (humor_b200/csrc/*.cuh) compiled for the host and checked against torch autograd:
the exact forward / hand-derived reverse code the CUDA kernels execute."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.smplh_lbs import rodrigues
from oracle import stage3_port as sp

HERE = os.path.dirname(os.path.abspath(__file__))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def L():
    so = os.path.join(HERE, 'host', 'geom_host.so')
    src = os.path.join(HERE, 'host', 'geom_host.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-o', so, src])
    return ctypes.CDLL(so)


def _check(L, fwd_t, cf, cb, x, outdim, ftol, btol):
    n = x.shape[0]
    rng = np.random.RandomState(1)
    xt = torch.tensor(x, requires_grad=True)
    y = fwd_t(xt)
    g = rng.randn(n, outdim).astype(np.float32)
    (y.reshape(n, outdim) * torch.tensor(g)).sum().backward()
    yo = np.zeros((n, outdim), np.float32)
    cf(n, P(x), P(yo))
    gi = np.zeros_like(x)
    cb(n, P(x), P(g), P(gi))
    assert np.abs(yo - y.detach().numpy().reshape(n, outdim)).max() < ftol
    ref = xt.grad.numpy().reshape(n, -1)
    assert np.abs(gi.reshape(n, -1) - ref).max() / np.abs(ref).max() < btol


def test_rodrigues_mat2aa_w2a(L):
    rng = np.random.RandomState(0)
    n = 3000
    aa = (rng.randn(n, 3) * np.array([0.01, 0.5, 2.5])[rng.randint(0, 3, n)][:, None]).astype(np.float32)
    _check(L, lambda t: rodrigues(t), L.h_rodrigues_fwd, L.h_rodrigues_bwd, aa, 9, 2e-6, 2e-5)
    a2 = rng.randn(n, 3).astype(np.float32)
    a2 *= (rng.uniform(0.05, 3.1, (n, 1)) / np.linalg.norm(a2, axis=1, keepdims=True)).astype(np.float32)
    R = rodrigues(torch.tensor(a2)).numpy().reshape(n, 9).copy()
    _check(L, lambda t: sp.mat2aa(t.reshape(n, 3, 3)), L.h_mat2aa_fwd, L.h_mat2aa_bwd, R, 3, 2e-6, 2e-5)
    _check(L, lambda t: sp.world2aligned(t.reshape(n, 3, 3)), L.h_w2a_fwd, L.h_w2a_bwd, R, 9, 2e-6, 2e-5)


def test_glue_rollout_forward_and_bptt(L):
    """glue_step_fwd / glue_step_bwd chained over steps (decoder outputs given) vs autograd through the port's
    decode-composition + canonicalisation + world transform."""
    rng = np.random.RandomState(0)
    B, S = 5, 7
    x0 = np.zeros((B, 339), np.float32)
    x0[:, 0:3] = rng.randn(B, 3) * 0.1 + [0, 0, 0.9]
    x0[:, 3:6] = rng.randn(B, 3) * 0.1
    R = rodrigues(torch.tensor(rng.randn(B * 22, 3).astype(np.float32) * 0.5)).numpy().reshape(B, 22, 9)
    x0[:, 6:15], x0[:, 15:18], x0[:, 18:207] = R[:, 0], rng.randn(B, 3) * 0.1, R[:, 1:].reshape(B, 189)
    x0[:, 207:273], x0[:, 273:339] = rng.randn(B, 66) * 0.3, rng.randn(B, 66) * 0.1
    raws = (rng.randn(S, B, 216) * 0.1).astype(np.float32)
    xt, rt = torch.tensor(x0, requires_grad=True), torch.tensor(raws, requires_grad=True)
    zero = torch.zeros(B, 1)
    t2j_t = -torch.cat([xt[:, 207:209], zero], 1)
    Gr, Gt = torch.eye(3)[None].expand(B, 3, 3), torch.zeros(B, 3)
    past, worlds = xt, []
    names = ['trans', 'trans_vel', 'root_orient', 'root_orient_vel', 'pose_body', 'joints', 'joints_vel']
    for t in range(S):
        i, o = past, rt[t]
        R_in = torch.cat([i[:, 6:15], i[:, 18:207]], 1).reshape(B * 22, 3, 3)
        d_aa = torch.cat([o[:, 6:9], o[:, 12:75]], 1).reshape(B * 22, 3)
        R_out = torch.bmm(rodrigues(d_aa), R_in).reshape(B, 198)
        x = sp._split348(torch.cat([o[:, 0:3] + i[:, 0:3], o[:, 3:6] + i[:, 3:6], R_out[:, :9], o[:, 9:12] + i[:, 15:18],
                                    R_out[:, 9:], o[:, 75:141] + i[:, 207:273], o[:, 141:207] + i[:, 273:339], o[:, 207:216]], 1))
        Ra = sp.world2aligned(x['root_orient'].reshape(B, 3, 3))
        ta = torch.cat([-x['trans'][:, :2], zero], 1)
        nxt, w = sp._rigid(x, Ra, ta, t2j_t, False), sp._rigid(x, Gr, Gt, t2j_t, True)
        Gt, Gr = torch.cat([-w['trans'][:, :2], zero], 1), torch.bmm(Gr, Ra)
        worlds.append(torch.cat([w[k] for k in names + ['contacts']], 1))
        past = torch.cat([nxt[k] for k in names], 1)
    Wt = torch.stack(worlds, 0)
    gw = rng.randn(S, B, 348).astype(np.float32)
    (Wt * torch.tensor(gw)).sum().backward()
    xins = np.zeros((S + 1, B, 416), np.float32)
    xins[0, :, :339] = x0
    rawp = np.zeros((S, B, 224), np.float32)
    rawp[:, :, :216] = raws
    Gs, t2j, wo = np.zeros((S + 1, B, 12), np.float32), np.zeros((B, 3), np.float32), np.zeros((S, B, 348), np.float32)
    t2j[:, :2] = -x0[:, 207:209]
    L.h_glue_rollout_fwd(B, S, P(xins), P(rawp), P(Gs), P(t2j), P(wo))
    assert np.abs(wo - Wt.detach().numpy()).max() < 5e-6
    assert np.abs(xins[S, :, :339] - past.detach().numpy()).max() < 5e-6
    dx0, dr, d2j = np.zeros((B, 339), np.float32), np.zeros((S, B, 216), np.float32), np.zeros((B, 3), np.float32)
    L.h_glue_rollout_bwd(B, S, P(xins), P(rawp), P(Gs), P(t2j), P(gw), P(dx0), P(dr), P(d2j))
    dx0[:, 207:209] -= d2j[:, :2]
    assert np.abs(dx0 - xt.grad.numpy()).max() / np.abs(xt.grad.numpy()).max() < 2e-5
    assert np.abs(dr - rt.grad.numpy()).max() / np.abs(rt.grad.numpy()).max() < 2e-5


def test_lbs_chain_forward_and_reverse(L):
    from humor_b200 import synth
    rng = np.random.RandomState(0)
    n = 6
    par = synth.smplh_parents()
    par32 = par.astype(np.int32).copy()
    par32[0] = -1
    pose = (rng.randn(n, 66) * 0.4).astype(np.float32)
    Jrest = (rng.randn(n, 52, 3) * 0.3).astype(np.float32)
    pt, Jt = torch.tensor(pose, requires_grad=True), torch.tensor(Jrest, requires_grad=True)
    Rm = rodrigues(torch.cat([pt, torch.zeros(n, 90)], 1).reshape(-1, 3)).view(n, 52, 3, 3)
    feat = (Rm[:, 1:22] - torch.eye(3)).reshape(n, 189)
    rel = torch.cat([Jt[:, :1], Jt[:, 1:] - Jt[:, par[1:]]], 1)
    G = torch.cat([torch.cat([Rm, rel[..., None]], -1), torch.tensor([0, 0, 0, 1.]).expand(n, 52, 1, 4)], 2)
    chain = [G[:, 0]]
    for i in range(1, 52):
        chain.append(chain[par[i]] @ G[:, i])
    Gw = torch.stack(chain, 1)
    Jp = Gw[:, :, :3, 3]
    corr = torch.matmul(Gw, torch.cat([Jt, torch.zeros(n, 52, 1)], -1)[..., None])
    A = torch.cat([Gw[..., :3], Gw[..., 3:] - corr], -1)[:, :, :3, :]
    gA, gJ, gF = rng.randn(n, 52, 3, 4).astype(np.float32), rng.randn(n, 52, 3).astype(np.float32), rng.randn(n, 189).astype(np.float32)
    ((A * torch.tensor(gA)).sum() + (Jp * torch.tensor(gJ)).sum() + (feat * torch.tensor(gF)).sum()).backward()
    fo, Ao, Jo = np.zeros((n, 189), np.float32), np.zeros((n, 624), np.float32), np.zeros((n, 156), np.float32)
    L.h_lbs_chain_fwd(n, P(pose), P(Jrest), P(par32), P(fo), P(Ao), P(Jo))
    assert np.abs(fo - feat.detach().numpy()).max() < 1e-6
    assert np.abs(Ao - A.detach().numpy().reshape(n, 624)).max() < 2e-6
    assert np.abs(Jo - Jp.detach().numpy().reshape(n, 156)).max() < 2e-6
    dp, dJ = np.zeros((n, 66), np.float32), np.zeros((n, 156), np.float32)
    L.h_lbs_chain_bwd(n, P(pose), P(Jrest), P(par32), P(gA), P(gJ), P(gF), P(dp), P(dJ))
    assert np.abs(dp - pt.grad.numpy()).max() / np.abs(pt.grad.numpy()).max() < 2e-5
    assert np.abs(dJ - Jt.grad.numpy().reshape(n, 156)).max() / np.abs(Jt.grad.numpy()).max() < 2e-5
