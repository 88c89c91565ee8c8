"""GPU parity of the whole Stage-III closure (loss, every energy term, every parameter gradient)
against the CPU oracle port, for both benchmark configurations and all three stage-3 phases."""
import numpy as np
import pytest
import torch

from humor_b200 import synth
from tests import util_stage3 as U

pytestmark = pytest.mark.gpu


def compare(optim_floor, B, T, nsteps, scale, seed=4, overlap=3, precision='exact'):
    W = synth.RGB_STAGE3_WEIGHTS if optim_floor else synth.AMASS_STAGE3_WEIGHTS
    prob = synth.make_stage3_problem(B, T, seed=seed, overlap=overlap, cam=optim_floor)
    port = U.build_port(B, T, W, optim_floor, prob)
    if optim_floor:
        # informative 2-D observations: project the oracle's own camera-frame joints
        _, _, aux = U.closure_port(port, prob, optim_floor)
        cj = torch.cat([aux['inter']['cam_pred']['joints3d'], aux['inter']['cam_pred']['joints3d_extra']], 2)
        prob = U.project_joints2d(prob, cj.detach().numpy())
    l_c, g_c, aux_c = U.closure_port(port, prob, optim_floor, nsteps, scale)
    mo = U.build_product(B, T, W, optim_floor, prob)
    mo.set_precision(precision)
    l_g, g_g, aux_g = U.closure_product(mo, prob, nsteps, scale)
    assert abs(l_g - l_c) / max(1.0, abs(l_c)) < 2e-5, (l_g, l_c)
    for k, v in aux_c['stats'].items():
        assert k in aux_g['stats'], k
        assert abs(aux_g['stats'][k] - v) <= 2e-4 * max(1.0, abs(v)), (k, aux_g['stats'][k], v)
    for k in g_c:
        err = float((g_g[k].cpu() - g_c[k]).abs().max() / (g_c[k].abs().max() + 1e-8))
        assert err < (1e-4 if precision == 'exact' else 1e-2), (k, err)
    # vertices of the final camera-frame SMPL evaluation
    v_c = aux_c['inter']['cam_pred']['points3d']
    v_g = aux_g['cam_pred']['points3d'].cpu()
    assert float((v_g - v_c).abs().max()) < 1e-4
    return l_g


@pytest.mark.parametrize('precision', ['exact', 'tensor'])
@pytest.mark.parametrize('nsteps,scale', [(None, 1.0), (4, 1.0), (None, 2.0)])
def test_closure_rgb_config(nsteps, scale, precision):
    compare(True, 4, 8, nsteps, scale, precision=precision)


@pytest.mark.parametrize('precision', ['exact', 'tensor'])
@pytest.mark.parametrize('nsteps,scale', [(None, 1.0), (4, 1.0)])
def test_closure_amass_keypts_config(nsteps, scale, precision):
    compare(False, 4, 7, nsteps, scale, precision=precision)


def test_closure_is_deterministic():
    prob = synth.make_stage3_problem(3, 6, seed=2, overlap=2)
    mo = U.build_product(3, 6, synth.RGB_STAGE3_WEIGHTS, True, prob)
    l1, g1, _ = U.closure_product(mo, prob)
    l2, g2, _ = U.closure_product(mo, prob)
    assert l1 == l2
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k


def test_full_size_properties():
    """BASELINE config 3 size (B=64, T=60).  The oracle takes minutes at that size on CPU, so: (1) sequences are
    independent when the overlap energy is off -> a sub-batch must reproduce its slice of the full batch;
    (2) that sub-batch (8 x 60 frames, full rollout length) is checked against the CPU oracle."""
    B, T = 64, 60
    W = dict(synth.RGB_STAGE3_WEIGHTS)
    W['rgb_overlap_consist'] = 0.0
    prob = synth.make_stage3_problem(B, T, seed=9, overlap=10)
    sub = 8
    p2 = {'params': {k: v[:sub] for k, v in prob['params'].items()}, 'obs': {k: v[:sub] for k, v in prob['obs'].items()},
          'cam_mat': prob['cam_mat'][:sub]}
    port = U.build_port(sub, T, W, True, p2)
    _, _, aux = U.closure_port(port, p2, True)
    cj = torch.cat([aux['inter']['cam_pred']['joints3d'], aux['inter']['cam_pred']['joints3d_extra']], 2).detach().numpy()
    p2 = U.project_joints2d(p2, cj)
    prob['obs']['joints2d'][:sub] = p2['obs']['joints2d']
    l_c, g_c, _ = U.closure_port(port, p2, True)
    mo = U.build_product(B, T, W, True, prob)
    l_full, g_full, aux = U.closure_product(mo, prob)
    assert np.isfinite(l_full)
    for k, g in g_full.items():
        assert torch.isfinite(g).all(), k
    mo2 = U.build_product(sub, T, W, True, p2)
    mo2.set_precision('exact')
    mo.set_precision('exact')
    l_full, g_full, aux = U.closure_product(mo, prob)
    l_sub, g_sub, _ = U.closure_product(mo2, p2)
    assert abs(l_sub - l_c) / max(1.0, abs(l_c)) < 1e-4, (l_sub, l_c)
    for k in g_sub:
        scale = float(g_c[k].abs().max()) + 1e-8
        assert float((g_sub[k].cpu() - g_c[k]).abs().max()) / scale < 5e-3, k          # vs oracle, T=60 BPTT
        assert float((g_sub[k] - g_full[k][:sub]).abs().max()) / scale < 5e-3, k       # slice of the full batch
    # the default 'tensor' mode (tcgen05 3xTF32, persistent chain) through the full T = 60 reverse pass: its gradients agree with
    # the exact-fp32 kernels far better than either agrees with the fp32 CPU oracle (profiles/r02g_tolerances.jsonl)
    mo2.set_precision('tensor')
    l_t, g_t, _ = U.closure_product(mo2, p2)
    assert abs(l_t - l_sub) / max(1.0, abs(l_sub)) < 1e-5
    for k in g_sub:
        scale = float(g_sub[k].abs().max()) + 1e-8
        assert float((g_t[k] - g_sub[k]).abs().max()) / scale < 5e-3, (k, float((g_t[k] - g_sub[k]).abs().max()) / scale)


def test_motion_optimizer_run_smoke():
    """MotionOptimizer.run end to end (all three stages, tiny iteration counts) with the reference's contract."""
    B, T = 2, 8
    prob = synth.make_stage3_problem(B, T, seed=3, overlap=3)
    mo = U.build_product(B, T, synth.RGB_STAGE3_WEIGHTS, True, prob, contact_refine_only=True)
    mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = 4, 1, 2
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    res, stages = mo.run(obs, num_iter=[1, 1, 3], lbfgs_max_iter=2)
    assert res['trans'].shape == (B, T, 3) and res['pose_body'].shape == (B, T, 63) and res['betas'].shape == (B, 16)
    assert res['latent_motion'].shape == (B, T - 1, 48) and res['floor_plane'].shape == (B, 4) and res['contacts'].shape == (B, T, 22)
    assert set(['stage1', 'stage2', 'stage3']) <= set(stages)
    s3 = stages['stage3']
    assert s3['verts3d'].shape == (B, T, 43, 3) and s3['points3d'].shape == (B, T, 6890, 3) and 'prior_trans' in s3
    for v in res.values():
        assert torch.isfinite(v).all()


@pytest.mark.parametrize('precision', ['exact', 'tensor', 'tensor16'])
@pytest.mark.parametrize('name', ['stage3_rgb', 'stage3_rgb_phase1', 'stage3_rgb_refine', 'stage3_amass', 'stage3_rgb_xbatch'])
def test_closure_matches_reference_golden(name, precision):
    """CUDA path against fixtures produced by the UNMODIFIED reference in the build container.
    'exact' (fp32 FFMA GEMMs): every gradient within 1e-4 of its scale.  'tensor' (tcgen05 3xTF32): forward states,
    loss and energy terms to the same bounds; gradients through the reverse rollout within 1e-2 (measured ~3e-3:
    the BPTT amplifies the tensor core's product rounding)."""
    from tests.golden_util import load_case, check_against_golden
    g, prob, c = load_case(name)
    mo = U.build_product(c['B'], c['T'], c['W'], c['optim_floor'], prob)
    mo.set_precision(precision)
    loss, grads, aux = U.closure_product(mo, prob, c['nsteps'], c['scale'])
    check_against_golden(g, loss, aux['stats'], grads, loss_tol=1e-5, stat_tol=1e-4, grad_tol=1e-4 if precision == 'exact' else 1e-2)
    assert np.abs(aux['cam_pred']['verts3d'].detach().cpu().numpy() - g['cam_verts3d']).max() < 1e-4
    assert np.abs(aux['roll']['trans'].detach().cpu().numpy() - g['rollout_trans']).max() < 2e-5
    pm = aux['roll']['cond_prior'][0].detach().cpu().numpy()
    assert np.abs(pm - g['cond_prior_mean']).max() / np.abs(g['cond_prior_mean']).max() < 1e-5


def test_cuda_graph_step_matches_eager():
    """stage3_step with use_cuda_graph replays a captured forward+backward: same loss and gradients as eager,
    and it tracks in-place parameter updates (what L-BFGS does between evaluations)."""
    B, T = 4, 8
    prob = synth.make_stage3_problem(B, T, seed=6, overlap=3)
    mo = U.build_product(B, T, synth.RGB_STAGE3_WEIGHTS, True, prob)
    names = mo.set_stage3_state(prob['params'])
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    params = [getattr(mo, n) for n in names]
    mo.use_cuda_graph = False
    l_e = float(mo.stage3_step(obs, params=params))
    g_e = [p.grad.clone() for p in params]
    mo.use_cuda_graph = True
    l_g = float(mo.stage3_step(obs, params=params))
    assert mo.use_cuda_graph and len(mo._graphs) == 1, 'capture fell back to eager'
    assert abs(l_g - l_e) <= 1e-6 * abs(l_e)
    for a, p in zip(g_e, params):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6 * float(a.abs().max()))
    with torch.no_grad():
        mo.latent_motion.mul_(0.9)
        mo.betas.add_(0.05)
    l_g2 = float(mo.stage3_step(obs, params=params))        # replay on the updated values
    g_g2 = [p.grad.clone() for p in params]
    mo.use_cuda_graph = False
    l_e2 = float(mo.stage3_step(obs, params=params))
    assert abs(l_g2 - l_e2) <= 1e-6 * abs(l_e2) and abs(l_g2 - l_g) > 1e-3
    for a, p in zip(g_g2, params):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6 * float(a.abs().max()))


@pytest.mark.parametrize('graph', [False, True])
def test_dense_vertices_queued_behind_the_reverse_chain_are_the_same_vertices(graph):
    """The gradient-free dense LBS pass of a closure evaluation is queued on the side stream behind the launch of the reverse decoder
    chain (short CTAs on the SMs the chain leaves idle).  What it writes must be what the immediate pass writes and
    the loss must not change - eagerly and inside the captured graph - and a forward without a reverse pass must still get its
    vertices (join_dense queues the pass itself)."""
    B, T = 128, 12                                              # >= 128 frames: the tensor-core dense path
    prob = synth.make_stage3_problem(B, T, seed=9, overlap=3, cam=True)
    mo = U.build_product(B, T, synth.RGB_STAGE3_WEIGHTS, True, prob)
    names = mo.set_stage3_state(prob['params'])
    params = [getattr(mo, n) for n in names]
    obs = {k: U.obs_to(v, mo.device) for k, v in prob['obs'].items() if k in U.obs_keys(True, prob)}
    mo.use_cuda_graph = False
    loss0, _, _, _, cam_pred0 = mo.stage3_forward(obs, None, 1.0)           # immediate placement
    mo.join_dense()
    torch.cuda.synchronize()
    v0 = cam_pred0['points3d'].clone()
    stash, seen = {}, []
    orig = mo.stage3_forward

    def spy(*a, **k):
        out = orig(*a, **k)
        stash['v'] = out[4]['points3d']
        seen.append(mo._dense_deferred is not None)             # still un-queued when the forward returns
        return out

    mo.stage3_forward = spy
    if not graph:
        loss1, _, _ = mo._eval_on_aliases(obs, None, 1.0, params)
    else:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            mo._eval_on_aliases(obs, None, 1.0, params)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss1, _, _ = mo._eval_on_aliases(obs, None, 1.0, params)
        stash['v'].zero_()
        g.replay()
    torch.cuda.synchronize()
    assert all(seen) and mo._dense_deferred is None
    v1 = stash['v']
    assert torch.isfinite(v1).all() and torch.equal(v1.reshape(v0.shape), v0)
    assert abs(float(loss1) - float(loss0)) <= 1e-6 * abs(float(loss0))
    # a forward with no reverse pass behind it
    mo.stage3_forward = orig
    mo._defer_dense_join = True
    try:
        _, _, _, _, cam_pred2 = mo.stage3_forward(obs, None, 1.0)
        assert mo._dense_deferred is not None
    finally:
        mo._defer_dense_join = False
        mo.join_dense()
    torch.cuda.synchronize()
    assert torch.equal(cam_pred2['points3d'], v0)
