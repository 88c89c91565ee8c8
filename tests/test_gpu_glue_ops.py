"""GPU parity of the two per-sequence glue kernel pairs of round 2 against the torch-op forms they replaced:
humor_cam2prior_fwd/_bwd (fitting_utils.compute_cam2prior_torch, float64 on the CPU) and humor_rollout_outputs_fwd/_bwd
(MotionOptimizer._rollout_outputs_torch on the GPU).  The CPU twins run on the emulation: tests/test_emul_product.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cam2prior_kernel_pair_matches_torch_float64():
    from humor_b200 import fitting_utils as FU
    B = 300
    rng = np.random.RandomState(5)
    floor = rng.randn(B, 3) * np.array([0.3, 1.0, 0.3]) * (1.0 + rng.rand(B, 1) * 2.0)
    floor[::2, 1] = -np.abs(floor[::2, 1]) - 0.2
    floor[1::2, 1] = np.abs(floor[1::2, 1]) + 0.2
    trans, orient, joints = rng.randn(B, 3) * 1.5, rng.randn(B, 3) * 0.9, rng.randn(B, 22, 3)
    gR, gt, gh = rng.randn(B, 3, 3), rng.randn(B, 3), rng.randn(B, 1)

    def run(fn, dtype, dev):
        v = [torch.tensor(x, dtype=dtype, device=dev, requires_grad=True) for x in (floor, trans, orient, joints)]
        R, t, h = fn(*v)
        w = [torch.tensor(x, dtype=dtype, device=dev) for x in (gR, gt, gh)]
        ((R * w[0]).sum() + (t * w[1]).sum() + (h * w[2]).sum()).backward()
        return [x.detach().double().cpu().numpy() for x in (R, t, h)], [x.grad.double().cpu().numpy() for x in v]

    def torch64(f, t, r, j):
        K = torch.zeros(B, 3, 3, dtype=torch.float64)
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -r[:, 2], r[:, 1], r[:, 2], -r[:, 0], -r[:, 1], r[:, 0]
        Rm = torch.linalg.matrix_exp(K)
        saved = FU.batch_rodrigues
        FU.batch_rodrigues = lambda aa: Rm
        try:
            return FU.compute_cam2prior_torch(f, t, r, j)
        finally:
            FU.batch_rodrigues = saved

    (o_ref, g_ref), (o_k, g_k) = run(torch64, torch.float64, 'cpu'), run(FU.compute_cam2prior, torch.float32, 'cuda')
    for a, b in zip(o_ref, o_k):
        assert np.abs(a - b).max() < 2e-5                      # 5e-6 measured; the largest on sequences whose right axis nearly lies in the floor normal
    for a, b in zip(g_ref, g_k):
        assert np.abs(a - b).max() < 1e-4 * np.abs(a).max()
    assert np.abs(g_k[3][:, 1:]).max() == 0.0                    # only the root joint is read


@pytest.mark.parametrize('cam', [True, False])
def test_rollout_outputs_kernel_pair_matches_the_torch_form(cam):
    from humor_b200 import motion_optimizer as MO
    from humor_b200.transforms import batch_rodrigues
    B, S = 37, 59
    g = torch.Generator().manual_seed(3)
    rn = lambda *sh: torch.randn(*sh, generator=g)

    class Stub:
        _rollout_outputs_torch = MO.MotionOptimizer._rollout_outputs_torch
        apply_cam2prior = MO.MotionOptimizer.apply_cam2prior

    m = Stub()
    m.optim_floor = cam
    m._contact_idx = torch.tensor(MO.CONTACT_INDS, dtype=torch.long, device='cuda')
    m._contact_idx32 = m._contact_idx.to(torch.int32)
    m.init_fidx, m.cam2prior_root_height = np.zeros(B), None
    world = rn(S, B, 348).cuda()
    R_all = batch_rodrigues((rn(S * B * 22, 3) * 0.9).cuda()).reshape(S, B, 22 * 9)
    world[..., 6:15], world[..., 18:207] = R_all[..., :9], R_all[..., 9:]
    base = {'world': world, 'trans': rn(B, 1, 3).cuda(), 'root_orient': (rn(B, 1, 3) * 0.8).cuda(), 'body_pose': (rn(B, 1, 63) * 0.5).cuda(),
            'joints': rn(B, 1, 22, 3).cuda(), 'R': batch_rodrigues((rn(B, 3) * 0.9).cuda()).reshape(B, 3, 3), 't': rn(B, 3).cuda()}
    keys = ['trans', 'root_orient', 'pose_body', 'joints', 'contacts_logits']
    res, weights = {}, None
    for form in ('kernel', 'torch'):
        v = {k: x.detach().clone().requires_grad_(True) for k, x in base.items()}
        m.cam2prior_R, m.cam2prior_t = v['R'], v['t']
        if form == 'kernel':
            r = MO._RolloutOutputs.apply(v['world'], v['trans'][:, 0], v['root_orient'][:, 0], v['body_pose'][:, 0], v['joints'][:, 0].reshape(B, 66),
                                         v['R'] if cam else None, v['t'] if cam else None, m._contact_idx32)
            o = dict(zip(keys + ['contacts_conf', 'contacts'], r[:7]))
            c = {'trans': r[7], 'root_orient': r[8]} if cam else {'trans': r[0], 'root_orient': r[1]}
        else:
            o, c = m._rollout_outputs_torch(v['world'], None, v['trans'], v['root_orient'], v['body_pose'], None, v['joints'], None, None, None,
                                            False, False)
        if weights is None:
            weights = {k: torch.randn(o[k].shape, generator=g).cuda() for k in keys}
            weights.update({'cam_' + k: torch.randn(c[k].shape, generator=g).cuda() for k in ('trans', 'root_orient')})
        (sum((o[k] * weights[k]).sum() for k in keys) + sum((c[k] * weights['cam_' + k]).sum() for k in ('trans', 'root_orient'))).backward()
        res[form] = ({**{k: o[k].detach() for k in keys + ['contacts_conf', 'contacts']}, **{'cam_' + k: c[k].detach() for k in c}},
                     {k: x.grad for k, x in v.items()})
    for k in res['kernel'][0]:
        assert float((res['kernel'][0][k] - res['torch'][0][k]).abs().max()) < 2e-6, k
    for k, gt in res['torch'][1].items():
        gk = res['kernel'][1][k]
        if gt is None:
            assert gk is None, k
        else:
            assert float((gk - gt).abs().max()) < 2e-5 * float(gt.abs().max()), k
