"""ORACLE (test infrastructure) — drive the UNMODIFIED reference's Stage-III closure.

The closure in the reference is a nested function (motion_optimizer.py:514-608) that cannot
be called from outside ``MotionOptimizer.run``.  ``stage3_closure`` below performs the same
calls, in the same order, on a reference ``MotionOptimizer`` whose optimisation variables were
set by hand — every arithmetic step is executed by the reference's own methods.
Container-only (needs /root/reference).
"""
import os
import tempfile
import numpy as np
import torch

from oracle import ref_import
from humor_b200 import synth

_ASSET_PATH = {}


def asset_npz(seed=0):
    """npz of the synthetic asset, keyed by its content so a changed generator never reuses a stale file."""
    if seed not in _ASSET_PATH:
        import hashlib
        asset = synth.make_smplh_asset(seed)
        tag = hashlib.sha1(asset['v_template'].tobytes() + asset['weights'].tobytes()).hexdigest()[:12]
        p = os.path.join(tempfile.gettempdir(), f'humor_b200_smplh_seed{seed}_{tag}.npz')
        if not os.path.exists(p):
            np.savez(p, **asset)
        _ASSET_PATH[seed] = p
    return _ASSET_PATH[seed]


class _RefChamferFn(torch.autograd.Function):
    """The CPU branch of the reference's ChamferDistanceFunction (utils/chamfer_distance/chamfer_distance.py:13-55)
    driving the reference's own compiled C++ (oracle/_ref/cd_ref.so, oracle/build_ref.py).  The reference's Python
    wrapper cannot be imported as is: it JIT-builds the .cpp AND the .cu at import (chamfer_distance.py:10)."""

    @staticmethod
    def forward(ctx, cd, xyz1, xyz2):
        b, n, _ = xyz1.size()
        m = xyz2.size(1)
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        dist1, dist2 = torch.zeros(b, n), torch.zeros(b, m)
        idx1, idx2 = torch.zeros(b, n, dtype=torch.int), torch.zeros(b, m, dtype=torch.int)
        cd.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        ctx.cd = cd
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gx1, gx2 = torch.zeros(xyz1.size()), torch.zeros(xyz2.size())
        ctx.cd.backward(xyz1, xyz2, gx1, gx2, g1.contiguous(), g2.contiguous(), idx1, idx2)
        return None, gx1, gx2


class RefChamfer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        from oracle.build_ref import build as build_cd, load_cd_ref
        build_cd()
        self.cd = load_cd_ref()
        if self.cd is None:
            raise RuntimeError('reference chamfer module could not be built (oracle/build_ref.py)')

    def forward(self, xyz1, xyz2):
        return _RefChamferFn.apply(self.cd, xyz1.detach() if not xyz1.requires_grad else xyz1, xyz2)


def build(B, T, weights, optim_floor, cam_mat=None, humor_sd=None, gmm=None, vposer=None,
          stage3_contact_refine_only=True):
    """Reference BodyModel + HumorModel + MotionOptimizer on CPU with synthetic assets."""
    ref = ref_import.load()
    ref_import.quiet_logger(ref)
    import io
    import contextlib
    dev = torch.device('cpu')
    with contextlib.redirect_stdout(io.StringIO()):
        bm = ref.body_model.BodyModel(bm_path=asset_npz(), num_betas=16, batch_size=B * T,
                                      use_vtx_selector=optim_floor).to(dev)
        humor = ref.humor_model.HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48,
                                           model_data_config='smpl+joints+contacts', steps_in=1)
    humor.load_state_dict(humor_sd if humor_sd is not None else synth.make_humor_state_dict())
    humor.eval()
    gmm = gmm if gmm is not None else synth.make_gmm()
    vposer = vposer if vposer is not None else synth.FakeVPoser()
    stage_w = [dict(x) for x in weights] if isinstance(weights, (list, tuple)) else [dict(weights)] * 3
    w = {k: max(x[k] for x in stage_w) for k in stage_w[0]}
    mo = ref.motion_optimizer.MotionOptimizer(
        dev, bm, 16, B, T, ['joints2d'] if optim_floor else ['verts3d'], [dict(x) for x in stage_w],
        vposer, humor, {'gmm': gmm}, optim_floor,
        None if cam_mat is None else torch.as_tensor(cam_mat),
        'bisquare', 4.6851, 100.0,
        stage3_contact_refine_only=stage3_contact_refine_only)
    if w.get('points3d', 0.0) > 0.0:
        # what FittingLoss(use_chamfer=True) sets up (fitting_loss.py:52-54), minus the import-time JIT build
        mo.fitting_loss.chamfer_dist = RefChamfer()
    mo.fitting_loss.set_stage(2)
    return ref, mo, bm, humor


def set_params(mo, params, requires_grad=True):
    """Install stage-3 variables (motion_optimizer.py:380-404)."""
    t = lambda k: torch.as_tensor(params[k]).clone().requires_grad_(requires_grad)
    mo.trans, mo.root_orient, mo.latent_pose, mo.betas = t('trans'), t('root_orient'), t('latent_pose'), t('betas')
    mo.latent_motion = t('latent_motion')
    mo.trans_vel, mo.joints_vel, mo.root_orient_vel = t('trans_vel'), t('joints_vel'), t('root_orient_vel')
    if mo.optim_floor:
        mo.floor_plane = t('floor_plane')
    names = ['trans', 'root_orient', 'latent_pose', 'betas', 'latent_motion',
             'trans_vel', 'joints_vel', 'root_orient_vel'] + (['floor_plane'] if mo.optim_floor else [])
    return names


def stage3_closure(ref, mo, observed_data, nsteps=None, init_motion_scale=1.0, backward=True):
    """motion_optimizer.py:514-608 with the phase logic reduced to (nsteps, init_motion_scale).

    nsteps=None  -> full sequence (iterations >= stage3_tune_init_freeze_start).
    Returns loss, stats_dict, dict of intermediate tensors.
    """
    prior_opt_params = [mo.trans_vel, mo.joints_vel, mo.root_orient_vel]
    T = mo.seq_len
    cur_body_pose = mo.latent2pose(mo.latent_pose)
    if mo.optim_floor:
        cam_smpl_data, _ = mo.smpl_results(mo.trans, mo.root_orient, cur_body_pose, mo.betas)
        mo.cam2prior_R, mo.cam2prior_t, mo.cam2prior_root_height = ref.fitting_utils.compute_cam2prior(
            mo.floor_plane, mo.trans[:, 0], mo.root_orient[:, 0], cam_smpl_data['joints3d'][:, 0])
    z = mo.latent_motion
    saved_ov = mo.fitting_loss.loss_weights['rgb_overlap_consist']
    if nsteps is not None:
        z = z[:, :nsteps - 1]
    rollout, cam_rollout = mo.rollout_latent_motion(mo.trans, mo.root_orient, cur_body_pose, mo.betas,
                                                    prior_opt_params, z, return_prior=mo.cond_prior,
                                                    fit_gender='neutral')
    cur_latent_pose = mo.pose2latent(rollout['pose_body'])
    pred, _ = mo.smpl_results(rollout['trans'], rollout['root_orient'], rollout['pose_body'], mo.betas)
    pred['latent_pose'] = cur_latent_pose
    pred['betas'] = mo.betas
    pred['latent_motion'] = z
    pred['joints_vel'], pred['trans_vel'], pred['root_orient_vel'] = mo.joints_vel, mo.trans_vel, mo.root_orient_vel
    pred['joints3d_rollout'] = rollout['joints']
    pred['contacts'], pred['contacts_conf'] = rollout['contacts'], rollout['contacts_conf']
    cam_pred = pred
    if mo.optim_floor:
        cam_pred, _ = mo.smpl_results(cam_rollout['trans'], cam_rollout['root_orient'], rollout['pose_body'], mo.betas)
        cam_pred['latent_pose'] = cur_latent_pose
        cam_pred['betas'] = mo.betas
        cam_pred['floor_plane'] = mo.floor_plane
    loss_obs, loss_nsteps = observed_data, T
    if nsteps is not None:
        loss_obs = {k: v[:, :nsteps] for k, v in observed_data.items() if k != 'prev_batch_overlap_res'}
        loss_nsteps = nsteps
        mo.fitting_loss.loss_weights['rgb_overlap_consist'] = 0.0
    loss, stats = mo.fitting_loss.motion_fit(loss_obs, pred, cam_pred, loss_nsteps,
                                             cond_prior=rollout.get('cond_prior'),
                                             init_motion_scale=init_motion_scale)
    mo.fitting_loss.loss_weights['rgb_overlap_consist'] = saved_ov
    if backward:
        loss.backward()
    inter = {'rollout': rollout, 'cam_rollout': cam_rollout, 'pred': pred, 'cam_pred': cam_pred,
             'body_pose0': cur_body_pose}
    return loss, stats, inter


def set_params12(mo, params, stage):
    """Stage-I/II variables with the reference's requires_grad pattern (motion_optimizer.py:224-228,276-280)."""
    full = stage == 1
    t = lambda k, g: torch.as_tensor(params[k]).clone().requires_grad_(g)
    mo.trans, mo.root_orient = t('trans', True), t('root_orient', True)
    mo.betas, mo.latent_pose = t('betas', full), t('latent_pose', full)
    return ['trans', 'root_orient'] + (['betas', 'latent_pose'] if full else [])


def stage12_closure(ref, mo, observed_data, stage, backward=True):
    """The closure bodies of motion_optimizer.py:237-250 (stage 0) and :289-304 (stage 1), executed by the reference's own
    methods on a reference MotionOptimizer."""
    mo.fitting_loss.set_stage(stage)
    body_pose = mo.latent2pose(mo.latent_pose)
    pred, _ = mo.smpl_results(mo.trans, mo.root_orient, body_pose, mo.betas)
    if stage == 1:
        pred['latent_pose'] = mo.latent_pose
        pred['betas'] = mo.betas
        loss, stats = mo.fitting_loss.smpl_fit(observed_data, pred, mo.seq_len)
    else:
        loss, stats = mo.fitting_loss.root_fit(observed_data, pred)
    if backward:
        loss.backward()
    return loss, stats, {'pred': pred, 'body_pose': body_pose}
