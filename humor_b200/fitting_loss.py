"""FittingLoss — the Stage I/II/III energies of humor/fitting/fitting_loss.py:20-518 evaluated by
ONE fused sm_100a kernel (csrc/losses.cu) that also produces every gradient, plus the GMM
init-state prior kernel.  Same stage/weight bookkeeping as the reference (``loss_weights`` is the
mutable per-stage dict MotionOptimizer edits; a weight of 0 removes the term exactly).
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _ext
from ._ext import TERM, HB_NUM_TERMS
from .chamfer import ChamferDistance
from .fitting_utils import apply_robust_weighting_sq

SMPL2OP = [52, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62]
# reference weight key -> (term ids sharing it)
WEIGHT_TERMS = {
    'joints2d': ['joints2d'], 'joints3d': ['joints3d'], 'verts3d': ['verts3d'],
    'rgb_overlap_consist': ['rgb_overlap_consist_verts3d_pos', 'rgb_overlap_consist_verts3d_vel',
                            'rgb_overlap_consist_betas', 'rgb_overlap_consist_floor'],
    'pose_prior': ['pose_prior'], 'shape_prior': ['shape_prior'], 'joints3d_smooth': ['joints3d_smooth'],
    'motion_prior': ['motion_prior'], 'init_motion_prior': ['init_motion_prior'],
    'joint_consistency': ['joint_consistency'], 'bone_length': ['bone_length'],
    'joints3d_rollout': ['joints3d_rollout'], 'contact_vel': ['contact_vel'], 'contact_height': ['contact_height'],
    'floor_reg': ['floor_reg'],
}

_GRAD_SLOTS = ['cam_joints', 'cam_verts', 'prior_joints', 'roll_joints', 'contact_logits', 'betas', 'floor', 'z',
               'prior_out', 'latent_pose']


class _FitFn(torch.autograd.Function):
    """inputs (10 tensors, some None) -> (loss scalar, terms[24]); gradients come from the same kernel."""

    @staticmethod
    def forward(ctx, cfg, *inp):
        L = _ext.lib()
        t = dict(zip(_GRAD_SLOTS, [None if x is None else _ext.f32c(x) for x in inp]))
        _ext.require_cuda(t['cam_joints'])
        dev = t['cam_joints'].device
        B, T, njx = cfg['B'], cfg['T'], cfg['njx']
        a = _ext.HbFitArgs()
        a.B, a.T, a.njx, a.T_obs = B, T, njx, cfg['T_obs']
        for i in range(HB_NUM_TERMS):
            a.coef[i] = float(cfg['coef'][i])
        a.sigma2d = float(cfg['sigma2d'])
        grads = {}
        for k in _GRAD_SLOTS:
            setattr(a, k, None if t[k] is None else t[k].data_ptr())
            grads[k] = None if t[k] is None else torch.empty_like(t[k])
            setattr(a, 'd_' + k, None if grads[k] is None else grads[k].data_ptr())
        if T > 1 and grads['contact_logits'] is not None:
            pass
        for k in ('obs_joints2d', 'obs_joints3d', 'obs_verts3d', 'obs_floor', 'seq_interval', 'cam_f', 'cam_c'):
            v = cfg.get(k)
            setattr(a, k, None if v is None else v.data_ptr())
        terms = torch.empty(HB_NUM_TERMS, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        partials = torch.empty(B * T + 64, HB_NUM_TERMS, device=dev, dtype=torch.float32)      # rows + the 64 block sums of level 1
        a.terms, a.loss, a.partials = terms.data_ptr(), loss.data_ptr(), partials.data_ptr()
        nl = C.c_int64(0)
        _ext.check(L.humor_fit_losses(C.byref(a), C.byref(nl), _ext.stream_ptr()), 'humor_fit_losses')
        _ext.LaunchCounter.total += nl.value
        ctx.grads = [grads[k] for k in _GRAD_SLOTS]
        ctx.unit = cfg.get('assume_unit_grad', False)
        ctx.mark_non_differentiable(terms)
        return loss[0], terms

    @staticmethod
    def backward(ctx, dloss, dterms):
        if ctx.unit:
            return (None,) + tuple(ctx.grads)
        return (None,) + tuple(None if g is None else g * dloss for g in ctx.grads)


class _GmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gmm, x):
        x = _ext.f32c(x)
        B, D = x.shape
        nll = torch.empty(B, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        L = _ext.lib()
        ws = gmm.get('_ws')
        if ws is None or ws[0] != (B, D):                      # scratch of the three-launch form, kept with the constants
            ws = ((B, D), torch.empty(L.humor_gmm_workspace_bytes(B, D, gmm['K']) // 4, device=x.device, dtype=torch.float32))
            gmm['_ws'] = ws
        _ext.check(L.humor_gmm_nll_ws(B, D, gmm['K'], _ext.ptr(x), _ext.ptr(gmm['logw']), _ext.ptr(gmm['mean']), _ext.ptr(gmm['Linv']),
                                      _ext.ptr(gmm['logdet']), _ext.ptr(nll), _ext.ptr(dx), _ext.ptr(ws[1]), ws[1].numel() * 4,
                                      _ext.stream_ptr()), 'humor_gmm_nll_ws')
        _ext.LaunchCounter.total += 3
        ctx.dx = dx
        return nll

    @staticmethod
    def backward(ctx, dnll):
        return None, ctx.dx * dnll[:, None]


def build_gmm(weights, means, covs):
    """torch.distributions MixtureSameFamily(Categorical(w), MultivariateNormal(means, covs)) constants
    (fitting_loss.py:85-89) in the layout of humor_gmm_nll: log mixing weights, means, inverse Cholesky
    factors and sum(log diag L).  Factorised once in fp64 on the host side of the device."""
    w = weights.double()
    Lc = torch.linalg.cholesky(covs.double())
    eye = torch.eye(Lc.shape[-1], dtype=torch.float64, device=Lc.device).expand_as(Lc)
    Linv = torch.linalg.solve_triangular(Lc, eye, upper=False)
    f = lambda t: t.float().contiguous()
    return {'K': int(w.shape[0]), 'logw': f(torch.log(w / w.sum())), 'mean': f(means),
            'Linv': f(Linv), 'logdet': f(torch.log(torch.diagonal(Lc, dim1=-2, dim2=-1)).sum(-1))}


class FittingLoss(nn.Module):
    """Reference-compatible shell (fitting_loss.py:20-92) around the fused kernel."""

    def __init__(self, loss_weights, init_motion_prior=None, smpl2op_map=None, ignore_op_joints=None, cam_f=None,
                 cam_cent=None, robust_loss='none', robust_tuning_const=4.6851, joints2d_sigma=100, use_chamfer=False):
        super().__init__()
        # fitting_loss.py:52-54: the point-cloud term needs the native nearest-neighbour module.  Only obs -> pred is
        # consumed (fitting_loss.py:385-392), so the other direction of the search is skipped.
        self.chamfer_dist = ChamferDistance(one_way=True) if use_chamfer else None
        if smpl2op_map is not None and list(smpl2op_map) != SMPL2OP:
            raise NotImplementedError('the fused kernel hard-wires the SMPL+H -> OpenPose BODY_25 map')
        self.all_stage_loss_weights = loss_weights
        self.cur_stage_idx = 0
        self.loss_weights = self.all_stage_loss_weights[0]
        self.cam_f, self.cam_cent = cam_f, cam_cent
        self.can_reproj = cam_f is not None and cam_cent is not None
        self.joints2d_sigma = joints2d_sigma
        self.robust_loss, self.robust_tuning_const = robust_loss, robust_tuning_const
        self.gmm = None
        tot = sum(w['init_motion_prior'] for w in loss_weights)
        if init_motion_prior is not None and tot > 0.0:
            self.gmm = build_gmm(*init_motion_prior['gmm'])
        self.cur_optim_step = 0
        self.assume_unit_grad = False
        self._dummies = {}

    def set_stage(self, idx):
        self.cur_stage_idx = idx
        self.loss_weights = self.all_stage_loss_weights[idx]

    def _dummy(self, key, shape, dev):
        d = self._dummies.get((key, shape))
        if d is None or d.device != dev:
            d = torch.zeros(shape, device=dev, dtype=torch.float32)
            self._dummies[(key, shape)] = d
        return d

    def coefficients(self, nsteps, init_motion_scale, have):
        """weight*scale per term (0 disables).  `have` = set of available inputs/observations."""
        w = self.loss_weights
        c = [0.0] * HB_NUM_TERMS

        def on(key):
            return w.get(key, 0.0) > 0.0

        if on('joints2d') and 'joints2d' in have:
            c[TERM['joints2d']] = w['joints2d']
        if on('joints3d') and 'joints3d' in have:
            c[TERM['joints3d']] = w['joints3d']
        if on('verts3d') and 'verts3d' in have:
            c[TERM['verts3d']] = w['verts3d']
        if on('rgb_overlap_consist') and 'seq_interval' in have:
            c[TERM['rgb_overlap_consist_verts3d_pos']] = w['rgb_overlap_consist']
            c[TERM['rgb_overlap_consist_verts3d_vel']] = w['rgb_overlap_consist']
            if 'smpl' in have:
                c[TERM['rgb_overlap_consist_betas']] = w['rgb_overlap_consist']
            if 'floor' in have and 'motion' in have:
                c[TERM['rgb_overlap_consist_floor']] = w['rgb_overlap_consist']
        if 'smpl' in have:
            if on('pose_prior') and 'latent_pose' in have:
                c[TERM['pose_prior']] = w['pose_prior']
            if on('shape_prior'):
                c[TERM['shape_prior']] = w['shape_prior'] * nsteps
            if on('joints3d_smooth'):
                c[TERM['joints3d_smooth']] = w['joints3d_smooth']
        if 'motion' in have:
            if on('motion_prior'):
                c[TERM['motion_prior']] = w['motion_prior']
            if on('joint_consistency'):
                c[TERM['joint_consistency']] = w['joint_consistency']
            if on('bone_length'):
                c[TERM['bone_length']] = w['bone_length']
            if on('joints3d_rollout') and 'joints3d' in have:
                c[TERM['joints3d_rollout']] = w['joints3d_rollout']
            if on('contact_vel'):
                c[TERM['contact_vel']] = w['contact_vel']
            if on('contact_height'):
                c[TERM['contact_height']] = w['contact_height']
            if on('floor_reg') and 'floor' in have and 'obs_floor' in have:
                c[TERM['floor_reg']] = w['floor_reg'] * nsteps
        return c

    def evaluate(self, observed, pred, nsteps, mode, init_motion_scale=1.0):
        """mode: 'root' (stage I), 'smpl' (stage II), 'motion' (stage III).

        pred (camera frame unless noted):
          Jtr (B,T,njx,3), verts3d (B,T,43,3), betas (B,16), [latent_pose (B,T,32)], [floor_plane (B,3)]
          motion only: prior_joints3d (B,T,22,3), joints3d_rollout (B,T,22,3), contacts_logits (B,T-1,9),
          latent_motion (B,T-1,48), prior_out (T-1,B,96)|None, joints_vel/trans_vel/root_orient_vel (B,1,·)
        Returns (loss, stats_dict) like FittingLoss.motion_fit / smpl_fit / root_fit.
        """
        Jtr = pred['Jtr']
        B, T, njx = Jtr.shape[:3]
        dev = Jtr.device
        have = set()
        for k in ('joints2d', 'joints3d', 'verts3d', 'seq_interval'):
            if k in observed:
                have.add(k)
        if 'floor_plane' in observed:
            have.add('obs_floor')
        if mode in ('smpl', 'motion'):
            have.add('smpl')
        if mode == 'motion':
            have.add('motion')
        if pred.get('floor_plane') is not None:
            have.add('floor')
        if pred.get('latent_pose') is not None:
            have.add('latent_pose')
        if 'joints2d' in have and not self.can_reproj and self.loss_weights['joints2d'] > 0:
            raise RuntimeError('Must provide camera intrinsics to use the re-projection loss')
        coef = self.coefficients(nsteps, init_motion_scale, have)
        i32 = lambda t: t.to(device=dev, dtype=torch.int32).contiguous()
        T_obs = None
        cfg = {'B': B, 'T': T, 'njx': njx, 'coef': coef, 'sigma2d': self.joints2d_sigma,
               'assume_unit_grad': self.assume_unit_grad}
        for k, name in (('joints2d', 'obs_joints2d'), ('joints3d', 'obs_joints3d'), ('verts3d', 'obs_verts3d')):
            if k in observed:
                o = _ext.f32c(observed[k])
                cfg[name] = o
                T_obs = o.shape[1] if T_obs is None else T_obs
                if o.shape[1] != T_obs or o.shape[1] < T:
                    raise ValueError('observation tensors must share their frame count and cover the prediction')
        cfg['T_obs'] = T if T_obs is None else T_obs
        if 'floor_plane' in observed:
            cfg['obs_floor'] = _ext.f32c(observed['floor_plane'])
        if 'seq_interval' in observed:
            cfg['seq_interval'] = i32(observed['seq_interval'])
        if self.can_reproj:
            cfg['cam_f'] = _ext.f32c(self.cam_f.reshape(-1, 2).expand(B, 2))
            cfg['cam_c'] = _ext.f32c(self.cam_cent.reshape(-1, 2).expand(B, 2))
        cfg['_keep'] = [v for v in cfg.values() if torch.is_tensor(v)]
        motion = mode == 'motion'
        betas16 = pred['betas']
        if betas16.shape[-1] > 16:
            raise ValueError('the fused energy kernel holds at most 16 shape coefficients')
        if betas16.shape[-1] < 16:          # the kernel reads / writes rows of 16 (csrc/losses.cu); autograd slices the gradient back
            betas16 = torch.nn.functional.pad(betas16, (0, 16 - betas16.shape[-1]))
        zdummy = self._dummy('z', (B, max(T - 1, 1), 48), dev)
        inp = [
            Jtr, pred['verts3d'],
            pred['prior_joints3d'] if motion else Jtr[:, :, :22],
            pred['joints3d_rollout'] if motion else Jtr[:, :, :22],
            pred['contacts_logits'] if motion else self._dummy('cl', (B, max(T - 1, 1), 9), dev),
            betas16, pred.get('floor_plane'),
            pred['latent_motion'] if motion else zdummy,
            pred.get('prior_out') if motion else None,
            pred.get('latent_pose') if coef[TERM['pose_prior']] != 0.0 else None,
        ]
        loss, terms = _FitFn.apply(cfg, *inp)
        stats = {}
        if motion and self.gmm is not None and self.loss_weights['init_motion_prior'] > 0.0:
            x = torch.cat([pred['prior_joints3d'][:, 0].reshape(B, 66), pred['joints_vel'].reshape(B, 66),
                           pred['trans_vel'].reshape(B, 3), pred['root_orient_vel'].reshape(B, 3)], 1)
            nll = _GmmFn.apply(self.gmm, x).sum()
            loss = loss + self.loss_weights['init_motion_prior'] * init_motion_scale * nll
            stats['init_motion_prior'] = nll
        if 'points3d' in observed and pred.get('points3d') is not None and self.loss_weights.get('points3d', 0.0) > 0.0:
            # fitting_loss.py:114-117 (every stage: root_fit / smpl_fit / motion_fit all go through it)
            cur = self.points3d_loss(observed['points3d'][:, :T], pred['points3d'])
            loss = loss + self.loss_weights['points3d'] * cur
            stats['points3d'] = cur
        for name, i in TERM.items():
            if coef[i] != 0.0:
                stats[name] = terms[i]
        if 'prev_batch_overlap_res' in observed and 'seq_interval' in observed and self.loss_weights['rgb_overlap_consist'] > 0.0:
            loss, stats = self._xbatch_terms(loss, stats, observed, pred, mode, T)
        return loss, stats

    def _xbatch_terms(self, loss, stats, observed, pred, mode, T):
        """Second and later batches of a split video: the FIRST sequence is tied to the cached result of the previous batch's
        LAST one - key vertices over the shared frames (positions + finite-difference velocities) in every stage, betas from
        Stage II, the floor in Stage III (fitting_loss.py:159-179, :216-222, :301-307; cache written by run_fitting.py:428-435).
        The previous result is constant: B-independent work on <= overlap x 43 points, plain torch ops.  The shared-frame count
        is read on the host ONCE per (previous result, interval tensor) pair - never inside a captured closure."""
        prev = observed['prev_batch_overlap_res']
        dev = pred['verts3d'].device
        iv = observed['seq_interval']
        key = (prev['seq_interval'].data_ptr(), iv.data_ptr())
        if getattr(self, '_xbatch_key', None) != key:
            self._xbatch_ov = int(prev['seq_interval'].reshape(-1)[1].item()) - int(iv[0, 0].item())
            self._xbatch_key = key
        cur_ov = self._xbatch_ov
        w = self.loss_weights['rgb_overlap_consist']
        zero = pred['verts3d'].sum() * 0.0
        pos, vel = zero, zero
        ov_len = min(T, cur_ov)
        if ov_len > 0:
            a = _ext.f32c(prev['verts3d']).to(dev)[-cur_ov:][:ov_len]
            c = pred['verts3d'][0, :ov_len]
            d = torch.where(torch.isinf(a), torch.zeros_like(c), a - c)          # verts3d_loss: invisible (inf) entries carry no energy
            pos = 0.5 * (d ** 2).sum()
            if cur_ov > 1 and ov_len > 1:
                da, dc = a[1:] - a[:-1], c[1:] - c[:-1]
                dv = torch.where(torch.isinf(da) | torch.isnan(da), torch.zeros_like(dc), da - dc)
                vel = 0.5 * (dv ** 2).sum()
        stats['rgb_overlap_xbatch_verts3d_pos'], stats['rgb_overlap_xbatch_verts3d_vel'] = pos, vel
        loss = loss + w * (pos + vel)
        if mode in ('smpl', 'motion'):
            bet = 0.5 * ((pred['betas'][0] - _ext.f32c(prev['betas']).to(dev).reshape(-1)) ** 2).sum()
            stats['rgb_overlap_xbatch_betas'] = bet
            loss = loss + w * bet
        if mode == 'motion' and pred.get('floor_plane') is not None:
            o = _ext.f32c(prev['floor_plane']).to(dev).reshape(-1)                  # 4-parameter plane (parse_floor_plane)
            fl = 0.5 * ((pred['floor_plane'][0] - o[:3] * o[3:]) ** 2).sum()
            stats['rgb_overlap_xbatch_floor'] = fl
            loss = loss + w * fl
        return loss, stats

    def points3d_loss(self, points3d_obs, points3d_pred):
        """One-way chamfer of the observed cloud against the predicted vertices with robust (bisquare/MAD) weights
        (fitting_loss.py:378-396).  Nearest neighbours + deterministic scatter gradient: csrc/chamfer.cu."""
        if self.chamfer_dist is None:
            raise RuntimeError('FittingLoss was built with use_chamfer=False but a points3d term is active')
        B, T, N_obs, _ = points3d_obs.size()
        obs = points3d_obs.reshape(B * T, N_obs, 3)
        pred = points3d_pred.reshape(B * T, -1, 3)
        obs2pred_sqr_dist, _ = self.chamfer_dist(obs, pred)
        weighted, _ = apply_robust_weighting_sq(obs2pred_sqr_dist.reshape(B, T * N_obs), self.robust_loss,
                                                self.robust_tuning_const)
        return 0.5 * torch.sum(weighted)

    # reference-named entry points ------------------------------------------------------------------
    def root_fit(self, observed_data, pred_data):
        return self.evaluate(observed_data, pred_data, pred_data['Jtr'].shape[1], 'root')

    def smpl_fit(self, observed_data, pred_data, nsteps):
        return self.evaluate(observed_data, pred_data, nsteps, 'smpl')

    def motion_fit(self, observed_data, pred_data, cam_pred_data, nsteps, cond_prior=None, init_motion_scale=1.0):
        p = dict(cam_pred_data)
        for k in ('prior_joints3d', 'joints3d_rollout', 'contacts_logits', 'latent_motion', 'prior_out',
                  'joints_vel', 'trans_vel', 'root_orient_vel'):
            if k in pred_data:
                p[k] = pred_data[k]
        return self.evaluate(observed_data, p, nsteps, 'motion', init_motion_scale)
