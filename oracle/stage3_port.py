"""ORACLE (test infrastructure) — plain-torch restatement of the reference's Stage-III closure.

Portable twin of the reference code so that parity can be checked on the GPU box (which has no
/root/reference) and timed as the CPU baseline (``cpu_baseline.kind = "port"``).  Pinned against
the reference executed in the build container by tests/test_oracle_vs_reference.py and by the
fixtures of oracle/make_golden.py.  Gradients come from torch autograd.

Each function cites the reference lines it follows (paths relative to /root/reference/humor).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from oracle.smplh_lbs import rodrigues, SMPLHOracle

# body_model/utils.py:5-19, datasets/amass_utils.py:22-23, fitting/fitting_utils.py:678-680
SMPL_PARENTS_REF = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 12, 12, 13, 14, 16, 17, 18, 19]
KEYPT_VERTS = [4404, 920, 3076, 3169, 823, 4310, 1010, 1085, 4495, 4569, 6615, 3217, 3313, 6713,
               6785, 3383, 6607, 3207, 1241, 1508, 4797, 4122, 1618, 1569, 5135, 5040, 5691, 5636,
               5404, 2230, 2173, 2108, 134, 3645, 6543, 3123, 3024, 4194, 1306, 182, 3694, 4294, 744]
CONTACT_INDS = [0, 4, 5, 7, 8, 10, 11, 20, 21]
SMPL2OP = [52, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62]
OP_IGNORE = [1, 9, 12]
CONTACT_HEIGHT_THRESH = 0.08


# ------------------------------------------------------------------------------------------------
# rotations  (utils/transforms.py)
# ------------------------------------------------------------------------------------------------
def mat2aa(R):
    """rotation_matrix_to_angle_axis, transforms.py:243-389 (quaternion route, NaN -> 0)."""
    R = R.reshape(-1, 3, 3)
    m = R.transpose(1, 2)                      # the reference works on the transposed matrix
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    d2 = m22 < 1e-6
    d01 = m00 > m11
    d0n1 = m00 < -m11
    t0 = 1 + m00 - m11 - m22
    t1 = 1 - m00 + m11 - m22
    t2 = 1 - m00 - m11 + m22
    t3 = 1 + m00 + m11 + m22
    q0 = torch.stack([m[:, 1, 2] - m[:, 2, 1], t0, m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], -1)
    q1 = torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t1, m[:, 1, 2] + m[:, 2, 1]], -1)
    q2 = torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t2], -1)
    q3 = torch.stack([t3, m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], -1)
    c0 = (d2 & d01).to(R.dtype)[:, None]
    c1 = (d2 & ~d01).to(R.dtype)[:, None]
    c2 = (~d2 & d0n1).to(R.dtype)[:, None]
    c3 = (~d2 & ~d0n1).to(R.dtype)[:, None]
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0[:, None] * c0 + t1[:, None] * c1 + t2[:, None] * c2 + t3[:, None] * c3)
    q = q * 0.5
    # quaternion_to_angle_axis, transforms.py:345-389
    s2 = (q[:, 1:] ** 2).sum(-1)
    s = torch.sqrt(s2)
    c = q[:, 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    aa = q[:, 1:] * k[:, None]
    return torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)


def world2aligned(R):
    """compute_world2aligned_mat / compute_aligned_from_right, transforms.py:17-42."""
    right = -R[:, :, 0]
    xproj = right[:, 0:1] / (torch.norm(right[:, :2], dim=1, keepdim=True) + 1e-6)
    ang = torch.acos(torch.clamp(xproj, -1.0, 1.0))
    r_xy = right * torch.tensor([1.0, 1.0, 0.0], dtype=R.dtype, device=R.device)
    xaxis = torch.tensor([[1.0, 0.0, 0.0]], dtype=R.dtype, device=R.device).expand_as(r_xy)
    axis = torch.cross(r_xy, xaxis, dim=1)
    aa = axis / (torch.norm(axis, dim=1, keepdim=True) + 1e-6) * ang
    return rodrigues(aa)


# ------------------------------------------------------------------------------------------------
# HuMoR CVAE  (models/humor_model.py)
# ------------------------------------------------------------------------------------------------
def mlp_forward(sd, prefix, x, n_linear, skip_from=None):
    """MLP.forward, humor_model.py:1231-1241: Linear, [GroupNorm(16), ReLU, (cat skip), Linear]*."""
    skip = None if skip_from is None else x[:, skip_from:]
    idx = 0
    x = F.linear(x, sd[f'{prefix}.net.0.weight'], sd[f'{prefix}.net.0.bias'])
    for _ in range(1, n_linear):
        idx += 1
        x = F.group_norm(x, 16, sd[f'{prefix}.net.{idx}.weight'], sd[f'{prefix}.net.{idx}.bias'], 1e-5)
        x = F.relu(x)
        idx += 2
        if skip is not None:
            x = torch.cat([x, skip], 1)
        x = F.linear(x, sd[f'{prefix}.net.{idx}.weight'], sd[f'{prefix}.net.{idx}.bias'])
    return x


def prior_net(sd, past_in):
    """HumorModel.prior, humor_model.py:407-418."""
    o = mlp_forward(sd, 'prior_net', past_in, 5)
    return o[:, :48], torch.exp(o[:, 48:])


def decode(sd, z, past_in):
    """HumorModel.decode with output_delta, in 'mat' / out 'aa', humor_model.py:445-498.
    past_in (B,339) = [trans3 tvel3 R0 9 rvel3 pose 189 joints66 jvel66]; returns (B,348)."""
    B = z.shape[0]
    raw = mlp_forward(sd, 'decoder', torch.cat([past_in, z], 1), 4, skip_from=339)   # (B,216)
    i, o = past_in, raw
    R_in = torch.cat([i[:, 6:15], i[:, 18:207]], 1).reshape(B * 22, 3, 3)
    d_aa = torch.cat([o[:, 6:9], o[:, 12:75]], 1).reshape(B * 22, 3)
    R_out = torch.bmm(rodrigues(d_aa), R_in).reshape(B, 22 * 9)
    return torch.cat([o[:, 0:3] + i[:, 0:3], o[:, 3:6] + i[:, 3:6], R_out[:, :9],
                      o[:, 9:12] + i[:, 15:18], R_out[:, 9:], o[:, 75:141] + i[:, 207:273],
                      o[:, 141:207] + i[:, 273:339], o[:, 207:216]], 1)


def _split348(x):
    return {'trans': x[:, 0:3], 'trans_vel': x[:, 3:6], 'root_orient': x[:, 6:15],
            'root_orient_vel': x[:, 15:18], 'pose_body': x[:, 18:207], 'joints': x[:, 207:273],
            'joints_vel': x[:, 273:339], 'contacts': x[:, 339:348]}


def _rigid(d, Rm, t, t2j, invert):
    """apply_world2local_trans on one step's dict of (B,d) tensors, humor_model.py:696-772."""
    B = Rm.shape[0]
    M = Rm.transpose(1, 2) if invert else Rm
    rot = lambda v: torch.einsum('bij,bnj->bni', M, v.reshape(B, -1, 3))
    out = dict(d)
    out['root_orient'] = torch.bmm(M, d['root_orient'].reshape(B, 3, 3)).reshape(B, 9)
    for k in ('trans_vel', 'root_orient_vel', 'joints_vel'):
        out[k] = rot(d[k]).reshape(B, -1)
    if invert:
        out['trans'] = rot(d['trans'])[:, 0] - t
        out['joints'] = (rot(d['joints'].reshape(B, 22, 3) + t2j[:, None]) - t2j[:, None] - t[:, None]).reshape(B, 66)
    else:
        out['trans'] = rot(d['trans'] + t)[:, 0]
        out['joints'] = (rot(d['joints'].reshape(B, 22, 3) + t[:, None] + t2j[:, None]) - t2j[:, None]).reshape(B, 66)
    return out


def roll_out(sd, init, z_seq):
    """HumorModel.roll_out(x_past=None, z_seq=..., return_prior=True), humor_model.py:785-1017.
    init: dict of (B,d) tensors in the local frame, rotations as matrices.
    Returns world-frame dict of (B,S,d) and (pm, pv) each (B,S,48)."""
    names = ['trans', 'trans_vel', 'root_orient', 'root_orient_vel', 'pose_body', 'joints', 'joints_vel']
    B, S = z_seq.shape[0], z_seq.shape[1]
    past = torch.cat([init[k] for k in names], 1)
    zero = torch.zeros(B, 1, dtype=past.dtype, device=past.device)
    t2j = -torch.cat([init['joints'][:, :2], zero], 1)
    Gr = torch.eye(3, dtype=past.dtype, device=past.device)[None].expand(B, 3, 3)
    Gt = torch.zeros(B, 3, dtype=past.dtype, device=past.device)
    world, pms, pvs = [], [], []
    for t in range(S):
        pm, pv = prior_net(sd, past)
        pms.append(pm)
        pvs.append(pv)
        x = _split348(decode(sd, z_seq[:, t], past))
        Ra = world2aligned(x['root_orient'].reshape(B, 3, 3))
        ta = torch.cat([-x['trans'][:, :2], zero], 1)
        nxt = _rigid(x, Ra, ta, t2j, invert=False)
        w = _rigid(x, Gr, Gt, t2j, invert=True)
        Gt = torch.cat([-w['trans'][:, :2], zero], 1)
        Gr = torch.bmm(Gr, Ra)
        world.append(w)
        past = torch.cat([nxt[k] for k in names], 1)
    out = {k: torch.stack([w[k] for w in world], 1) for k in world[0]}
    return out, (torch.stack(pms, 1), torch.stack(pvs, 1))


# ------------------------------------------------------------------------------------------------
# camera <-> prior frame  (fitting/fitting_utils.py, fitting/motion_optimizer.py)
# ------------------------------------------------------------------------------------------------
def parse_floor_plane(fp3):
    """fitting_utils.py:88-103."""
    off = torch.norm(fp3, dim=1, keepdim=True)
    n = fp3 / off
    neg = n[:, 1:2] > 0.0
    n = torch.where(neg.expand_as(n), -n, n)
    off = torch.where(neg, -off, off)
    return torch.cat([n, off], 1)


def _plane_hit(p, d, plane):
    """compute_plane_intersection, fitting_utils.py:61-77."""
    s = (plane[:, 3] - (plane[:, :3] * p).sum(-1)) / (plane[:, :3] * d).sum(-1)
    return p + s[:, None] * d, s


def compute_cam2prior(floor3, trans, root_orient, joints):
    """fitting_utils.py:149-190."""
    plane = parse_floor_plane(floor3) if floor3.shape[1] == 3 else floor3
    n = plane[:, :3]
    floor_trans, _ = _plane_hit(trans, -n, plane)
    right_body = -rodrigues(root_orient)[:, :, 0]
    hit, s = _plane_hit(trans, right_body, plane)
    right = hit - floor_trans
    right = torch.where(s[:, None] < 0, -right, right)
    right = right / torch.norm(right, dim=1, keepdim=True)
    fwd = torch.cross(n, right, dim=1)
    fwd = fwd / torch.norm(fwd, dim=1, keepdim=True)
    R = torch.stack([right, fwd, n], 2).transpose(2, 1)
    _, s_root = _plane_hit(joints[:, 0], -n, plane)
    return R, -trans, s_root[:, None]


class Stage3Port:
    """The Stage-III closure (motion_optimizer.py:514-608) as one forward function."""

    def __init__(self, asset, humor_sd, gmm, vposer, weights, B, T, optim_floor, cam_mat=None,
                 joints2d_sigma=100.0, dtype=torch.float32, device='cpu'):
        self.smpl = SMPLHOracle(asset, 16, dtype, device)
        self.sd = {k: v.to(dtype=dtype, device=device) for k, v in humor_sd.items()}
        self.vposer = vposer
        self.w = dict(weights)
        self.B, self.T, self.optim_floor = B, T, optim_floor
        self.sigma = joints2d_sigma
        self.robust_loss, self.robust_tuning_const = 'bisquare', 4.6851
        gw, gm, gc = [g.to(dtype=dtype, device=device) for g in gmm]
        self.gmm_logw = torch.log(gw / gw.sum())
        self.gmm_mean = gm
        self.gmm_chol = torch.linalg.cholesky(gc)
        if cam_mat is not None:
            cm = torch.as_tensor(cam_mat, dtype=dtype, device=device)
            self.cam_f = torch.stack([cm[:, 0, 0], cm[:, 1, 1]], 1)
            self.cam_c = torch.stack([cm[:, 0, 2], cm[:, 1, 2]], 1)

    # -- SMPL ------------------------------------------------------------------------------------
    def smpl_results(self, trans, root_orient, body_pose, betas):
        """motion_optimizer.py:1065-1110 (the expand/pad to B*T rows changes cost, not values)."""
        B, T = trans.shape[:2]
        bt = betas[:, None].expand(B, T, betas.shape[1]).reshape(B * T, -1)
        v, J, _ = self.smpl.forward(bt, root_orient.reshape(B * T, 3), body_pose.reshape(B * T, 63),
                                    trans.reshape(B * T, 3))
        J = J.reshape(B, T, -1, 3)
        v = v.reshape(B, T, -1, 3)
        return {'joints3d': J[:, :, :22], 'joints3d_extra': J[:, :, 22:], 'points3d': v,
                'verts3d': v[:, :, KEYPT_VERTS]}

    def latent2pose(self, z):
        """motion_optimizer.py:1041-1051."""
        B, T = z.shape[:2]
        R = self.vposer.decode(z.reshape(B * T, -1), output_type='matrot')
        return mat2aa(R.reshape(-1, 3, 3)).reshape(B, T, 63)

    def apply_cam2prior(self, trans, root_orient, R, t, h, body_pose, betas, inverse=False):
        """motion_optimizer.py:678-741 with key_frame_idx = 0."""
        B, T = trans.shape[:2]
        Rm = rodrigues(root_orient.reshape(-1, 3)).reshape(B, T, 3, 3)
        Rt = R[:, None].expand(B, T, 3, 3)
        ro = mat2aa(torch.matmul(Rt.transpose(3, 2) if inverse else Rt, Rm).reshape(-1, 3, 3)).reshape(B, T, 3)
        if inverse:
            tr = trans - trans[:, 0:1]
            tr = torch.matmul(Rt.transpose(3, 2), tr[..., None])[..., 0] - t[:, None]
        else:
            tr = torch.matmul(Rt, (trans + t[:, None])[..., None])[..., 0]
            cur_h = self.smpl_results(tr, ro, body_pose, betas)['joints3d'][:, 0, 0, 2:3]
            off = torch.cat([torch.zeros(B, 2, dtype=tr.dtype, device=tr.device), h - cur_h], 1)
            tr = tr + off[:, None]
        return tr, ro

    def rollout_latent_motion(self, p, body_pose0, z, cam2prior):
        """motion_optimizer.py:876-1019."""
        B = z.shape[0]
        trans, root_orient = p['trans'], p['root_orient']
        if self.optim_floor:
            trans, root_orient = self.apply_cam2prior(trans, root_orient, *cam2prior, body_pose0, p['betas'])
        joints = self.smpl_results(trans, root_orient, body_pose0, p['betas'])['joints3d']
        init = {'trans': trans[:, 0], 'trans_vel': p['trans_vel'][:, 0],
                'root_orient': rodrigues(root_orient.reshape(-1, 3)).reshape(B, 9),
                'root_orient_vel': p['root_orient_vel'][:, 0],
                'pose_body': rodrigues(body_pose0.reshape(-1, 3)).reshape(B, 189),
                'joints': joints.reshape(B, 66), 'joints_vel': p['joints_vel'][:, 0]}
        pred, prior = roll_out(self.sd, init, z)
        S = z.shape[1]
        out = {
            'trans': torch.cat([trans, pred['trans']], 1),
            'root_orient': torch.cat([root_orient, mat2aa(pred['root_orient'].reshape(-1, 3, 3)).reshape(B, S, 3)], 1),
            'pose_body': torch.cat([body_pose0, mat2aa(pred['pose_body'].reshape(-1, 3, 3)).reshape(B, S, 63)], 1),
            'joints': torch.cat([joints, pred['joints'].reshape(B, S, 22, 3)], 1),
            'cond_prior': prior, 'contacts_logits': pred['contacts'], 'raw': pred,
        }
        conf9 = torch.sigmoid(pred['contacts'])
        conf = torch.zeros(B, S, 22, dtype=conf9.dtype, device=conf9.device)
        conf[:, :, CONTACT_INDS] = conf[:, :, CONTACT_INDS] + conf9
        out['contacts_conf'] = torch.cat([conf[:, 0:1], conf], 1)
        out['contacts'] = (out['contacts_conf'] > 0.5).to(conf.dtype)
        if self.optim_floor:
            ct, cr = self.apply_cam2prior(out['trans'], out['root_orient'], *cam2prior, None, None, inverse=True)
        else:
            ct, cr = out['trans'], out['root_orient']
        return out, {'trans': ct, 'root_orient': cr, 'pose_body': out['pose_body']}

    # -- losses (fitting/fitting_loss.py) --------------------------------------------------------
    @staticmethod
    def _l2_vis(obs, pred):
        """joints3d_loss / verts3d_loss, fitting_loss.py:360-376."""
        vis = ~torch.isinf(obs)
        return 0.5 * ((obs[vis] - pred[vis]) ** 2).sum()

    def joints2d_loss(self, obs, j3d, j3d_extra):
        """fitting_loss.py:317-358 + perspective_projection fitting_utils.py:647-676 (R=I,t=0) + gmof :250-258."""
        B, T = obs.shape[:2]
        full = torch.cat([j3d, j3d_extra], 2)[:, :, SMPL2OP]
        proj = full[..., :2] / full[..., 2:3]
        px = proj * self.cam_f[:, None, None, :] + self.cam_c[:, None, None, :]
        conf = obs[..., 2:3].clone()
        conf[:, :, OP_IGNORE] = 0.0
        r2 = (px - obs[..., :2]) ** 2
        s2 = self.sigma ** 2
        return ((conf ** 2) * (s2 * r2) / (s2 + r2)).sum()

    def gmm_nll(self, x):
        """init_motion_prior_loss, fitting_loss.py:416-429 (MixtureSameFamily log_prob)."""
        d = x[:, None, :] - self.gmm_mean[None]                                   # (B,K,D)
        Lk = self.gmm_chol[None].expand(x.shape[0], -1, -1, -1)
        y = torch.linalg.solve_triangular(Lk, d[..., None], upper=False)[..., 0]
        maha = (y ** 2).sum(-1)
        logdet = torch.log(torch.diagonal(self.gmm_chol, dim1=-2, dim2=-1)).sum(-1)
        logp = -0.5 * (x.shape[1] * math.log(2 * math.pi) + maha) - logdet[None]
        return -torch.logsumexp(self.gmm_logw[None] + logp, 1).sum()

    def motion_fit(self, obs, pred, cam_pred, p, z, prior, nsteps, init_scale, w, mode='motion'):
        """FittingLoss.motion_fit -> smpl_fit -> root_fit, fitting_loss.py:94-309.  mode 'root' stops after root_fit (:94-179),
        'smpl' after smpl_fit (:181-224), 'motion' runs all of motion_fit."""
        smpl, motion = mode in ('smpl', 'motion'), mode == 'motion'
        st = {}
        loss = 0.0
        if 'joints3d' in obs and w['joints3d'] > 0:
            st['joints3d'] = self._l2_vis(obs['joints3d'], cam_pred['joints3d'])
            loss = loss + w['joints3d'] * st['joints3d']
        if 'verts3d' in obs and w['verts3d'] > 0:
            st['verts3d'] = self._l2_vis(obs['verts3d'], cam_pred['verts3d'])
            loss = loss + w['verts3d'] * st['verts3d']
        if 'points3d' in obs and w.get('points3d', 0.0) > 0:
            # fitting_loss.py:114-117,378-396 (restated in oracle/chamfer.py; robust settings of run_fitting.py:398-399)
            from oracle.chamfer import points3d_loss
            st['points3d'] = points3d_loss(obs['points3d'], cam_pred['points3d'], self.robust_loss, self.robust_tuning_const)
            loss = loss + w['points3d'] * st['points3d']
        if 'joints2d' in obs and w['joints2d'] > 0:
            st['joints2d'] = self.joints2d_loss(obs['joints2d'], cam_pred['joints3d'], cam_pred['joints3d_extra'])
            loss = loss + w['joints2d'] * st['joints2d']
        ov_on = 'seq_interval' in obs and w['rgb_overlap_consist'] > 0
        if ov_on:
            iv = obs['seq_interval']
            ov = (iv[:-1, 1] - iv[1:, 0]).tolist()
            v = cam_pred['verts3d']
            pos = vel = 0.0
            for b in range(1, v.shape[0]):
                o = int(ov[b - 1])
                a, c = v[b - 1, -o:], v[b, :o]
                pos = pos + 0.5 * ((a - c) ** 2).sum()
                if o > 1:
                    vel = vel + 0.5 * (((a[1:] - a[:-1]) - (c[1:] - c[:-1])) ** 2).sum()
            st['rgb_overlap_consist_verts3d_pos'], st['rgb_overlap_consist_verts3d_vel'] = pos, vel
            loss = loss + w['rgb_overlap_consist'] * (pos + vel)
            if 'prev_batch_overlap_res' in obs:            # fitting_loss.py:159-179: first sequence vs the previous batch's last one
                prev = obs['prev_batch_overlap_res']
                cur_ov = int(prev['seq_interval'][1] - iv[0, 0])
                ov_len = min(v.shape[1], cur_ov)
                a, c = prev['verts3d'][-cur_ov:][:ov_len], v[0, :ov_len]
                vis = ~torch.isinf(a)
                xpos = 0.5 * ((a[vis] - c[vis]) ** 2).sum()
                xvel = 0.0
                if cur_ov > 1:
                    da, dc = a[1:] - a[:-1], c[1:] - c[:-1]
                    vv = ~torch.isinf(da)
                    xvel = 0.5 * ((da[vv] - dc[vv]) ** 2).sum()
                st['rgb_overlap_xbatch_verts3d_pos'], st['rgb_overlap_xbatch_verts3d_vel'] = xpos, xvel
                loss = loss + w['rgb_overlap_consist'] * (xpos + xvel)
        if smpl and w['pose_prior'] > 0:
            st['pose_prior'] = (cam_pred['latent_pose'] ** 2).sum()
            loss = loss + w['pose_prior'] * st['pose_prior']
        if smpl and w['shape_prior'] > 0:
            st['shape_prior'] = (p['betas'] ** 2).sum()
            loss = loss + w['shape_prior'] * nsteps * st['shape_prior']
        if smpl and w['joints3d_smooth'] > 0:
            j = cam_pred['joints3d']
            st['joints3d_smooth'] = 0.5 * ((j[:, 1:] - j[:, :-1]) ** 2).sum()
            loss = loss + w['joints3d_smooth'] * st['joints3d_smooth']
        if smpl and ov_on:
            st['rgb_overlap_consist_betas'] = 0.5 * ((p['betas'][:-1] - p['betas'][1:]) ** 2).sum()
            loss = loss + w['rgb_overlap_consist'] * st['rgb_overlap_consist_betas']
            if 'prev_batch_overlap_res' in obs:            # fitting_loss.py:217-222
                st['rgb_overlap_xbatch_betas'] = 0.5 * ((p['betas'][0] - obs['prev_batch_overlap_res']['betas']) ** 2).sum()
                loss = loss + w['rgb_overlap_consist'] * st['rgb_overlap_xbatch_betas']
        if not motion:
            return loss, st
        if w['motion_prior'] > 0:
            pm, pv = prior
            lp = -torch.log(torch.sqrt(pv)) - math.log(math.sqrt(2 * math.pi)) - (z - pm) ** 2 / (2 * pv)
            st['motion_prior'] = -lp.sum()
            loss = loss + w['motion_prior'] * st['motion_prior']
        if w['init_motion_prior'] > 0:
            x = torch.cat([pred['joints3d'][:, 0].reshape(-1, 66), p['joints_vel'].reshape(-1, 66),
                           p['trans_vel'].reshape(-1, 3), p['root_orient_vel'].reshape(-1, 3)], 1)
            st['init_motion_prior'] = self.gmm_nll(x)
            loss = loss + w['init_motion_prior'] * init_scale * st['init_motion_prior']
        if w['joint_consistency'] > 0:
            st['joint_consistency'] = 0.5 * ((pred['joints3d'] - pred['joints3d_rollout']) ** 2).sum()
            loss = loss + w['joint_consistency'] * st['joint_consistency']
        if w['bone_length'] > 0:
            jr = pred['joints3d_rollout']
            bl = torch.norm(jr[:, :, 1:] - jr[:, :, SMPL_PARENTS_REF[1:]], dim=-1)
            st['bone_length'] = 0.5 * ((bl[:, 1:] - bl[:, :-1]) ** 2).sum()
            loss = loss + w['bone_length'] * st['bone_length']
        if 'joints3d' in obs and w['joints3d_rollout'] > 0:
            st['joints3d_rollout'] = self._l2_vis(obs['joints3d'], pred['joints3d_rollout'])
            loss = loss + w['joints3d_rollout'] * st['joints3d_rollout']
        if w['contact_vel'] > 0:
            j = pred['joints3d']
            st['contact_vel'] = 0.5 * (((j[:, 1:] - j[:, :-1]) ** 2).sum(-1) * pred['contacts_conf'][:, 1:]).sum()
            loss = loss + w['contact_vel'] * st['contact_vel']
        if w['contact_height'] > 0:
            st['contact_height'] = (F.relu(pred['joints3d'][..., 2].abs() - CONTACT_HEIGHT_THRESH) * pred['contacts_conf']).sum()
            loss = loss + w['contact_height'] * st['contact_height']
        if self.optim_floor and w['floor_reg'] > 0 and 'floor_plane' in obs:
            o = obs['floor_plane']
            st['floor_reg'] = 0.5 * ((p['floor_plane'] - o[:, :3] * o[:, 3:]) ** 2).sum()
            loss = loss + w['floor_reg'] * nsteps * st['floor_reg']
        if self.optim_floor and ov_on:
            fp = p['floor_plane']
            st['rgb_overlap_consist_floor'] = 0.5 * ((fp[:-1] - fp[1:]) ** 2).sum()
            loss = loss + w['rgb_overlap_consist'] * st['rgb_overlap_consist_floor']
            if 'prev_batch_overlap_res' in obs:            # fitting_loss.py:302-307: floor_reg_loss against the previous 4-parameter plane
                o = obs['prev_batch_overlap_res']['floor_plane']
                st['rgb_overlap_xbatch_floor'] = 0.5 * ((fp[0] - o[:3] * o[3:]) ** 2).sum()
                loss = loss + w['rgb_overlap_consist'] * st['rgb_overlap_xbatch_floor']
        return loss, st

    # -- Stage I / II closures ---------------------------------------------------------------------
    def closure12(self, p, obs, stage, weights):
        """motion_optimizer.py:237-250 (stage 0, root_fit) / :289-304 (stage 1, smpl_fit).  p: trans (B,T,3), root_orient
        (B,T,3), betas (B,16), latent_pose (B,T,32) leaf tensors; weights: that stage's column of the config."""
        body_pose = self.latent2pose(p['latent_pose'])
        pred = self.smpl_results(p['trans'], p['root_orient'], body_pose, p['betas'])
        pred['latent_pose'] = p['latent_pose']
        loss, st = self.motion_fit(obs, pred, pred, p, None, None, self.T, 1.0, dict(weights), mode='root' if stage == 0 else 'smpl')
        return loss, st, {'pred': pred, 'body_pose': body_pose}

    # -- the closure -----------------------------------------------------------------------------
    def closure(self, p, obs, nsteps=None, init_motion_scale=1.0):
        """p: dict of leaf tensors (stage-3 variables).  Returns loss, stats, intermediates."""
        body_pose0 = self.latent2pose(p['latent_pose'])
        cam2prior = None
        if self.optim_floor:
            j0 = self.smpl_results(p['trans'], p['root_orient'], body_pose0, p['betas'])['joints3d'][:, 0]
            cam2prior = compute_cam2prior(p['floor_plane'], p['trans'][:, 0], p['root_orient'][:, 0], j0)
        z = p['latent_motion'] if nsteps is None else p['latent_motion'][:, :nsteps - 1]
        roll, cam = self.rollout_latent_motion(p, body_pose0, z, cam2prior)
        pred = self.smpl_results(roll['trans'], roll['root_orient'], roll['pose_body'], p['betas'])
        pred['joints3d_rollout'] = roll['joints']
        pred['contacts_conf'] = roll['contacts_conf']
        cam_pred = pred
        if self.optim_floor:
            cam_pred = self.smpl_results(cam['trans'], cam['root_orient'], roll['pose_body'], p['betas'])
        if self.w['pose_prior'] > 0:
            Bz, Tz = roll['pose_body'].shape[:2]
            cam_pred['latent_pose'] = self.vposer.encode(roll['pose_body'].reshape(Bz * Tz, 63)).mean.reshape(Bz, Tz, -1)
        w = dict(self.w)
        o, n = obs, self.T
        if nsteps is not None:
            o = {k: v[:, :nsteps] for k, v in obs.items() if k != 'prev_batch_overlap_res'}     # motion_optimizer.py:590 (overlap weight is 0 in this phase)
            n = nsteps
            w['rgb_overlap_consist'] = 0.0
        loss, st = self.motion_fit(o, pred, cam_pred, p, z, roll['cond_prior'], n, init_motion_scale, w)
        return loss, st, {'rollout': roll, 'cam_rollout': cam, 'pred': pred, 'cam_pred': cam_pred,
                          'body_pose0': body_pose0, 'cam2prior': cam2prior}
