"""Seeded synthetic stand-ins for the licensed assets the Stage-III path needs.

Nothing here is an oracle and nothing here is on the compute path: these are the
*inputs* a user of the reference would normally download (SMPL+H ``model.npz``,
HuMoR ``best_model.pth``, VPoser snapshot, ``prior_gmm.npz``) plus synthetic
observations, generated with ``numpy.random.RandomState`` so they are bit-identical
on every machine (the GPU box has no ``/root/reference`` and no network).

Shapes follow what the reference loads:
  * SMPL+H npz keys read by ``humor/body_model/body_model.py:37-48`` and by smplx:
    ``v_template (6890,3) shapedirs (6890,3,16) posedirs (6890,3,459)
    J_regressor (52,6890) weights (6890,52) kintree_table (2,52) f (13776,3)``
  * HuMoR state-dict keys ``prior_net.net.N.*`` / ``decoder.net.N.*`` / ``encoder.net.N.*``
    (``humor/models/humor_model.py:181-206,1206-1241``)
  * GMM npz ``weights (12,) means (12,138) covariances (12,138,138)``
    (``humor/fitting/run_fitting.py:251-258``)
  * VPoser contract used by ``humor/fitting/motion_optimizer.py:77,1049,1061``.
"""
import math
import numpy as np
import torch
import torch.nn as nn

NUM_VERTS = 6890
NUM_JOINTS = 52          # SMPL+H
NUM_BODY_JOINTS = 22
NUM_BETAS = 16
NUM_FACES = 13776
POSE_FEAT = 459          # 51 * 9

# true SMPL+H kinematic tree (smplx kintree_table[0]); NOT body_model/utils.py:9 SMPL_PARENTS
BODY_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]


def smplh_parents():
    par = list(BODY_PARENTS)
    for wrist in (20, 21):
        base = len(par)
        for f in range(5):
            par += [wrist, base + 3 * f, base + 3 * f + 1]
    assert len(par) == NUM_JOINTS
    return np.asarray(par, dtype=np.int64)


_REST_BODY = np.array([
    [0.00, -0.22, 0.03], [0.07, -0.31, 0.02], [-0.07, -0.31, 0.02], [0.00, -0.10, 0.00],
    [0.10, -0.70, 0.02], [-0.10, -0.70, 0.02], [0.00, 0.03, 0.02], [0.09, -1.10, -0.02],
    [-0.09, -1.10, -0.02], [0.00, 0.09, 0.02], [0.12, -1.16, 0.10], [-0.12, -1.16, 0.10],
    [0.00, 0.30, -0.01], [0.08, 0.21, 0.00], [-0.08, 0.21, 0.00], [0.00, 0.38, 0.03],
    [0.19, 0.24, -0.01], [-0.19, 0.24, -0.01], [0.45, 0.23, -0.03], [-0.45, 0.23, -0.03],
    [0.70, 0.23, -0.02], [-0.70, 0.23, -0.02]], dtype=np.float64)


def _rest_skeleton():
    par = smplh_parents()
    J = np.zeros((NUM_JOINTS, 3))
    J[:22] = _REST_BODY
    for side, wrist in ((1.0, 20), (-1.0, 21)):
        base = 22 if wrist == 20 else 37
        for f in range(5):
            for k in range(3):
                J[base + 3 * f + k] = J[wrist] + np.array(
                    [side * (0.08 + 0.03 * k), 0.01 * (f - 2), 0.015 * (f - 2)])
    return J, par


def make_smplh_asset(seed=0):
    """Synthetic SMPL+H-shaped model dict (float32 arrays, npz key names of the real asset)."""
    rng = np.random.RandomState(seed)
    J, par = _rest_skeleton()
    V = NUM_VERTS
    # every vertex hangs off a bone (child joint c, parent p)
    child = rng.randint(1, NUM_JOINTS, size=V)
    # bias towards the body: 80 % of the vertices on body bones
    body_mask = rng.rand(V) < 0.8
    child[body_mask] = rng.randint(1, NUM_BODY_JOINTS, size=int(body_mask.sum()))
    # vertex ids are region-contiguous, as in the SMPL template (an artist-made mesh whose index ranges follow body
    # parts: head, torso, arms, hands, legs ...): neighbouring ids are skinned to the same few joints
    child = np.sort(child, kind='stable')
    parent = par[child]
    u = rng.rand(V, 1)
    radial = rng.randn(V, 3) * np.where(child[:, None] >= 22, 0.008, 0.045)
    v_template = (1 - u) * J[parent] + u * J[child] + radial
    # skinning weights: <= 4 non-zeros per vertex
    W = np.zeros((V, NUM_JOINTS))
    gp = np.where(par[parent] >= 0, par[parent], parent)
    # 4th influence: a kinematic NEIGHBOUR of the bone (a child of either end, i.e. the next bone or a sibling), as in SMPL,
    # whose skinning weights are local by construction (initialised from an artist's segmentation and regularised towards
    # it).  An earlier generator drew this joint uniformly from all 52, which made every 128-vertex range touch ~48 joints.
    kids = [[c for c in range(1, NUM_JOINTS) if par[c] == j] for j in range(NUM_JOINTS)]
    pick = rng.rand(V)
    other = np.empty(V, np.int64)
    for v in range(V):
        cand = [c for c in kids[child[v]] + kids[parent[v]] if c != child[v]] or [gp[v]]
        other[v] = cand[int(pick[v] * len(cand))]
    raw = rng.dirichlet([2.0, 2.0, 0.5, 0.3], size=V)
    raw[:, 0] *= (0.3 + u[:, 0])
    raw[:, 1] *= (1.3 - u[:, 0])
    for k, idx in enumerate((child, parent, gp, other)):
        np.add.at(W, (np.arange(V), idx), raw[:, k])
    W /= W.sum(1, keepdims=True)
    # joint regressor: sparse positive rows that sum to one, supported near the joint
    Jreg = np.zeros((NUM_JOINTS, V))
    for j in range(NUM_JOINTS):
        d = np.linalg.norm(v_template - J[j], axis=1)
        nn_idx = np.argsort(d)[:24]
        w = rng.rand(24) + 0.1
        Jreg[j, nn_idx] = w / w.sum()
    shape_scale = 0.03 * (0.75 ** np.arange(NUM_BETAS))
    shapedirs = rng.randn(V, 3, NUM_BETAS) * shape_scale
    posedirs = rng.randn(V, 3, POSE_FEAT) * 0.004
    faces = rng.randint(0, V, size=(NUM_FACES, 3)).astype(np.int64)
    kintree = np.stack([par, np.arange(NUM_JOINTS)]).astype(np.int64)
    kintree[0, 0] = 4294967295  # what the real file holds for the root (uint32 -1)
    return {
        'v_template': v_template.astype(np.float32),
        'shapedirs': shapedirs.astype(np.float32),
        'posedirs': posedirs.astype(np.float32),
        'J_regressor': Jreg.astype(np.float32),
        'weights': W.astype(np.float32),
        'kintree_table': kintree,
        'f': faces,
    }


def write_smplh_npz(path, seed=0):
    np.savez(path, **make_smplh_asset(seed))
    return path


# ----------------------------------------------------------------------------------------------
# HuMoR CVAE weights
# ----------------------------------------------------------------------------------------------
def mlp_param_shapes(layers, skip):
    """(key-suffix, shape) in ModuleList order of humor_model.py:1206-1229."""
    out = []
    idx = 0
    out.append((f'net.{idx}.weight', (layers[1], layers[0])))
    out.append((f'net.{idx}.bias', (layers[1],)))
    idx += 1
    for li in range(1, len(layers) - 1):
        out.append((f'net.{idx}.weight', (layers[li],)))      # GroupNorm gamma
        out.append((f'net.{idx}.bias', (layers[li],)))        # GroupNorm beta
        idx += 2                                              # GN, ReLU
        out.append((f'net.{idx}.weight', (layers[li + 1], layers[li] + skip)))
        out.append((f'net.{idx}.bias', (layers[li + 1],)))
        idx += 1
    return out


PRIOR_LAYERS = [339, 1024, 1024, 1024, 1024, 96]
DECODER_LAYERS = [339 + 48, 1024, 1024, 512, 216]
POSTERIOR_LAYERS = [678, 1024, 1024, 1024, 1024, 96]


def make_humor_state_dict(seed=1, decoder_out_scale=0.05, logvar_scale=0.3):
    """Random-init HuMoR weights with the reference's state-dict keys (torch tensors, fp32)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for prefix, layers, skip in (('encoder', POSTERIOR_LAYERS, 0),
                                 ('decoder', DECODER_LAYERS, 48),
                                 ('prior_net', PRIOR_LAYERS, 0)):
        shapes = mlp_param_shapes(layers, skip)
        last_w = shapes[-2][0]
        for key, shp in shapes:
            if len(shp) == 2:
                bound = 1.0 / math.sqrt(shp[1])
                val = rng.uniform(-bound, bound, size=shp)
            elif key.endswith('weight'):      # GN gamma
                val = 1.0 + 0.1 * rng.randn(*shp)
            else:
                val = 0.05 * rng.randn(*shp)
            if prefix == 'decoder' and key in (last_w, last_w.replace('weight', 'bias')):
                val = val * decoder_out_scale
            if prefix != 'decoder' and key in (last_w, last_w.replace('weight', 'bias')):
                val = val.copy()
                val[48:] *= logvar_scale
            sd[f'{prefix}.{key}'] = torch.from_numpy(val.astype(np.float32))
    return sd


# ----------------------------------------------------------------------------------------------
# VPoser stand-in (duck-typed; the real one is a third-party object injected by the user)
# ----------------------------------------------------------------------------------------------
class FakeVPoser(nn.Module):
    latentD = 32

    def __init__(self, seed=2):
        super().__init__()
        rng = np.random.RandomState(seed)
        f32 = lambda a: nn.Parameter(torch.from_numpy(a.astype(np.float32)), requires_grad=False)
        self.w1 = f32(rng.randn(32, 64) / math.sqrt(32))
        self.b1 = f32(0.1 * rng.randn(64))
        self.w2 = f32(rng.randn(64, 126) * 0.25 / math.sqrt(64))
        self.b2 = f32(0.02 * rng.randn(126))
        self.we = f32(rng.randn(63, 32) / math.sqrt(63))
        self.be = f32(0.1 * rng.randn(32))
        self.base6d = f32(np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]]))

    def decode(self, z, output_type='matrot'):
        assert output_type == 'matrot'
        h = torch.tanh(z @ self.w1 + self.b1)
        o = (h @ self.w2 + self.b2).reshape(-1, 21, 3, 2)
        o = o + self.base6d
        a1, a2 = o[..., 0], o[..., 1]
        b1 = a1 / a1.norm(dim=-1, keepdim=True)
        b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
        b2 = b2 / b2.norm(dim=-1, keepdim=True)
        b3 = torch.cross(b1, b2, dim=-1)
        R = torch.stack([b1, b2, b3], dim=-1)          # (M,21,3,3) columns b1 b2 b3
        return R.reshape(-1, 1, 21, 9)

    def encode(self, pose):
        mean = pose @ self.we + self.be
        return torch.distributions.Normal(mean, torch.ones_like(mean))


# ----------------------------------------------------------------------------------------------
# Init-state GMM
# ----------------------------------------------------------------------------------------------
def make_gmm(seed=3, ncomp=12, dim=138):
    rng = np.random.RandomState(seed)
    w = rng.rand(ncomp) + 0.2
    w /= w.sum()
    means = rng.randn(ncomp, dim) * 0.3
    covs = np.zeros((ncomp, dim, dim))
    for k in range(ncomp):
        A = rng.randn(dim, dim) * 0.15
        covs[k] = A @ A.T / dim + np.diag(0.05 + 0.1 * rng.rand(dim))
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    return f(w), f(means), f(covs)


# ----------------------------------------------------------------------------------------------
# Stage-III state + observations
# ----------------------------------------------------------------------------------------------
CAM_F = (1060.531, 1060.3856)
CAM_C = (951.2999, 536.7704)
DEFAULT_FLOOR = (0.0, -1.0, 0.0, -0.5)   # rgb_dataset.py:16  (a,b,c,d)


def _aa_compose_x_pi_yaw(yaw):
    """axis-angle of  Rx(pi) @ Ry(yaw): a y-up body standing in the y-down camera frame."""
    from scipy.spatial.transform import Rotation as Rsc
    rx = Rsc.from_euler('x', np.pi + 0.08)
    out = []
    for a in yaw:
        out.append((rx * Rsc.from_euler('y', a) * Rsc.from_euler('z', 0.05)).as_rotvec())
    return np.asarray(out)


def make_stage3_problem(B, T, seed=4, overlap=10, cam=True, dtype=np.float32):
    """Synthetic Stage-III optimisation variables + observations (numpy, host).

    Returns dict(params=..., obs=..., cam_mat=...). ``params`` are the 2 961 floats/sequence
    the reference optimises in stage 3 (motion_optimizer.py:400-404); ``obs`` mirrors
    ``observed_data`` of run_fitting.py (joints2d, floor_plane, seq_interval).
    """
    rng = np.random.RandomState(seed)
    p = {}
    yaw = rng.uniform(-0.6, 0.6, size=B)
    if cam:
        p['trans'] = np.stack([rng.uniform(-0.5, 0.5, B), rng.uniform(-0.45, -0.35, B),
                               rng.uniform(3.0, 4.5, B)], 1)[:, None, :]
        p['root_orient'] = _aa_compose_x_pi_yaw(yaw)[:, None, :]
        p['floor_plane'] = (np.asarray(DEFAULT_FLOOR[:3]) * DEFAULT_FLOOR[3])[None].repeat(B, 0) \
            + 0.02 * rng.randn(B, 3)
    else:
        # canonical frame: z up, floor z=0
        from scipy.spatial.transform import Rotation as Rsc
        p['trans'] = np.stack([rng.uniform(-0.1, 0.1, B), rng.uniform(-0.1, 0.1, B),
                               rng.uniform(0.9, 1.0, B)], 1)[:, None, :]
        p['root_orient'] = np.asarray(
            [(Rsc.from_euler('z', a) * Rsc.from_euler('x', np.pi / 2)).as_rotvec() for a in yaw])[:, None, :]
    p['latent_pose'] = rng.randn(B, 1, 32) * 0.7
    p['betas'] = rng.randn(B, NUM_BETAS) * 0.5
    p['latent_motion'] = rng.randn(B, T - 1, 48) * 0.5
    p['trans_vel'] = rng.randn(B, 1, 3) * 0.1
    p['joints_vel'] = rng.randn(B, 1, 66) * 0.1
    p['root_orient_vel'] = rng.randn(B, 1, 3) * 0.1
    obs = {}
    conf = rng.uniform(0.3, 1.0, size=(B, T, 25, 1))
    conf[rng.rand(B, T, 25, 1) < 0.1] = 0.0
    xy = np.stack([rng.uniform(500, 1400, (B, T, 25)), rng.uniform(100, 1000, (B, T, 25))], -1)
    obs['joints2d'] = np.concatenate([xy, conf], -1)
    obs['floor_plane'] = np.asarray(DEFAULT_FLOOR)[None].repeat(B, 0)
    step = T - overlap
    obs['seq_interval'] = np.stack([np.arange(B) * step, np.arange(B) * step + T], 1).astype(np.int32)
    # 3-D keypoint-vertex observations with occlusion (inf) for the AMASS-style config
    v3 = rng.randn(B, T, 43, 3) * 0.3 + np.array([0, 0, 0.9])
    occl = v3[..., 2:3] < 0.6
    v3 = np.where(np.broadcast_to(occl, v3.shape), np.inf, v3)
    obs['verts3d'] = v3
    cam_mat = np.zeros((B, 3, 3))
    cam_mat[:, 0, 0], cam_mat[:, 1, 1] = CAM_F
    cam_mat[:, 0, 2], cam_mat[:, 1, 2] = CAM_C
    cam_mat[:, 2, 2] = 1.0
    cast = lambda d: {k: (v.astype(dtype) if v.dtype.kind == 'f' else v) for k, v in d.items()}
    return {'params': cast(p), 'obs': cast(obs), 'cam_mat': cam_mat.astype(dtype)}


# stage-3 column of configs/fit_rgb_demo_use_split.cfg and configs/fit_amass_keypts.cfg
RGB_STAGE3_WEIGHTS = {
    'joints2d': 0.001, 'joints3d': 0.0, 'joints3d_rollout': 0.0, 'verts3d': 0.0, 'points3d': 0.0,
    'pose_prior': 0.0, 'shape_prior': 0.05, 'motion_prior': 0.075, 'init_motion_prior': 0.075,
    'joint_consistency': 100.0, 'bone_length': 2000.0, 'joints3d_smooth': 0.0,
    'contact_vel': 100.0, 'contact_height': 10.0, 'floor_reg': 0.167, 'rgb_overlap_consist': 200.0,
}
AMASS_STAGE3_WEIGHTS = {
    'joints2d': 0.0, 'joints3d': 0.0, 'joints3d_rollout': 0.0, 'verts3d': 1.0, 'points3d': 0.0,
    'pose_prior': 0.0, 'shape_prior': 1.67e-4, 'motion_prior': 5e-4, 'init_motion_prior': 5e-4,
    'joint_consistency': 1.0, 'bone_length': 10.0, 'joints3d_smooth': 0.0,
    'contact_vel': 1.0, 'contact_height': 1.0, 'floor_reg': 0.0, 'rgb_overlap_consist': 0.0,
}
# stage-3 column of configs/fit_proxd.cfg (PROX RGB-D: point cloud + 2-D keypoints, floor optimised)
PROXD_STAGE3_WEIGHTS = {
    'joints2d': 0.001, 'joints3d': 0.0, 'joints3d_rollout': 0.0, 'verts3d': 0.0, 'points3d': 1.0,
    'pose_prior': 0.0, 'shape_prior': 0.034, 'motion_prior': 0.075, 'init_motion_prior': 0.075,
    'joint_consistency': 100.0, 'bone_length': 2000.0, 'joints3d_smooth': 0.0,
    'contact_vel': 100.0, 'contact_height': 10.0, 'floor_reg': 1.0, 'rgb_overlap_consist': 0.0,
}
WEIGHT_SETS = {'rgb': RGB_STAGE3_WEIGHTS, 'amass': AMASS_STAGE3_WEIGHTS, 'proxd': PROXD_STAGE3_WEIGHTS}


def stage12_weights(wset):
    """Stage-I/II columns (identical in the shipped configs) of fit_rgb_demo_use_split.cfg / fit_amass_keypts.cfg /
    fit_proxd.cfg: the data terms of the stage-3 column plus pose prior and temporal smoothness, no motion terms."""
    w = {k: 0.0 for k in RGB_STAGE3_WEIGHTS}
    if wset == 'rgb':
        w.update(joints2d=0.001, pose_prior=0.04, shape_prior=0.05, joints3d_smooth=100.0, rgb_overlap_consist=200.0)
    elif wset == 'amass':
        w.update(verts3d=1.0, pose_prior=2e-4, shape_prior=1.67e-4, joints3d_smooth=0.1)
    elif wset == 'proxd':
        w.update(points3d=1.0, joints2d=0.001, pose_prior=0.1, shape_prior=0.034, joints3d_smooth=100.0)
    else:
        raise ValueError(wset)
    return w


def make_stage12_params(B, T, seed=0, dtype=np.float32):
    """Stage-I/II variables (motion_optimizer.py:66-77,224-283): per-frame trans / root_orient / latent_pose, betas."""
    rng = np.random.RandomState(seed)
    trans = np.zeros((B, T, 3)) + np.array([0.0, 0.1, 3.0]) + np.cumsum(rng.randn(B, T, 3) * 0.02, 1)
    ro = np.zeros((B, T, 3)) + np.array([np.pi, 0.0, 0.0]) + rng.randn(B, T, 3) * 0.1
    return {'trans': trans.astype(dtype), 'root_orient': ro.astype(dtype), 'betas': (rng.randn(B, 16) * 0.5).astype(dtype),
            'latent_pose': (rng.randn(B, T, 32) * 0.5).astype(dtype)}


def sample_point_cloud(verts, n_obs, seed=0, noise=0.01, outlier_frac=0.05, outlier_sigma=0.6):
    """Synthetic depth-camera cloud for the points3d energy: n_obs points per frame drawn from the given vertices
    (B,T,V,3) with Gaussian noise, a fraction of them pushed far away (what the bisquare weights must reject)."""
    v = np.asarray(verts)
    B, T, V, _ = v.shape
    rng = np.random.RandomState(seed)
    ids = rng.randint(0, V, size=(B, T, n_obs))
    pts = np.take_along_axis(v, ids[..., None], axis=2) + rng.randn(B, T, n_obs, 3) * noise
    out = rng.rand(B, T, n_obs) < outlier_frac
    pts = pts + out[..., None] * rng.randn(B, T, n_obs, 3) * outlier_sigma
    return pts.astype(v.dtype)
