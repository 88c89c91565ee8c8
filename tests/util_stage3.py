"""Shared builders for the Stage-III parity tests: the product (CUDA) optimiser and the CPU oracle port
on the same seeded synthetic problem."""
import numpy as np
import torch

from humor_b200 import synth

PARAM_NAMES = ['trans', 'root_orient', 'latent_pose', 'betas', 'latent_motion', 'trans_vel', 'joints_vel', 'root_orient_vel']


def obs_keys(optim_floor, prob=None):
    keys = ('joints2d', 'floor_plane', 'seq_interval') if optim_floor else ('verts3d',)
    if prob is not None and 'points3d' in prob['obs']:
        keys = keys + ('points3d',)          # PROX RGB-D style problems (golden case stage3_proxd)
    if prob is not None and 'prev_batch_overlap_res' in prob['obs']:
        keys = keys + ('prev_batch_overlap_res',)          # second batch of a split video (golden case stage3_rgb_xbatch)
    return keys


def obs_to(v, device):
    if isinstance(v, dict):
        return {a: torch.as_tensor(b).to(device).clone() for a, b in v.items()}
    return torch.as_tensor(v).to(device).clone()


def project_joints2d(prob, cam_joints73, seed=5, noise=2.0):
    """Replace the random 2-D keypoints by the projection of the given camera-frame joints (+noise), keeping
    the synthetic confidences: the re-projection term then sits in its informative (non-saturated) range."""
    from oracle.stage3_port import SMPL2OP
    rng = np.random.RandomState(seed)
    j = np.asarray(cam_joints73)[:, :, SMPL2OP]                       # (B,T,25,3)
    f, c = np.asarray(synth.CAM_F), np.asarray(synth.CAM_C)
    xy = j[..., :2] / j[..., 2:3] * f + c + rng.randn(*j.shape[:3], 2) * noise
    prob['obs']['joints2d'][..., :2] = xy.astype(np.float32)
    return prob


def build_port(B, T, weights, optim_floor, prob, dtype=torch.float32, device='cpu'):
    from oracle.stage3_port import Stage3Port
    return Stage3Port(synth.make_smplh_asset(), synth.make_humor_state_dict(), synth.make_gmm(),
                      synth.FakeVPoser().to(dtype).to(device), weights, B, T, optim_floor, prob['cam_mat'], dtype=dtype,
                      device=device)


def closure_port(port, prob, optim_floor, nsteps=None, scale=1.0, device='cpu'):
    names = PARAM_NAMES + (['floor_plane'] if optim_floor else [])
    p = {k: torch.as_tensor(prob['params'][k]).to(device).clone().requires_grad_(True) for k in names}
    obs = {k: obs_to(v, device) for k, v in prob['obs'].items() if k in obs_keys(optim_floor, prob)}
    loss, stats, inter = port.closure(p, obs, nsteps, scale)
    loss.backward()
    return float(loss.detach()), {k: p[k].grad.detach() for k in names}, {'stats': {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in stats.items()}, 'inter': inter}


def build_product(B, T, weights, optim_floor, prob, device='cuda', contact_refine_only=True):
    from humor_b200.body_model import BodyModel
    from humor_b200.humor_model import HumorModel
    from humor_b200.motion_optimizer import MotionOptimizer
    dev = torch.device(device)
    bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=B * T, use_vtx_selector=optim_floor).to(dev)
    humor = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
    humor.load_state_dict(synth.make_humor_state_dict())
    humor.to(dev).eval()
    gmm = tuple(g.to(dev) for g in synth.make_gmm())
    w = dict(weights)
    mo = MotionOptimizer(dev, bm, 16, B, T, list(obs_keys(optim_floor, prob)), [dict(w), dict(w), dict(w)],
                         synth.FakeVPoser().to(dev), humor, {'gmm': gmm}, optim_floor,
                         torch.as_tensor(prob['cam_mat']).to(dev) if optim_floor else None, 'bisquare', 4.6851, 100.0,
                         stage3_contact_refine_only=contact_refine_only,
                         use_chamfer='points3d' in prob['obs'])      # run_fitting.py:405
    return mo


def closure_product(mo, prob, nsteps=None, scale=1.0):
    names = mo.set_stage3_state(prob['params'])
    obs = {k: obs_to(v, mo.device) for k, v in prob['obs'].items() if k in obs_keys(mo.optim_floor, prob)}
    loss, stats, roll, cam, cam_pred = mo.stage3_forward(obs, nsteps, scale)
    loss.backward()
    grads = {n: getattr(mo, n).grad.detach() for n in names}
    return float(loss.detach()), grads, {'stats': {k: float(v.detach()) for k, v in stats.items()}, 'roll': roll, 'cam': cam, 'cam_pred': cam_pred}
