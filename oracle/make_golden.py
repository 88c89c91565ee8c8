"""ORACLE (test infrastructure) — generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only) on seeded synthetic inputs.

    python -m oracle.make_golden

Each fixture holds the observations that were used (2-D keypoints are the projection of the reference's
own camera-frame joints + seeded noise), the reference's loss, every energy term, the gradient of every
stage-3 variable and a few intermediates.  Inputs other than the observations are regenerated from
humor_b200.synth seeds by the tests.
"""
import os
import numpy as np
import torch

from humor_b200 import synth
from oracle import ref_closure

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
# NB: B == 3 is avoided on purpose: the reference calls torch.cross without `dim` (fitting_utils.py:181,
# transforms.py:26), which picks the FIRST size-3 axis - the batch axis when B == 3.  The reference itself pads
# a batch of 3 to 4 for this reason (run_fitting.py:61-63,288-318); port and product use dim=1.
CASES = {
    'stage3_rgb': dict(optim_floor=True, B=4, T=8, seed=21, overlap=3, nsteps=None, scale=1.0),
    'stage3_rgb_phase1': dict(optim_floor=True, B=4, T=8, seed=22, overlap=3, nsteps=4, scale=1.0),
    'stage3_rgb_refine': dict(optim_floor=True, B=2, T=7, seed=23, overlap=2, nsteps=None, scale=7.0 / 4.0),
    'stage3_amass': dict(optim_floor=False, B=2, T=6, seed=24, overlap=2, nsteps=None, scale=1.0),
    # PROX RGB-D (configs/fit_proxd.cfg): point-cloud energy through the reference's compiled chamfer module
    'stage3_proxd': dict(optim_floor=True, B=2, T=6, seed=25, overlap=2, nsteps=None, scale=1.0, wset='proxd', n_obs=96),
    # second batch of a split video: the first sequence is tied to the LAST sequence of the previous batch
    # (observed_data['prev_batch_overlap_res'], run_fitting.py:428-435 -> fitting_loss.py:159-179,216-222,301-307)
    'stage3_rgb_xbatch': dict(optim_floor=True, B=4, T=8, seed=26, overlap=3, nsteps=None, scale=1.0, xbatch=3),
}
SMPL2OP = [52, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62]


def run_case(c):
    W = synth.WEIGHT_SETS[c['wset']] if 'wset' in c else (synth.RGB_STAGE3_WEIGHTS if c['optim_floor'] else synth.AMASS_STAGE3_WEIGHTS)
    prob = synth.make_stage3_problem(c['B'], c['T'], seed=c['seed'], overlap=c['overlap'], cam=c['optim_floor'])
    keys = ('joints2d', 'floor_plane', 'seq_interval') if c['optim_floor'] else ('verts3d',)
    pts = W.get('points3d', 0.0) > 0.0
    ref, mo, _, _ = ref_closure.build(c['B'], c['T'], W, c['optim_floor'], prob['cam_mat'] if c['optim_floor'] else None)
    obs = {k: torch.as_tensor(prob['obs'][k]) for k in keys}
    if c['optim_floor']:
        ref_closure.set_params(mo, prob['params'], requires_grad=False)
        with torch.no_grad():
            _, _, inter = ref_closure.stage3_closure(ref, mo, {k: v.clone() for k, v in obs.items()}, backward=False)
        j = torch.cat([inter['cam_pred']['joints3d'], inter['cam_pred']['joints3d_extra']], 2)[:, :, SMPL2OP].numpy()
        rng = np.random.RandomState(c['seed'] + 100)
        xy = j[..., :2] / j[..., 2:3] * np.asarray(synth.CAM_F) + np.asarray(synth.CAM_C) + rng.randn(*j.shape[:3], 2) * 2.0
        obs['joints2d'] = obs['joints2d'].clone()
        obs['joints2d'][..., :2] = torch.as_tensor(xy.astype(np.float32))
        if pts:
            obs['points3d'] = torch.as_tensor(synth.sample_point_cloud(inter['cam_pred']['points3d'].numpy(), c['n_obs'],
                                                                        seed=c['seed'] + 200))
            keys = keys + ('points3d',)
        if c.get('xbatch'):
            # what run_fitting.py:428-435 caches after the previous batch: key vertices of its last sequence (here: the current
            # first sequence's own start + noise, so the term sits in a realistic range), betas, 4-parameter floor, interval
            ov, T = c['xbatch'], c['T']
            rng = np.random.RandomState(c['seed'] + 300)
            cv = inter['cam_pred']['verts3d'][0].numpy()                      # (T, 43, 3)
            pv = cv[0][None] + np.cumsum(rng.randn(T, 43, 3) * 0.01, 0)
            pv[T - ov:] = cv[:ov] + rng.randn(ov, 43, 3) * 0.02
            n = rng.randn(3) * 0.1 + [0.0, -1.0, 0.0]
            n /= np.linalg.norm(n)
            s0 = int(obs['seq_interval'][0, 0])
            obs['prev_batch_overlap_res'] = {
                'verts3d': torch.as_tensor(pv.astype(np.float32)),
                'betas': torch.as_tensor((prob['params']['betas'][0] + rng.randn(16) * 0.1).astype(np.float32)),
                'floor_plane': torch.as_tensor(np.concatenate([n, [1.1 + 0.1 * rng.randn()]]).astype(np.float32)),
                'seq_interval': torch.as_tensor(np.array([s0 + ov - T, s0 + ov], dtype=np.int64))}
            keys = keys + ('prev_batch_overlap_res',)
    names = ref_closure.set_params(mo, prob['params'])
    clone = lambda v: {a: b.clone() for a, b in v.items()} if isinstance(v, dict) else v.clone()
    loss, stats, inter = ref_closure.stage3_closure(ref, mo, {k: clone(v) for k, v in obs.items()}, c['nsteps'], c['scale'])
    out = {'loss': np.float32(loss.item())}
    for k, v in stats.items():
        out['stat_' + k] = np.float32(float(v))
    for n in names:
        out['grad_' + n] = getattr(mo, n).grad.numpy()
    for k in keys:
        if isinstance(obs[k], dict):
            for a, b in obs[k].items():
                out['obs_prevres_' + a] = b.numpy()
        else:
            out['obs_' + k] = obs[k].numpy()
    out['cam_verts3d'] = inter['cam_pred']['verts3d'].detach().numpy()
    out['cam_joints3d'] = inter['cam_pred']['joints3d'].detach().numpy()
    out['rollout_trans'] = inter['rollout']['trans'].detach().numpy()
    out['rollout_pose_body'] = inter['rollout']['pose_body'].detach().numpy()
    out['cond_prior_mean'] = inter['rollout']['cond_prior'][0].detach().numpy()
    out['cond_prior_var'] = inter['rollout']['cond_prior'][1].detach().numpy()
    out['meta'] = np.array([c['B'], c['T'], c['seed'], c['overlap'], -1 if c['nsteps'] is None else c['nsteps'], int(c['optim_floor'])])
    out['scale'] = np.float32(c['scale'])
    if 'wset' in c:
        out['wset'] = np.array(c['wset'])
    return out


CASES12 = {
    'stage1_rgb': dict(wset='rgb', stage=0, optim_floor=True, B=4, T=6, seed=41, overlap=2),
    'stage2_rgb': dict(wset='rgb', stage=1, optim_floor=True, B=4, T=6, seed=42, overlap=2),
    'stage2_amass': dict(wset='amass', stage=1, optim_floor=False, B=2, T=7, seed=43, overlap=2),
    'stage2_proxd': dict(wset='proxd', stage=1, optim_floor=True, B=2, T=5, seed=44, overlap=2, n_obs=80),
}


def run_case12(c):
    """Stage-I/II closure of the unmodified reference (root_fit / smpl_fit) on per-frame variables."""
    W12 = synth.stage12_weights(c['wset'])
    W3 = synth.WEIGHT_SETS[c['wset']]
    B, T = c['B'], c['T']
    prob = synth.make_stage3_problem(B, T, seed=c['seed'], overlap=c['overlap'], cam=c['optim_floor'])
    params = synth.make_stage12_params(B, T, seed=c['seed'] + 1)
    keys = ('joints2d', 'floor_plane', 'seq_interval') if c['optim_floor'] else ('verts3d',)
    if c['wset'] == 'proxd':
        keys = ('joints2d', 'floor_plane')
    ref, mo, _, _ = ref_closure.build(B, T, [W12, W12, W3], c['optim_floor'], prob['cam_mat'] if c['optim_floor'] else None)
    obs = {k: torch.as_tensor(prob['obs'][k]) for k in keys}
    if c['optim_floor']:
        ref_closure.set_params12(mo, params, c['stage'])
        with torch.no_grad():
            _, _, inter = ref_closure.stage12_closure(ref, mo, {k: v.clone() for k, v in obs.items()}, c['stage'], backward=False)
        j = torch.cat([inter['pred']['joints3d'], inter['pred']['joints3d_extra']], 2)[:, :, SMPL2OP].numpy()
        rng = np.random.RandomState(c['seed'] + 100)
        xy = j[..., :2] / j[..., 2:3] * np.asarray(synth.CAM_F) + np.asarray(synth.CAM_C) + rng.randn(*j.shape[:3], 2) * 2.0
        obs['joints2d'] = obs['joints2d'].clone()
        obs['joints2d'][..., :2] = torch.as_tensor(xy.astype(np.float32))
        if W12.get('points3d', 0.0) > 0.0:
            obs['points3d'] = torch.as_tensor(synth.sample_point_cloud(inter['pred']['points3d'].numpy(), c['n_obs'], seed=c['seed'] + 200))
            keys = keys + ('points3d',)
    names = ref_closure.set_params12(mo, params, c['stage'])
    loss, stats, inter = ref_closure.stage12_closure(ref, mo, {k: v.clone() for k, v in obs.items()}, c['stage'])
    out = {'loss': np.float32(loss.item())}
    for k, v in stats.items():
        out['stat_' + k] = np.float32(float(v))
    for n in names:
        out['grad_' + n] = getattr(mo, n).grad.numpy()
    for k in keys:
        out['obs_' + k] = obs[k].numpy()
    out['pred_verts3d'] = inter['pred']['verts3d'].detach().numpy()
    out['pred_joints3d'] = inter['pred']['joints3d'].detach().numpy()
    out['meta'] = np.array([B, T, c['seed'], c['overlap'], c['stage'], int(c['optim_floor'])])
    out['wset'] = np.array(c['wset'])
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    import sys
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        out = run_case(c)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(name, 'loss', out['loss'], {k[5:]: float(v) for k, v in out.items() if k.startswith('stat_')})
    for name, c in CASES12.items():
        if only and name not in only:
            continue
        out = run_case12(c)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(name, 'loss', out['loss'], {k[5:]: float(v) for k, v in out.items() if k.startswith('stat_')})


if __name__ == '__main__':
    main()
