"""Child process: humor_rollout_outputs_fwd / _bwd (csrc/rot.cu) on the CPU emulation against the torch-op form of the same step
(MotionOptimizer._rollout_outputs_torch: permute / slices / matrix->axis-angle / concatenations / apply_cam2prior inverse) with
autograd, with and without the camera frame.  Prints JSON."""
import json
import sys

import numpy as np
import torch

root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import motion_optimizer as MO  # noqa: E402
from humor_b200.transforms import batch_rodrigues  # noqa: E402

B, S = 5, 7
rng = np.random.RandomState(11)


class Stub:
    """the attributes of MotionOptimizer the two forms read"""
    _rollout_outputs_torch = MO.MotionOptimizer._rollout_outputs_torch
    apply_cam2prior = MO.MotionOptimizer.apply_cam2prior

    def __init__(self, cam):
        self.optim_floor = cam
        self._contact_idx = torch.tensor(MO.CONTACT_INDS, dtype=torch.long)
        self._contact_idx32 = self._contact_idx.to(torch.int32)
        self.init_fidx = np.zeros(B)
        self.cam2prior_root_height = None


def rot(n):
    return batch_rodrigues(torch.tensor(rng.randn(n, 3).astype(np.float32) * 0.9)).detach()


def inputs():
    world = torch.tensor(rng.randn(S, B, 348).astype(np.float32))
    R_all = rot(S * B * 22).reshape(S, B, 22 * 9)
    world[..., 6:15] = R_all[..., :9]
    world[..., 18:207] = R_all[..., 9:]
    t = lambda *sh: torch.tensor(rng.randn(*sh).astype(np.float32))
    v = {'world': world, 'trans': t(B, 1, 3), 'root_orient': t(B, 1, 3) * 0.8, 'body_pose': t(B, 1, 63) * 0.5, 'joints': t(B, 1, 22, 3),
         'R': rot(B).reshape(B, 3, 3), 't': t(B, 3)}
    return {k: x.clone().requires_grad_(True) for k, x in v.items()}


out = {}
for cam in (True, False):
    v = inputs()
    weights = None
    res = {}
    for form in ('kernel', 'torch'):
        m = Stub(cam)
        vv = {k: x.detach().clone().requires_grad_(True) for k, x in v.items()}
        m.cam2prior_R, m.cam2prior_t = vv['R'], vv['t']
        if form == 'kernel':
            r = MO._RolloutOutputs.apply(vv['world'], vv['trans'][:, 0], vv['root_orient'][:, 0], vv['body_pose'][:, 0],
                                         vv['joints'][:, 0].reshape(B, 66), vv['R'] if cam else None, vv['t'] if cam else None,
                                         m._contact_idx32)
            o = {'trans': r[0], 'root_orient': r[1], 'pose_body': r[2], 'joints': r[3], 'contacts_logits': r[4], 'contacts_conf': r[5],
                 'contacts': r[6]}
            c = {'trans': r[7], 'root_orient': r[8]} if cam else {'trans': r[0], 'root_orient': r[1]}
        else:
            o, c = m._rollout_outputs_torch(vv['world'], None, vv['trans'], vv['root_orient'], vv['body_pose'], None, vv['joints'], None, None,
                                            None, False, False)
        keys = ['trans', 'root_orient', 'pose_body', 'joints', 'contacts_logits']
        if weights is None:
            weights = {k: torch.tensor(rng.randn(*o[k].shape).astype(np.float32)) for k in keys}
            weights.update({'cam_' + k: torch.tensor(rng.randn(*c[k].shape).astype(np.float32)) for k in ('trans', 'root_orient')})
        loss = sum((o[k] * weights[k]).sum() for k in keys) + sum((c[k] * weights['cam_' + k]).sum() for k in ('trans', 'root_orient'))
        loss.backward()
        res[form] = ({**{k: o[k].detach().numpy() for k in keys + ['contacts_conf', 'contacts']}, **{'cam_' + k: c[k].detach().numpy() for k in c}},
                     {k: (x.grad.numpy() if x.grad is not None else None) for k, x in vv.items()})
    fk, ft = res['kernel'], res['torch']
    tag = 'cam' if cam else 'nocam'
    out[tag + '_fwd'] = {k: float(np.abs(fk[0][k] - ft[0][k]).max()) for k in fk[0]}
    out[tag + '_bwd_rel'] = {k: (None if ft[1][k] is None else float(np.abs(fk[1][k] - ft[1][k]).max() / (np.abs(ft[1][k]).max() + 1e-30)))
                             for k in ft[1] if (cam or k not in ('R', 't'))}
    out[tag + '_unused_grads_none'] = bool(cam or (fk[1]['R'] is None and fk[1]['t'] is None))
print(json.dumps(out))
