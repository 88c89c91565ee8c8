"""Child process: the product's Stage-I/II closure (MotionOptimizer.stage12_forward: VPoser decode -> mat2aa kernel -> SMPL+H
LBS kernels -> fused energy kernel -> reverse) on CPU tensors through the emulated kernels, against a golden fixture of the
unmodified reference.  Prints JSON {loss, stats, grad errors}."""
import json
import sys

import numpy as np
import torch

root, lib, name = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from humor_b200.body_model import BodyModel  # noqa: E402
from humor_b200.motion_optimizer import MotionOptimizer  # noqa: E402
from tests.golden_util import load_case12  # noqa: E402


class NoMotionPrior:                     # Stage I/II never touch the motion prior; the constructor only reads these
    latent_size, use_conditional_prior = 48, True


g, c = load_case12(name)
B, T = c['B'], c['T']
dev = torch.device('cpu')
bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=B * T, use_vtx_selector=c['optim_floor'])
obs = {k: torch.as_tensor(v) for k, v in c['obs'].items()}
mo = MotionOptimizer(dev, bm, 16, B, T, list(obs.keys()), [dict(c['W12']), dict(c['W12']), dict(c['W3'])], synth.FakeVPoser(),
                     NoMotionPrior(), {'gmm': synth.make_gmm()}, c['optim_floor'],
                     torch.as_tensor(c['cam_mat']) if c['optim_floor'] else None, 'bisquare', 4.6851, 100.0,
                     use_chamfer='points3d' in obs)
full = c['stage'] == 1
names = ['trans', 'root_orient'] + (['betas', 'latent_pose'] if full else [])
for k, v in c['params'].items():
    setattr(mo, k, torch.as_tensor(v).clone().requires_grad_(k in names))
mo.fitting_loss.set_stage(c['stage'])
loss, stats, pred = mo.stage12_forward(obs, c['stage'])
loss.backward()
out = {'loss': float(loss.detach()), 'stats': {k: float(v.detach()) for k, v in stats.items()},
       'verts_err': float(np.abs(pred['verts3d'].detach().numpy() - g['pred_verts3d']).max()), 'grad_err': {}}
for n in names:
    ref = g['grad_' + n]
    out['grad_err'][n] = float(np.abs(getattr(mo, n).grad.numpy() - ref).max() / (np.abs(ref).max() + 1e-8))
print(json.dumps(out))
