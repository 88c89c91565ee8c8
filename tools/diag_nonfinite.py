"""Where does MotionOptimizer.run leave the finite numbers?  (tools/time_run.py, B=256 with the library L-BFGS: non-finite result.)
Every closure evaluation of the three stages is checked on the host; the first non-finite loss is reported with its stage, outer
iteration and evaluation index, the non-finite statistics terms, and the variables of that evaluation and of the last finite one
are written to gpurun_out/ so that the reference closure can be evaluated at the same point on the CPU (oracle side).
  python tools/diag_nonfinite.py [B] [precision] [lbfgs impl] [out prefix]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth  # noqa: E402
from tests import util_stage3 as U  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precision = sys.argv[2] if len(sys.argv) > 2 else 'tensor'
impl = sys.argv[3] if len(sys.argv) > 3 else 'torch'
prefix = sys.argv[4] if len(sys.argv) > 4 else 'gpurun_out/diag_nonfinite'
T = 60
iters = [30, 80, 70]

prob = synth.make_stage3_problem(B, T, seed=4, overlap=10, cam=True)
W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
mo = U.build_product(B, T, W3, True, prob, contact_refine_only=True)
mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
mo.fitting_loss.set_stage(0)
mo.set_precision(precision)
mo.lbfgs_impl = impl
obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}

NAMES = ['trans', 'root_orient', 'latent_pose', 'betas', 'latent_motion', 'trans_vel', 'joints_vel', 'root_orient_vel', 'floor_plane']
state = {'evals': 0, 'last_finite': None, 'bad': None, 'hist': []}


def snapshot():
    out = {}
    for n in NAMES:
        v = getattr(mo, n, None)
        if torch.is_tensor(v):
            out[n] = v.detach().cpu().numpy().copy()
            if v.grad is not None:
                out['grad_' + n] = v.grad.detach().cpu().numpy().copy()
    return out


def check(stage, loss, stats=None):
    state['evals'] += 1
    lv = float(loss.detach())
    fin = np.isfinite(lv)
    if state['bad'] is None:
        if fin:
            if state['evals'] % 16 == 0 or stage != state.get('stage'):
                state['last_finite'] = snapshot()
            state['stage'] = stage
            if len(state['hist']) < 4000:
                state['hist'].append((stage, int(mo.fitting_loss.cur_optim_step) if hasattr(mo.fitting_loss, 'cur_optim_step') else -1, lv))
        else:
            bad_terms = {}
            if stats:
                for k, v in stats.items():
                    x = float(torch.as_tensor(v).detach().float().sum())
                    if not np.isfinite(x):
                        bad_terms[k] = x
            snap = snapshot()
            nonfin = {k: int((~np.isfinite(v)).sum()) for k, v in snap.items() if (~np.isfinite(v)).any()}
            big = {k: float(np.nanmax(np.abs(v))) for k, v in snap.items()}
            state['bad'] = {'stage': stage, 'eval': state['evals'], 'loss': lv, 'bad_terms': bad_terms, 'nonfinite_entries': nonfin,
                            'max_abs': big, 'recent_losses': state['hist'][-12:]}
            np.savez_compressed(prefix + '_bad.npz', **snap)
            if state['last_finite'] is not None:
                np.savez_compressed(prefix + '_last_finite.npz', **state['last_finite'])
            print(json.dumps({'first_nonfinite': state['bad']}), flush=True)


orig12 = mo.stage12_forward


def stage12_forward(observed_data, stage):
    loss, st, pred = orig12(observed_data, stage)
    check(stage, loss, st)
    return loss, st, pred


orig3 = mo.stage3_step


def stage3_step(observed_data, nsteps=None, init_motion_scale=1.0, params=None):
    loss = orig3(observed_data, nsteps, init_motion_scale, params)
    check(2, loss)
    return loss


mo.stage12_forward = stage12_forward
mo.stage3_step = stage3_step
res, stages = mo.run(obs, num_iter=iters, lbfgs_max_iter=20)
torch.cuda.synchronize()
fin = {k: bool(torch.isfinite(v).all()) for k, v in res.items()}
h = state['hist']
summary = {'B': B, 'precision': precision, 'lbfgs': impl, 'evals': state['evals'], 'result_finite': fin,
           'first_nonfinite': None if state['bad'] is None else {k: state['bad'][k] for k in ('stage', 'eval')},
           'loss_first_last_by_stage': {s: [x[2] for x in h if x[0] == s][:1] + [x[2] for x in h if x[0] == s][-1:] for s in (0, 1, 2)}}
print(json.dumps(summary), flush=True)
