import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests import util_stage3 as U
from tests.golden_util import load_case
from humor_b200 import synth
g, prob, c = load_case('stage3_rgb')
p64 = {'params': {k: v.astype(np.float64) for k, v in prob['params'].items()}, 'obs': {k: (v.astype(np.float64) if v.dtype.kind == 'f' else v) for k, v in prob['obs'].items()}, 'cam_mat': prob['cam_mat']}
def run(W, tag):
    mo = U.build_product(c['B'], c['T'], W, True, prob)
    loss, grads, aux = U.closure_product(mo, prob, None, 1.0)
    port = U.build_port(c['B'], c['T'], W, True, prob, dtype=torch.float64)
    l64, g64, _ = U.closure_port(port, p64, True, None, 1.0)
    out = [f'{tag:22s} loss {abs(loss-l64)/max(abs(l64),1e-9):.1e}']
    for k in g64:
        r = g64[k].numpy(); s = np.abs(r).max()
        if s > 0: out.append(f'{k[:7]} {np.abs(grads[k].cpu().numpy()-r).max()/s:.1e}')
    print(' '.join(out))
run(synth.RGB_STAGE3_WEIGHTS, 'ALL')
for key in ['joints2d', 'rgb_overlap_consist', 'shape_prior', 'motion_prior', 'init_motion_prior', 'joint_consistency', 'bone_length', 'contact_vel', 'contact_height', 'floor_reg']:
    W = {k: 0.0 for k in synth.RGB_STAGE3_WEIGHTS}
    W[key] = synth.RGB_STAGE3_WEIGHTS[key]
    run(W, key)
