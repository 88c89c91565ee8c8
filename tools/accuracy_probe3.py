import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests import util_stage3 as U
from tests.golden_util import load_case
from humor_b200 import synth
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
g, prob, c = load_case('stage3_rgb')
W = c['W']
names = U.PARAM_NAMES + ['floor_plane']
def port_run(dt):
    port = U.build_port(c['B'], c['T'], W, True, prob, dtype=dt)
    p = {k: torch.as_tensor(prob['params'][k]).to(dt).clone().requires_grad_(True) for k in names}
    obs = {k: (torch.as_tensor(v).to(dt) if v.dtype.kind == 'f' else torch.as_tensor(v)) for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    loss, st, inter = port.closure(p, obs, None, 1.0)
    keep = {k: inter['rollout'][k] for k in ('trans', 'root_orient', 'pose_body', 'joints')}
    keep['cam_trans'] = inter['cam_rollout']['trans']; keep['cam_root_orient'] = inter['cam_rollout']['root_orient']
    keep['raw_trans'] = inter['rollout']['raw']['trans']; keep['raw_pose'] = inter['rollout']['raw']['pose_body']; keep['raw_joints'] = inter['rollout']['raw']['joints']
    keep['contacts_logits'] = inter['rollout']['contacts_logits']
    for v in keep.values(): v.retain_grad()
    loss.backward()
    return {k: v.grad for k, v in keep.items()}, {k: p[k].grad for k in names}
i64, g64 = port_run(torch.float64)
i32, g32 = port_run(torch.float32)
print('fp32 port intermediates vs fp64:', {k: f'{rel(i32[k], i64[k]):.1e}' for k in i64})
print('fp32 port params        vs fp64:', {k: f'{rel(g32[k], g64[k]):.1e}' for k in g64})
if torch.cuda.is_available():
    mo = U.build_product(c['B'], c['T'], W, True, prob)
    nm = mo.set_stage3_state(prob['params'])
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    loss, stats, roll, cam, cam_pred = mo.stage3_forward(obs, None, 1.0)
    keep = {k: roll[k] for k in ('trans', 'root_orient', 'pose_body', 'joints')}
    keep['cam_trans'] = cam['trans']; keep['cam_root_orient'] = cam['root_orient']; keep['contacts_logits'] = roll['contacts_logits']
    for v in keep.values(): v.retain_grad()
    loss.backward()
    print('cuda intermediates      vs fp64:', {k: f'{rel(v.grad, i64[k]):.1e}' for k, v in keep.items()})
    print('cuda params             vs fp64:', {k: f'{rel(getattr(mo, k).grad, g64[k]):.1e}' for k in nm})
