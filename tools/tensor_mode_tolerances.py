"""Measures what the 'tensor' precision mode (tcgen05 3xTF32, persistent chain, CUDA graph) does END TO END, to set the tolerances
of its GPU tests from data: (1) MotionOptimizer.run against the result of the unmodified reference's run() (tests/golden/run_rgb.npz),
next to the same run in 'exact' mode; (2) full-length (T=60) closure gradients of a sub-batch against the CPU oracle port."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth  # noqa: E402
from tests import util_stage3 as U  # noqa: E402
from oracle.make_golden_run import CFG  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests')
g = np.load(os.path.join(HERE, 'golden', 'run_rgb.npz'))
out = {}
for precision, graph in (('exact', False), ('tensor', False), ('tensor', True)):
    prob = synth.make_stage3_problem(CFG['B'], CFG['T'], seed=CFG['seed'], overlap=CFG['overlap'], cam=True)
    W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
    mo = U.build_product(CFG['B'], CFG['T'], W3, True, prob, contact_refine_only=True)
    mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
    mo.fitting_loss.set_stage(0)
    mo.set_precision(precision)
    mo.use_cuda_graph = graph
    mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = CFG['tune_init']
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    res, stages = mo.run(obs, num_iter=list(CFG['num_iter']), lbfgs_max_iter=CFG['lbfgs_max_iter'])
    got = {k: v.detach().cpu().numpy() for k, v in res.items()}
    got['stage3_verts3d'] = stages['stage3']['verts3d'].detach().cpu().numpy()
    out[f'run_{precision}_graph{int(graph)}'] = {k: float(np.abs(got[k] - g[k]).max()) for k in got if k in g.files and got[k].dtype.kind == 'f'}
    out[f'run_{precision}_graph{int(graph)}']['contacts_equal'] = bool(np.array_equal(got['contacts'], g['contacts']))
print(json.dumps(out), flush=True)

# ---- (2) T = 60 gradients of an 8-sequence sub-batch vs the oracle port
B, T = 8, 60
W = dict(synth.RGB_STAGE3_WEIGHTS)
W['rgb_overlap_consist'] = 0.0
prob = synth.make_stage3_problem(B, T, seed=9, overlap=10)
port = U.build_port(B, T, W, True, prob)
_, _, aux = U.closure_port(port, prob, True)
cj = torch.cat([aux['inter']['cam_pred']['joints3d'], aux['inter']['cam_pred']['joints3d_extra']], 2).detach().numpy()
prob = U.project_joints2d(prob, cj)
l_c, g_c, _ = U.closure_port(port, prob, True)
rec = {}
for precision in ('exact', 'tensor'):
    mo = U.build_product(B, T, W, True, prob)
    mo.set_precision(precision)
    l_g, g_g, _ = U.closure_product(mo, prob)
    rec[precision] = {'loss_rel': abs(l_g - l_c) / max(1.0, abs(l_c)),
                      'grad_rel': {k: float((g_g[k].cpu() - g_c[k]).abs().max() / (g_c[k].abs().max() + 1e-8)) for k in g_c}}
print(json.dumps({'T60_subbatch_vs_oracle': rec}), flush=True)
