"""Wall time of MotionOptimizer.run - the call a user of the reference makes (run_fitting.py:410-416): Stage I, II, Stage-III
initialisation and the three Stage-III phases with the reference's default iteration counts (30 / 80 / 70 outer iterations,
lbfgs_max_iter 20; configs/fit_rgb_demo_use_split.cfg) on the synthetic RGB problem - library L-BFGS (torch.optim.LBFGS, what the
reference builds) vs the native Gram-space L-BFGS, CUDA-graphed closure, default 'tensor' precision.  One JSON line per run.
  python tools/time_run.py [B] [iters "30,80,70"] [max_iter]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth, _ext  # noqa: E402
from tests import util_stage3 as U  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '30,80,70').split(',')]
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 20
T = 60
for impl in ('torch', 'native'):
    prob = synth.make_stage3_problem(B, T, seed=4, overlap=10, cam=True)
    W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
    mo = U.build_product(B, T, W3, True, prob, contact_refine_only=True)
    mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
    mo.fitting_loss.set_stage(0)
    mo.lbfgs_impl = impl
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    torch.cuda.synchronize()
    l0 = _ext.LaunchCounter.total
    t0 = time.perf_counter()
    res, stages = mo.run(obs, num_iter=iters, lbfgs_max_iter=max_iter)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({'B': B, 'T': T, 'num_iter': iters, 'lbfgs_max_iter': max_iter, 'lbfgs': impl, 'seconds': dt,
                      'frames': B * T, 'kernel_launches': _ext.LaunchCounter.total - l0,
                      'finite': bool(all(torch.isfinite(v).all() for v in res.values()))}), flush=True)
