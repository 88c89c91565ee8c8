"""ORACLE (test infrastructure) — tests/golden/run_rgb.npz: the result of the UNMODIFIED reference's MotionOptimizer.run
(all three stages, torch.optim.LBFGS with strong-Wolfe, Stage-III initialisation through the posterior; motion_optimizer.py:
200-676) on a seeded synthetic RGB problem, executed on the CPU of the build container.

    python -m oracle.make_golden_run
"""
import os

import numpy as np
import torch

from humor_b200 import synth
from oracle import ref_closure
from tests import golden_util as GU
from tests import util_stage3 as U

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'run_rgb.npz')
CFG = dict(B=2, T=6, seed=51, overlap=2, num_iter=[2, 2, 3], lbfgs_max_iter=3, tune_init=(4, 1, 2))


def reference_run(cfg=CFG):
    prob = synth.make_stage3_problem(cfg['B'], cfg['T'], seed=cfg['seed'], overlap=cfg['overlap'], cam=True)
    W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
    ref, mo, _, _ = ref_closure.build(cfg['B'], cfg['T'], [W12, W12, W3], True, prob['cam_mat'])
    mo.fitting_loss.set_stage(0)
    mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = cfg['tune_init']
    obs = {k: torch.as_tensor(v) for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        dirs = GU.make_stage_dirs(tmp, cfg['B'])
        res, stages = mo.run(obs, num_iter=list(cfg['num_iter']), lbfgs_max_iter=cfg['lbfgs_max_iter'], stages_res_out=dirs)
        files = GU.collect_stage_files(dirs)    # the per-stage dumps run() writes itself (motion_optimizer.py:424-455, 665-674)
    out = {k: v.detach().numpy() for k, v in res.items()}
    out.update(files)
    out['stage3_init_joints3d'] = stages['stage3_init']['joints3d'].detach().numpy()
    out['stage3_verts3d'] = stages['stage3']['verts3d'].detach().numpy()
    out['stage1_joints3d'] = stages['stage1']['joints3d'].detach().numpy()
    out['stage2_joints3d'] = stages['stage2']['joints3d'].detach().numpy()
    return out


if __name__ == '__main__':
    torch.set_num_threads(8)
    out = reference_run()
    np.savez_compressed(OUT, **out)
    print({k: v.shape for k, v in out.items()})
