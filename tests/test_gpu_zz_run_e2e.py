"""End-to-end on the device: MotionOptimizer.run (three stages, library L-BFGS, Stage-III initialisation, exact-fp32 kernels,
eager launches) against the result of the UNMODIFIED reference's run() on the same seeded problem
(tests/golden/run_rgb.npz).  The same comparison passes on the CPU emulation (tests/test_emul_product.py, HB_SLOW_TESTS=1)."""
import os

import numpy as np
import pytest
import torch

from humor_b200 import synth
from tests import util_stage3 as U

# First hardware run: round 2, call r02a (profiles/r02a_gpu_tests_ungated.txt): green on the B200 with the tolerances below.
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('precision,graph', [('exact', False), ('tensor', True)])
def test_run_end_to_end_matches_reference_run(precision, graph, tmp_path):
    """'exact': fp32 FFMA kernels, eager launches.  'tensor' (the default mode of the product: tcgen05 3xTF32 GEMMs, persistent
    decoder chain, fused LBS, CUDA-graphed closure): measured on the B200 it lands as close to the reference's result as 'exact'
    does (profiles/r02g_tolerances.jsonl: trans 2.5e-6, pose_body 1.5e-6, latent_motion 1.06e-4 in both modes), so both share
    the tolerances of tests/test_emul_product.py::RUN_TOL."""
    from oracle.make_golden_run import CFG
    from tests.test_emul_product import check_run_result
    prob = synth.make_stage3_problem(CFG['B'], CFG['T'], seed=CFG['seed'], overlap=CFG['overlap'], cam=True)
    W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
    mo = U.build_product(CFG['B'], CFG['T'], W3, True, prob, contact_refine_only=True)
    mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
    mo.fitting_loss.set_stage(0)
    mo.set_precision(precision)
    mo.use_cuda_graph = graph
    mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = CFG['tune_init']
    obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    from tests import golden_util as GU
    dirs = GU.make_stage_dirs(tmp_path, CFG['B'])
    res, stages = mo.run(obs, num_iter=list(CFG['num_iter']), lbfgs_max_iter=CFG['lbfgs_max_iter'], stages_res_out=dirs)
    got = {k: v.detach().cpu().numpy() for k, v in res.items()}
    got.update(GU.collect_stage_files(dirs))
    got['stage3_init_joints3d'] = stages['stage3_init']['joints3d'].detach().cpu().numpy()
    for s in ('stage1', 'stage2'):
        got[s + '_joints3d'] = stages[s]['joints3d'].detach().cpu().numpy()
    got['stage3_verts3d'] = stages['stage3']['verts3d'].detach().cpu().numpy()
    check_run_result(got, np.load(os.path.join(HERE, 'golden', 'run_rgb.npz')))
