"""Child process: MotionOptimizer.run (all three stages + Stage-III initialisation) of the product on CPU tensors through the
emulated kernels; writes the optimisation result to an npz.  argv: root lib out.npz B T seed n1 n2 n3 lbfgs_max_iter"""
import sys

import numpy as np
import torch

root, lib, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
B, T, seed, n1, n2, n3, mi = (int(x) for x in sys.argv[4:11])
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from tests import util_stage3 as U  # noqa: E402

prob = synth.make_stage3_problem(B, T, seed=seed, overlap=2, cam=True)
W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
mo = U.build_product(B, T, W3, True, prob, device='cpu', contact_refine_only=True)
mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
mo.fitting_loss.set_stage(0)
mo.use_cuda_graph = False
mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = 4, 1, 2
obs = {k: torch.as_tensor(v) for k, v in prob['obs'].items() if k in U.obs_keys(True)}
import tempfile  # noqa: E402
from tests import golden_util as GU  # noqa: E402
with tempfile.TemporaryDirectory() as tmp:
    dirs = GU.make_stage_dirs(tmp, B)
    res, stages = mo.run(obs, num_iter=[n1, n2, n3], lbfgs_max_iter=mi, stages_res_out=dirs)
    files = GU.collect_stage_files(dirs)
    assert not any(__import__('os').path.exists(d + '/stage3_results.npz') for d in dirs)      # written by save_optim_result only
np.savez(out_path, **{k: v.detach().numpy() for k, v in res.items()},
         stage3_verts3d=stages['stage3']['verts3d'].detach().numpy(), stage1_joints3d=stages['stage1']['joints3d'].detach().numpy(),
         stage2_joints3d=stages['stage2']['joints3d'].detach().numpy(),
         stage3_init_joints3d=stages['stage3_init']['joints3d'].detach().numpy(), **files)
print('{"ok": true}')
