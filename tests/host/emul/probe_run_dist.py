"""Child process (one per rank, gloo): MotionOptimizer.run of the product on CPU tensors through the emulated kernels with the
sub-sequences SHARDED over ranks — halo exchange of the overlap energies (parallel.py) + joint L-BFGS (lbfgs.py).  world = 1
runs the same problem in one process (same optimiser implementation) for comparison.
argv: root lib out.npz rank world port B_total T seed n1 n2 n3 lbfgs_max_iter"""
import os
import sys

import numpy as np
import torch

root, lib, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
rank, world, port, Bt, T, seed, n1, n2, n3, mi = (int(x) for x in sys.argv[4:14])
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from tests import util_stage3 as U  # noqa: E402

if world > 1:
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
prob = synth.make_stage3_problem(Bt, T, seed=seed, overlap=2, cam=True)
B = Bt // world
sl = slice(rank * B, (rank + 1) * B)
prob = {'params': {k: v[sl] for k, v in prob['params'].items()}, 'obs': {k: v[sl] for k, v in prob['obs'].items()},
        'cam_mat': prob['cam_mat'][sl]}
W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
mo = U.build_product(B, T, W3, True, prob, device='cpu', contact_refine_only=True)
mo.fitting_loss.all_stage_loss_weights = [dict(W12), dict(W12), dict(W3)]
mo.fitting_loss.set_stage(0)
mo.use_cuda_graph = False
mo.lbfgs_impl = 'native'
mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = 4, 1, 2
obs = {k: torch.as_tensor(v) for k, v in prob['obs'].items() if k in U.obs_keys(True)}
if world > 1:
    from humor_b200.parallel import Shard
    mo.shard = Shard(rank, world, None, ov_max=8)
    mo.shard.prepare(obs['seq_interval'])
res, stages = mo.run(obs, num_iter=[n1, n2, n3], lbfgs_max_iter=mi)
np.savez(out_path, **{k: v.detach().numpy() for k, v in res.items()}, stage3_verts3d=stages['stage3']['verts3d'].detach().numpy())
if world > 1:
    dist.destroy_process_group()
print('{"ok": true}')
