"""Child process: BodyModel forward/backward on CPU tensors through the emulated kernels vs the torch oracle (JSON out)."""
import json
import sys

import numpy as np
import torch

root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from humor_b200.body_model import BodyModel, lbs, KEYPT_VERTS  # noqa: E402
from oracle.smplh_lbs import OracleBodyModel  # noqa: E402

asset = synth.make_smplh_asset()
bm = BodyModel(asset, num_betas=16, batch_size=1, use_vtx_selector=True)
ob = OracleBodyModel(asset, use_vtx_selector=True)
rng = np.random.RandomState(0)
n = 5
ro, pb, be, tr = (torch.tensor(a, requires_grad=True) for a in (rng.randn(n, 3).astype(np.float32) * 0.8, (rng.randn(n, 63) * 0.4).astype(np.float32),
                                                                  (rng.randn(n, 16) * 0.7).astype(np.float32), rng.randn(n, 3).astype(np.float32)))
g = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
ro2, pb2, be2, tr2 = (x.detach().clone().requires_grad_(True) for x in (ro, pb, be, tr))
o = ob(root_orient=ro2, pose_body=pb2, betas=be2, trans=tr2)
out = {'v_err': float((g.v - o.v).abs().max()), 'J_err': float((g.Jtr - o.Jtr).abs().max())}
# the Stage-III pattern: gradient through 43 key vertices + 73 joints only
wv = torch.tensor(rng.randn(n, len(KEYPT_VERTS), 3).astype(np.float32))
wj = torch.tensor(rng.randn(n, 73, 3).astype(np.float32))
_, vs, J = lbs(bm.lbs_model, ro, pb, be, tr, 1, KEYPT_VERTS, False, False, 73)
((vs * wv).sum() + (J * wj).sum()).backward()
((o.v[:, KEYPT_VERTS] * wv).sum() + (o.Jtr * wj).sum()).backward()
for name, a, b in (('root_orient', ro, ro2), ('pose_body', pb, pb2), ('betas', be, be2), ('trans', tr, tr2)):
    out['g_' + name] = float((a.grad - b.grad).abs().max() / (b.grad.abs().max() + 1e-12))
print(json.dumps(out))
