"""Child process: the product's Stage-III closure (VPoser decode -> cam2prior -> CVAE rollout [exact-fp32 kernels] -> SMPL+H
LBS -> fused energies -> reverse through everything) on CPU tensors through the emulated kernels, against a golden fixture of
the unmodified reference.  Prints JSON."""
import json
import sys

import numpy as np
import torch

root, lib, name = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from tests import util_stage3 as U  # noqa: E402
from tests.golden_util import load_case  # noqa: E402

g, prob, c = load_case(name)
mo = U.build_product(c['B'], c['T'], c['W'], c['optim_floor'], prob, device='cpu')
mo.use_cuda_graph = False
import os  # noqa: E402
if os.environ.get('HB_EMUL_PRECISION'):                 # e.g. 'tensor16': forward decoder chain on fp16 hi/lo operand planes
    mo.set_precision(os.environ['HB_EMUL_PRECISION'])
loss, grads, aux = U.closure_product(mo, prob, c['nsteps'], c['scale'])
out = {'loss': loss, 'stats': aux['stats'], 'grad_err': {},
       'verts_err': float(np.abs(aux['cam_pred']['verts3d'].detach().numpy() - g['cam_verts3d']).max()),
       'trans_err': float(np.abs(aux['roll']['trans'].detach().numpy() - g['rollout_trans']).max()),
       'prior_mean_err': float(np.abs(aux['roll']['cond_prior'][0].detach().numpy() - g['cond_prior_mean']).max() / np.abs(g['cond_prior_mean']).max())}
for k in g:
    if k.startswith('grad_'):
        ref = g[k]
        out['grad_err'][k[5:]] = float(np.abs(grads[k[5:]].numpy() - ref).max() / (np.abs(ref).max() + 1e-8))
print(json.dumps(out))
