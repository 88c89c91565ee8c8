"""Child process: the persistent decoder-chain kernel (csrc/chain_persist.cuh) against the launch-per-layer chain it replaces
(same tcgen05 3xTF32 GEMMs, same glue arithmetic) on the CPU emulation: rollout forward (world states, prior) and reverse
(d init, d z) for a batch that spans two ragged 128-row tiles, with a given number of resident clusters.  Prints JSON.
  probe_chain.py <root> <lib> B S clusters"""
import json
import os
import sys

import numpy as np
import torch

root, lib = sys.argv[1], sys.argv[2]
B, S, ncl = int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
os.environ['HB_CHAIN_CLUSTERS'] = ncl
os.environ['HB_EMUL_TENSOR'] = '1'
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from humor_b200.humor_model import HumorModel  # noqa: E402
from tests.test_gpu_kernels import make_state  # noqa: E402

m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
m.load_state_dict(synth.make_humor_state_dict())
m.eval()
rng = np.random.RandomState(3)
x0n = make_state(B, 1)
zn = (rng.randn(B, S, 48) * 0.5).astype(np.float32)
gw = torch.tensor(rng.randn(S, B, 348).astype(np.float32))
gp = torch.tensor(rng.randn(S, B, 96).astype(np.float32))


def run(chain):
    os.environ['HB_CHAIN'] = chain
    x0 = torch.tensor(x0n).requires_grad_(True)
    z = torch.tensor(zn).requires_grad_(True)
    w, p = m.roll_out_raw(x0, z, True)
    ((w * gw).sum() + (p * gp).sum()).backward()
    return w.detach().clone(), p.detach().clone(), x0.grad.clone(), z.grad.clone()


a = run('1')
b = run('0')
rel = lambda u, v: float((u - v).abs().max() / (v.abs().max() + 1e-12))
print(json.dumps({'B': B, 'S': S, 'clusters': ncl, 'world': rel(a[0], b[0]), 'prior': rel(a[1], b[1]), 'd_init': rel(a[2], b[2]),
                  'd_z': rel(a[3], b[3]), 'finite': bool(all(torch.isfinite(t).all() for t in a))}))
