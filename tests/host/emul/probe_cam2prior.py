"""Child process: humor_cam2prior_fwd / _bwd (csrc/rot.cu) on the CPU emulation against the torch-op form of the same function
(fitting_utils.compute_cam2prior_torch, evaluated in float64 with autograd).  Prints JSON."""
import json
import sys

import numpy as np
import torch

root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
from humor_b200 import fitting_utils as FU  # noqa: E402
from humor_b200 import transforms as TR  # noqa: E402

B = 67
rng = np.random.RandomState(5)
floor = rng.randn(B, 3) * np.array([0.3, 1.0, 0.3]) * (1.0 + rng.rand(B, 1) * 2.0)      # normals around +-y, offsets 1..3
floor[::2, 1] = -np.abs(floor[::2, 1]) - 0.2                                          # both signs of n_y (the flip of parse_floor_plane)
floor[1::2, 1] = np.abs(floor[1::2, 1]) + 0.2
trans = rng.randn(B, 3) * 1.5
orient = rng.randn(B, 3) * 0.9
orient[3] = 1e-4 * rng.randn(3)                                                       # near-identity rotation
joints = rng.randn(B, 22, 3)
gR, gt, gh = rng.randn(B, 3, 3), rng.randn(B, 3), rng.randn(B, 1)


def run(fn, dtype):
    v = [torch.tensor(x, dtype=dtype, requires_grad=True) for x in (floor, trans, orient, joints)]
    R, t, h = fn(*v)
    loss = (R * torch.tensor(gR, dtype=dtype)).sum() + (t * torch.tensor(gt, dtype=dtype)).sum() + (h * torch.tensor(gh, dtype=dtype)).sum()
    loss.backward()
    return [x.detach().double().numpy() for x in (R, t, h)], [x.grad.double().numpy() for x in v]


def torch64(f, t, r, j):
    # float64 oracle: batch_rodrigues of the product is a float32 kernel, so the rotation comes from torch's matrix exponential
    K = torch.zeros(B, 3, 3, dtype=torch.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -r[:, 2], r[:, 1], r[:, 2], -r[:, 0], -r[:, 1], r[:, 0]
    Rm = torch.linalg.matrix_exp(K)
    saved = FU.batch_rodrigues
    FU.batch_rodrigues = lambda aa: Rm
    try:
        return FU.compute_cam2prior_torch(f, t, r, j)
    finally:
        FU.batch_rodrigues = saved


(o_ref, g_ref), (o_k, g_k) = run(torch64, torch.float64), run(FU.compute_cam2prior, torch.float32)
out = {'fwd': [float(np.abs(a - b).max()) for a, b in zip(o_ref, o_k)],
       'bwd_rel': [float(np.abs(a - b).max() / (np.abs(a).max() + 1e-30)) for a, b in zip(g_ref, g_k)],
       'orthonormal': float(np.abs(np.einsum('bij,bkj->bik', o_k[0], o_k[0]) - np.eye(3)).max()),
       'joint_grad_only_root': bool(np.abs(g_k[3][:, 1:]).max() == 0.0), 'finite': bool(all(np.isfinite(x).all() for x in o_k + g_k))}
# a parsed (B,4) plane keeps the torch-op form (same values)
R4, t4, h4 = FU.compute_cam2prior(FU.parse_floor_plane(torch.tensor(floor, dtype=torch.float32)), torch.tensor(trans, dtype=torch.float32),
                                  torch.tensor(orient, dtype=torch.float32), torch.tensor(joints, dtype=torch.float32))
out['parsed_plane_vs_kernel'] = float(max(np.abs(R4.numpy() - o_k[0]).max(), np.abs(h4.numpy() - o_k[2]).max()))
print(json.dumps(out))
