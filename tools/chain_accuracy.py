"""Accuracy of the persistent decoder chain against the launch-per-layer chain and the exact-fp32 kernels on the same inputs
(forward states / prior, d init, d z) for several rollout lengths: separates summation-order noise amplified by the recurrence
from a real defect.  One JSON line per (B, S)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth  # noqa: E402
from humor_b200.humor_model import HumorModel  # noqa: E402
from tests.test_gpu_kernels import make_state  # noqa: E402

m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
m.load_state_dict(synth.make_humor_state_dict())
m = m.cuda().eval()


def run(x0n, zn, gw, gp, chain, precision):
    os.environ['HB_CHAIN'] = chain
    m.set_precision(precision)
    x0 = torch.tensor(x0n).cuda().requires_grad_(True)
    z = torch.tensor(zn).cuda().requires_grad_(True)
    w, p = m.roll_out_raw(x0, z, True)
    ((w * gw).sum() + (p * gp).sum()).backward()
    torch.cuda.synchronize()
    return [t.detach().double().cpu() for t in (w, p, x0.grad, z.grad)]


rel = lambda u, v: float((u - v).abs().max() / (v.abs().max() + 1e-12))
for B, S in [(256, 6), (256, 20), (256, 59), (200, 5)]:
    rng = np.random.RandomState(B + S)
    x0n = make_state(B, 1)
    zn = (rng.randn(B, S, 48) * 0.5).astype(np.float32)
    gw = torch.tensor(rng.randn(S, B, 348).astype(np.float32)).cuda()
    gp = torch.tensor(rng.randn(S, B, 96).astype(np.float32)).cuda()
    c = run(x0n, zn, gw, gp, '1', 'tensor')
    l = run(x0n, zn, gw, gp, '0', 'tensor')
    e = run(x0n, zn, gw, gp, '0', 'exact')
    names = ['world', 'prior', 'd_init', 'd_z']
    rec = {'B': B, 'S': S}
    for i, n in enumerate(names):
        rec[n] = {'chain_vs_legacy': rel(c[i], l[i]), 'chain_vs_exact': rel(c[i], e[i]), 'legacy_vs_exact': rel(l[i], e[i])}
    # per-step growth of the forward difference (world states)
    rec['world_chain_vs_exact_by_step'] = [float((c[0][t] - e[0][t]).abs().max() / (e[0][t].abs().max() + 1e-12)) for t in range(0, S, max(1, S // 8))]
    rec['world_legacy_vs_exact_by_step'] = [float((l[0][t] - e[0][t]).abs().max() / (e[0][t].abs().max() + 1e-12)) for t in range(0, S, max(1, S // 8))]
    print(json.dumps(rec), flush=True)
