"""ORACLE side of tools/diag_nonfinite.py (test infrastructure, container only: needs /root/reference).

tools/diag_nonfinite.py (GPU) writes the Stage-III variables and the product's gradients of the last finite closure evaluation
before a run() left the finite numbers.  This script evaluates the UNMODIFIED reference's closure at the same variables on the CPU
and compares loss and per-sequence gradients; with --sub LO HI it re-evaluates the sub-problem of sequences [LO, HI) and prints the
gradient magnitudes at the roll-out's outputs and inputs (where the reverse pass through the decoder chain amplifies them).

    python -m oracle.diag_snapshot gpurun_out/r02m_diag_tensor_last_finite.npz [--scale 4.0] [--sub 80 85 --seq 82]
"""
import argparse
import json

import numpy as np
import torch

from humor_b200 import synth
from oracle import ref_closure
from tests import util_stage3 as U


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('snapshot')
    ap.add_argument('--scale', type=float, default=4.0, help='init_motion_scale of the phase the snapshot was taken in (T / 15 in the last one)')
    ap.add_argument('--sub', type=int, nargs=2, default=None)
    ap.add_argument('--seq', type=int, default=-1)
    ap.add_argument('--seed', type=int, default=4)
    args = ap.parse_args()
    snap = np.load(args.snapshot)
    B, T = snap['latent_motion'].shape[0], snap['latent_motion'].shape[1] + 1
    lo, hi = args.sub if args.sub else (0, B)
    torch.set_num_threads(16)
    prob = synth.make_stage3_problem(B, T, seed=args.seed, overlap=10, cam=True)
    W12, W3 = synth.stage12_weights('rgb'), synth.RGB_STAGE3_WEIGHTS
    cam = np.asarray(prob['cam_mat'])
    ref, mo, _, _ = ref_closure.build(hi - lo, T, [W12, W12, W3], True, cam[lo:hi] if cam.ndim == 3 else cam)
    mo.fitting_loss.set_stage(2)
    names = ref_closure.set_params(mo, {k: snap[k][lo:hi] for k in snap.files if not k.startswith('grad_')})
    obs = {k: torch.as_tensor(v[lo:hi]) for k, v in prob['obs'].items() if k in U.obs_keys(True)}
    loss, stats, inter = ref_closure.stage3_closure(ref, mo, obs, None, args.scale, backward=False)
    watch = {}
    if args.sub:
        for grp in ('rollout', 'cam_rollout', 'pred'):
            for k, v in inter[grp].items():
                if torch.is_tensor(v) and v.requires_grad and v.dim() >= 2 and v.shape[0] == hi - lo:
                    v.retain_grad()
                    watch[grp + '.' + k] = v
    loss.backward()
    out = {'ref_loss': float(loss.detach()), 'sequences': [lo, hi], 'per_variable': {}}
    for n in names:
        g = getattr(mo, n).grad.numpy().reshape(hi - lo, -1)
        o = snap['grad_' + n][lo:hi].reshape(hi - lo, -1)
        ng, no = np.linalg.norm(g, axis=1), np.linalg.norm(o, axis=1)
        rel = np.linalg.norm(g - o, axis=1) / (ng + 1e-30)
        worst = int(np.argmax(rel * (ng > 0)))
        out['per_variable'][n] = {'sequences_within_1e-2': int((rel < 1e-2).sum()), 'median_rel': float(np.median(rel)),
                                  'worst_sequence': worst + lo, 'worst_ref_norm': float(ng[worst]), 'worst_product_norm': float(no[worst]),
                                  'worst_cosine': float((g[worst] * o[worst]).sum() / (ng[worst] * no[worst] + 1e-30))}
    if args.sub and args.seq >= 0:
        i = args.seq - lo
        out['grad_magnitudes_of_sequence'] = {k: float(v.grad[i].abs().max()) for k, v in watch.items() if v.grad is not None}
        out['root_angle_per_frame'] = [round(float(x), 3) for x in inter['rollout']['root_orient'][i].detach().norm(dim=-1)]
    out['stats'] = {k: float(v) for k, v in stats.items()}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
