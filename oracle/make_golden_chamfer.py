"""ORACLE (test infrastructure) — writes tests/golden/chamfer_*.npz from the UNMODIFIED reference chamfer module
compiled by oracle/build_ref.py (cd.forward / cd.backward, chamfer_distance.cpp:90-177), i.e. outputs of the reference
itself, run in the build container.  The fixtures travel to the GPU box; /root/reference does not.

    python oracle/build_ref.py && python oracle/make_golden_chamfer.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_cd_ref  # noqa: E402

CASES = {
    # name: (b, n, m, seed, kind)
    'small': (3, 37, 53, 0, 'normal'),
    'ragged': (2, 1, 1300, 1, 'normal'),             # a single query, more than one 1024-target tile
    'ties': (2, 64, 96, 2, 'grid'),                  # integer lattice: many exactly tied distances -> first minimum wins
    'dup': (1, 50, 40, 3, 'dup'),                    # duplicated targets and queries lying ON targets (distance 0)
    'mesh': (2, 512, 2100, 4, 'body'),               # body-sized cloud, > 2 target tiles, 1 query block
}


def make_case(b, n, m, seed, kind):
    g = np.random.default_rng(seed)
    if kind == 'grid':
        a = g.integers(-3, 4, (b, n, 3)).astype(np.float32)
        c = g.integers(-3, 4, (b, m, 3)).astype(np.float32)
    elif kind == 'dup':
        c = g.normal(size=(b, m // 2, 3)).astype(np.float32)
        c = np.concatenate([c, c], 1)
        a = np.concatenate([c[:, :n // 2], g.normal(size=(b, n - n // 2, 3)).astype(np.float32)], 1)
    elif kind == 'body':
        c = (g.normal(size=(b, m, 3)) * np.array([0.25, 0.9, 0.15])).astype(np.float32) + np.float32(1.5)
        a = (c[:, g.integers(0, m, n)] + g.normal(size=(b, n, 3)).astype(np.float32) * np.float32(0.02)).astype(np.float32)
    else:
        a = g.normal(size=(b, n, 3)).astype(np.float32)
        c = g.normal(size=(b, m, 3)).astype(np.float32)
    gd1 = g.normal(size=(b, n)).astype(np.float32)
    gd2 = g.normal(size=(b, m)).astype(np.float32)
    return a, c, gd1, gd2


def run_reference(cd, a, c, gd1, gd2):
    ta, tc = torch.from_numpy(a), torch.from_numpy(c)
    b, n, m = a.shape[0], a.shape[1], c.shape[1]
    d1, d2 = torch.zeros(b, n), torch.zeros(b, m)
    i1, i2 = torch.zeros(b, n, dtype=torch.int), torch.zeros(b, m, dtype=torch.int)
    cd.forward(ta, tc, d1, d2, i1, i2)
    g1, g2 = torch.zeros(b, n, 3), torch.zeros(b, m, 3)
    cd.backward(ta, tc, g1, g2, torch.from_numpy(gd1), torch.from_numpy(gd2), i1, i2)
    # one-way gradient (what points3d_loss back-propagates: dist2 is unused, so its incoming gradient is zero)
    h1, h2 = torch.zeros(b, n, 3), torch.zeros(b, m, 3)
    cd.backward(ta, tc, h1, h2, torch.from_numpy(gd1), torch.zeros(b, m), i1, i2)
    return {'dist1': d1.numpy(), 'dist2': d2.numpy(), 'idx1': i1.numpy(), 'idx2': i2.numpy(),
            'grad_xyz1': g1.numpy(), 'grad_xyz2': g2.numpy(), 'grad_xyz1_oneway': h1.numpy(), 'grad_xyz2_oneway': h2.numpy()}


def main():
    cd = load_cd_ref()
    if cd is None:
        raise SystemExit('oracle/_ref/cd_ref.so missing: run python oracle/build_ref.py first')
    out = os.path.join(ROOT, 'tests', 'golden')
    for name, spec in CASES.items():
        a, c, gd1, gd2 = make_case(*spec)
        res = run_reference(cd, a, c, gd1, gd2)
        np.savez_compressed(os.path.join(out, f'chamfer_{name}.npz'), xyz1=a, xyz2=c, grad_dist1=gd1, grad_dist2=gd2, **res)
        print(name, a.shape, c.shape, 'dist1 sum', float(res['dist1'].sum()))


if __name__ == '__main__':
    main()
