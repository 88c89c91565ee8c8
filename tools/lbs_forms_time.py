"""Times the dense LBS forward (B x T = 256 x 60 frames) for the kernel forms of humor_lbs_configure and reports, per
form, whether its output differs bitwise from the default forms (a different summation order must show up in the last bits:
evidence that the form really ran).  python tools/lbs_forms_time.py [--only skin,blend] [--reps R] [--slab S]"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from humor_b200 import synth, _ext  # noqa: E402
from humor_b200.body_model import BodyModel, lbs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--only', default='')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--slab', type=int, default=512)
args = ap.parse_args()
B, T = 256, 60
N = B * T
bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=N, use_vtx_selector=True).to('cuda')
m = bm.lbs_model
g = torch.Generator(device='cpu').manual_seed(0)
ro = (torch.randn(N, 3, generator=g) * 0.5).cuda()
pb = (torch.randn(N, 63, generator=g) * 0.3).cuda()
be = (torch.randn(B, 16, generator=g) * 0.5).cuda()
tr = torch.randn(N, 3, generator=g).cuda()
L = _ext.lib()
cfgs = [(1, 1), (2, 1), (1, 2), (2, 2), (2, 3), (3, 1), (3, 3), (3, 4)]
if args.only:
    cfgs = [tuple(int(x) for x in args.only.split(','))]
ref = None
out = []
for skin, blend in cfgs:
    _ext.check(L.humor_lbs_configure(skin, blend, args.slab), 'configure')
    with torch.no_grad():
        for _ in range(2):
            v, _, J = lbs(m, ro, pb, be, tr, T, None, True, False, 73)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            v, _, J = lbs(m, ro, pb, be, tr, T, None, True, False, 73)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    import ctypes as C
    us, ub = C.c_int(0), C.c_int(0)
    L.humor_lbs_forms_used(C.byref(us), C.byref(ub))
    if ref is None:
        ref = v.clone()
    out.append({'skin': skin, 'blend': blend, 'used': [us.value, ub.value], 'slab': args.slab, 'ms': ms, 'GBps': N * 83896 / (ms * 1e-3) / 1e9,
                'max_abs_diff_vs_first': float((v - ref).abs().max()), 'bitwise_equal_to_first': bool(torch.equal(v, ref))})
    del v
print(json.dumps(out))
