"""Dense LBS forward: fused tcgen05 kernel vs the exact-fp32 FFMA path and the CPU oracle, plus device timing."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from humor_b200 import synth
from humor_b200.body_model import BodyModel, lbs


def inputs(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    N = B * T
    return ((torch.randn(N, 3, generator=g) * 0.5).cuda(), (torch.randn(N, 63, generator=g) * 0.3).cuda(),
            (torch.randn(B, 16, generator=g) * 0.5).cuda(), torch.randn(N, 3, generator=g).cuda())


def run(bm, args, T, mode):
    bm.set_precision(mode)
    with torch.no_grad():
        v, _, j = lbs(bm.lbs_model, *args, T, None, True, False, 73)
    torch.cuda.synchronize()
    return v, j


def timeit(bm, args, T, mode, reps=5):
    bm.set_precision(mode)
    with torch.no_grad():
        lbs(bm.lbs_model, *args, T, None, True, False, 73)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            lbs(bm.lbs_model, *args, T, None, True, False, 73)
            ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


asset = synth.make_smplh_asset()
bm = BodyModel(asset, num_betas=16, batch_size=1, use_vtx_selector=True).to('cuda')
# 1. ragged size against the CPU oracle
from oracle.smplh_lbs import OracleBodyModel
ob = OracleBodyModel(asset, use_vtx_selector=True)
B, T = 5, 60
a = inputs(B, T, 1)
v, j = run(bm, a, T, 'tensor')
o = ob(root_orient=a[0].cpu(), pose_body=a[1].cpu(), betas=a[2].cpu().repeat_interleave(T, 0), trans=a[3].cpu())
print('N=300 fused vs oracle: v', float((v.cpu().reshape(-1, 6890, 3) - o.v).abs().max()),
      'Jtr', float((j.cpu().reshape(-1, 73, 3) - o.Jtr).abs().max()), flush=True)
# 2. benchmark size against the exact path
B, T = 256, 60
a = inputs(B, T, 0)
vt, jt = run(bm, a, T, 'tensor')
ve, je = run(bm, a, T, 'exact')
print('N=15360 fused vs exact: v', float((vt - ve).abs().max()), 'Jtr', float((jt - je).abs().max()), flush=True)
for mode in ('tensor', 'exact'):
    ms = timeit(bm, a, T, mode)
    print(f'{mode}: {ms:.3f} ms  -> {B * T * 83896 / ms / 1e6:.1f} GB/s algorithmic', flush=True)
