"""Per-phase timeline of the persistent decoder-chain kernel (csrc/chain_persist.cuh) from the clock64 stamps CTA 0 records when
humor_chain_debug is armed: for every phase of a step (GEMM phases 0..3, glue 4) the median over steps of the intervals between
its events, in microseconds at the SM clock given.  Events (one thread each):
  TMA lane   0 before the dependency wait of the first k-block | 1 dependency satisfied | 2 last A tile issued
  MMA lane   3 first stage landed | 4 last stage landed
  epilogue   5 last chunk promoted | 6 peers' slabs free | 7 partials sent | 8 all partials here | 9 slab finalised+stored
             10 epilogue warps joined | 11 flag released
  glue warp  0 before the flag wait | 1 inputs ready | 2 row done | 3 flag released
  python tools/chain_timeline.py [B] [S] [MHz]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth, _ext  # noqa: E402
from humor_b200.humor_model import HumorModel  # noqa: E402
from tests.test_gpu_kernels import make_state  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 59
mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1965.0
precision = sys.argv[4] if len(sys.argv) > 4 else 'tensor'
m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
m.load_state_dict(synth.make_humor_state_dict())
m = m.cuda().eval()
m.set_precision(precision)
x0 = torch.tensor(make_state(B, 1)).cuda().requires_grad_(True)
z = (torch.randn(B, S, 48) * 0.5).cuda().requires_grad_(True)
L = _ext.lib()
EV = 16
bufs = {d: torch.zeros(S * 5 * EV, dtype=torch.int64, device='cuda') for d in ('fwd', 'bwd')}
for _ in range(2):
    w, p = m.roll_out_raw(x0, z, True)
    (w.sum() + p.sum()).backward()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
L.humor_chain_debug(C.c_void_p(bufs['fwd'].data_ptr()), bufs['fwd'].numel() * 8)
e[0].record()
w, p = m.roll_out_raw(x0, z, True)
e[1].record()
torch.cuda.synchronize()
L.humor_chain_debug(C.c_void_p(bufs['bwd'].data_ptr()), bufs['bwd'].numel() * 8)
(w.sum() + p.sum()).backward()
e[2].record()
torch.cuda.synchronize()
L.humor_chain_debug(None, 0)
out = {'B': B, 'S': S, 'sm_mhz': mhz, 'precision': precision, 'rollout_fwd_ms': e[0].elapsed_time(e[1]), 'rollout_bwd_ms': e[1].elapsed_time(e[2])}
us = lambda c: c / mhz
for d in ('fwd', 'bwd'):
    t = bufs[d].cpu().numpy().reshape(S, 5, EV).astype(np.float64)
    order = [4, 0, 1, 2, 3] if d == 'bwd' else [0, 1, 2, 3, 4]          # phase order inside a step
    rec = {}
    span = us(t[1:, order[0], 0 if order[0] != 4 else 0] - t[:-1, order[0], 0])          # step period seen by CTA 0
    rec['step_period_us_median'] = float(np.median(span[2:]))
    for ph in range(5):
        x = t[2:, ph]                                                  # skip the first steps (cold)
        if ph < 4:
            iv = {'dep_wait(0-1)': x[:, 1] - x[:, 0], 'issue_A(1-2)': x[:, 2] - x[:, 1], 'first_landed_after_dep(1-3)': x[:, 3] - x[:, 1],
                  'stream(3-4)': x[:, 4] - x[:, 3], 'last_landed_to_promoted(4-5)': x[:, 5] - x[:, 4], 'xfree_wait(5-6)': x[:, 6] - x[:, 5],
                  'send(6-7)': x[:, 7] - x[:, 6], 'xfull_wait(7-8)': x[:, 8] - x[:, 7], 'finalise(8-9)': x[:, 9] - x[:, 8],
                  'join(9-10)': x[:, 10] - x[:, 9], 'release(10-11)': x[:, 11] - x[:, 10], 'phase_total(0-11)': x[:, 11] - x[:, 0]}
        else:
            iv = {'flag_wait(0-1)': x[:, 1] - x[:, 0], 'row(1-2)': x[:, 2] - x[:, 1], 'release(2-3)': x[:, 3] - x[:, 2], 'phase_total(0-3)': x[:, 3] - x[:, 0]}
        rec[f'phase{ph}'] = {k: round(float(np.median(us(v))), 2) for k, v in iv.items()}
    # hand-over latencies between consecutive phases of CTA 0's own chain of events
    seq = []
    for i in range(len(order) - 1):
        a, b = order[i], order[i + 1]
        end_a = t[2:, a, 11 if a < 4 else 3]
        dep_b = t[2:, b, 1]
        seq.append((f'release_ph{a}->dep_seen_ph{b}', float(np.median(us(dep_b - end_a)))))
    rec['handover_us'] = {k: round(v, 2) for k, v in seq}
    out[d] = rec
print(json.dumps(out, indent=1))
