#!/usr/bin/env python
"""Per-kernel device time of ONE Stage-III step via torch.profiler (CUPTI; kernels are not serialised or
replayed, so the numbers are the in-situ ones).  Usage: python tools/profile_step.py [B] [T]"""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device('cuda', 0)
prob = bench.build_problem(B, T)
mo = bench.make_optimizer(B, T, prob, dev)
prob = bench.project_obs_from_product(mo, prob, dev)
names = mo.set_stage3_state(prob['params'])
obs = {k: torch.as_tensor(prob['obs'][k]).to(dev) for k in bench.OBS_KEYS}
params = [getattr(mo, n) for n in names]


mo.use_cuda_graph = False


def step():
    mo.stage3_step(obs, params=params)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name[:90]
    a = agg.setdefault(k, [0.0, 0])
    a[0] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
    a[1] += 1
tot = sum(v[0] for v in agg.values())
print(f'device busy {tot/1e3:.3f} ms in {sum(v[1] for v in agg.values())} kernels; B={B} T={T}')
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f'{100*us/tot:6.2f}% {us/1e3:9.3f} ms {n:6d} x {us/n:9.2f} us  {k}')
cpu = prof.key_averages().total_average()
print('cpu self time total (ms):', sum(e.self_cpu_time_total for e in prof.key_averages()) / 1e3)
