"""Small driver for ncu: a few eager Stage-III closure evaluations at the benchmark size (no CUDA graph, so that ncu sees every
kernel as its own launch).  Usage: python tools/run_step_once.py [B] [T] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda', 0)
prob = bench.build_problem(B, T)
mo = bench.make_optimizer(B, T, prob, dev)
prob = bench.project_obs_from_product(mo, prob, dev)
names = mo.set_stage3_state(prob['params'])
obs = {k: torch.as_tensor(prob['obs'][k]).to(dev) for k in bench.OBS_KEYS}
params = [getattr(mo, n) for n in names]
mo.use_cuda_graph = False
for _ in range(steps):
    mo.stage3_step(obs, params=params)
torch.cuda.synchronize()
