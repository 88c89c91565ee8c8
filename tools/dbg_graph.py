import sys, torch, warnings
sys.path.insert(0, '/root/repo')
import bench
from humor_b200 import synth
B, T = 8, 10
dev = torch.device('cuda', 0)
prob = bench.build_problem(B, T)
mo = bench.make_optimizer(B, T, prob, dev)
names = mo.set_stage3_state(prob['params'])
obs = {k: torch.as_tensor(prob['obs'][k]).to(dev) for k in bench.OBS_KEYS}
params = [getattr(mo, n) for n in names]
mo.use_cuda_graph = True
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter('always')
    l = mo.stage3_step(obs, params=params)
    for x in w: print('WARN:', str(x.message)[:1500])
print('graph on:', mo.use_cuda_graph, float(l))
