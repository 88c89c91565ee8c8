import sys, torch, warnings, traceback
sys.path.insert(0, '.')
from humor_b200 import synth
from tests import util_stage3 as U
B, T = 4, 8
prob = synth.make_stage3_problem(B, T, seed=6, overlap=3)
mo = U.build_product(B, T, synth.RGB_STAGE3_WEIGHTS, True, prob)
names = mo.set_stage3_state(prob['params'])
obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
params = [getattr(mo, n) for n in names]
mo.use_cuda_graph = False
mo.stage3_step(obs, params=params)
torch.cuda.synchronize()
for p in params:
    if p.grad is None: p.grad = torch.zeros_like(p)
# replicate the capture by hand to get the real traceback
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): mo._eval_into_static(obs, None, 1.0, params)
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        try:
            l = mo._eval_into_static(obs, None, 1.0, params)
        except Exception:
            traceback.print_exc()
            raise
    print('capture ok')
except Exception as e:
    print('CAPTURE FAILED:', str(e)[:300])
