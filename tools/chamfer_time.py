"""Times the chamfer nearest-neighbour kernel at the PROX-RGBD size (4096 observed points x 6890 vertices per frame) and
reports pair evaluations per second against the fp32-pipe bound (11 instructions per pair: 3 FADD(sub) + 3 FMUL + 2 FADD
+ FSETP + 2 SEL; 148 SMs x 128 lanes x clock).  python tools/chamfer_time.py [frames]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from humor_b200.chamfer import chamfer_nn  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 240
n, m = 4096, 6890
g = torch.Generator(device='cuda').manual_seed(0)
pred = torch.randn(b, m, 3, device='cuda', generator=g)
obs = torch.randn(b, n, 3, device='cuda', generator=g)
for _ in range(2):
    chamfer_nn(obs, pred, one_way=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    d, _, i, _ = chamfer_nn(obs, pred, one_way=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pairs = b * n * m
clk = 1.965e9
bound = 148 * 128 * clk / 11.0
pred.requires_grad_(True)
d, _, i, _ = chamfer_nn(obs, pred, one_way=True)
gd = torch.rand_like(d)
torch.cuda.synchronize()
e0.record()
d.backward(gd)
e1.record()
torch.cuda.synchronize()
first_bwd_ms = e0.elapsed_time(e1)
# the reverse kernel alone, warm, through the C-ABI
import ctypes as C
from humor_b200 import _ext
g2 = torch.empty_like(pred)
L = _ext.lib()
for _ in range(2):
    e0.record()
    _ext.check(L.humor_chamfer_bwd(b, n, _ext.ptr(obs), m, _ext.ptr(pred.detach()), _ext.ptr(gd), _ext.ptr(i), None, None, None,
                                   _ext.ptr(g2), None, _ext.stream_ptr()), 'bwd')
    e1.record()
    torch.cuda.synchronize()
print(json.dumps({'kernel': 'chamfer_nn_kernel (one-way, obs->verts)', 'frames': b, 'pairs_per_frame': n * m, 'ms': ms,
                  'pairs_per_s': pairs / (ms * 1e-3), 'fp32_pipe_bound_pairs_per_s': bound, 'frac_of_bound': pairs / (ms * 1e-3) / bound,
                  'frames_per_s': b / (ms * 1e-3), 'bwd_first_call_ms': first_bwd_ms, 'bwd_kernel_ms': e0.elapsed_time(e1)}))
