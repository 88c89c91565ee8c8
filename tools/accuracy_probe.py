import sys, numpy as np, torch
sys.path.insert(0, '.')
from humor_b200 import synth
from humor_b200.humor_model import HumorModel
from tests.test_gpu_kernels import make_state, port_rollout
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
m.load_state_dict(synth.make_humor_state_dict()); m = m.cuda().eval()
for B, S in ((4, 7), (4, 20)):
    x0 = make_state(B, 3); z = (np.random.RandomState(4).randn(B, S, 48) * 0.5).astype(np.float32)
    rng = np.random.RandomState(8)
    gw = rng.randn(B, S, 348).astype(np.float32); gp = rng.randn(B, S, 96).astype(np.float32)
    for only in ('world', 'prior', 'both'):
        res = {}
        for dt in (torch.float64, torch.float32):
            xc, zc = torch.tensor(x0, dtype=dt, requires_grad=True), torch.tensor(z, dtype=dt, requires_grad=True)
            wc, pm, pv = port_rollout(xc, zc)
            pr = torch.cat([pm, torch.log(pv)], -1)
            L = 0
            if only in ('world', 'both'): L = L + (wc * torch.tensor(gw, dtype=dt)).sum()
            if only in ('prior', 'both'): L = L + (pr * torch.tensor(gp, dtype=dt)).sum()
            L.backward(); res[dt] = (wc, pr, xc.grad, zc.grad)
        xg, zg = torch.tensor(x0, device='cuda', requires_grad=True), torch.tensor(z, device='cuda', requires_grad=True)
        wg, pg = m.roll_out_raw(xg, zg, True)
        L = 0
        if only in ('world', 'both'): L = L + (wg.permute(1, 0, 2) * torch.tensor(gw).cuda()).sum()
        if only in ('prior', 'both'): L = L + (pg.permute(1, 0, 2) * torch.tensor(gp).cuda()).sum()
        L.backward()
        r64, r32 = res[torch.float64], res[torch.float32]
        print(f'B={B} S={S} {only:5s} fwd world cuda {rel(wg.permute(1,0,2), r64[0]):.1e} (fp32 ref {rel(r32[0], r64[0]):.1e}) prior cuda {rel(pg.permute(1,0,2), r64[1]):.1e} (ref {rel(r32[1], r64[1]):.1e}) | dx0 cuda {rel(xg.grad, r64[2]):.1e} (ref {rel(r32[2], r64[2]):.1e}) dz cuda {rel(zg.grad, r64[3]):.1e} (ref {rel(r32[3], r64[3]):.1e})')
