import sys, numpy as np, torch
sys.path.insert(0, '.')
from humor_b200 import synth
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
gpu = torch.cuda.is_available()
from oracle.smplh_lbs import rodrigues, OracleBodyModel
from oracle.stage3_port import mat2aa
rng = np.random.RandomState(0)
n = 4000
aa = rng.randn(n, 3); aa *= (rng.uniform(0.01, 3.1, (n, 1)) / np.linalg.norm(aa, axis=1, keepdims=True))
G9 = rng.randn(n, 3, 3); g3 = rng.randn(n, 3)
out = {}
for dt in (torch.float64, torch.float32):
    a = torch.tensor(aa, dtype=dt, requires_grad=True)
    R = rodrigues(a); (R * torch.tensor(G9, dtype=dt)).sum().backward()
    Rm = rodrigues(torch.tensor(aa, dtype=torch.float64)).to(dt).detach().requires_grad_(True)
    x = mat2aa(Rm); (x * torch.tensor(g3, dtype=dt)).sum().backward()
    out[dt] = (R, a.grad, x, Rm.grad)
print('fp32 port  : rod fwd %.1e bwd %.1e | mat2aa fwd %.1e bwd %.1e' % tuple(rel(out[torch.float32][i], out[torch.float64][i]) for i in range(4)))
if gpu:
    from humor_b200.transforms import batch_rodrigues, rotation_matrix_to_angle_axis
    a = torch.tensor(aa, dtype=torch.float32, device='cuda', requires_grad=True)
    R = batch_rodrigues(a); (R * torch.tensor(G9, dtype=torch.float32).cuda()).sum().backward()
    Rm = rodrigues(torch.tensor(aa, dtype=torch.float64)).float().cuda().requires_grad_(True)
    x = rotation_matrix_to_angle_axis(Rm); (x * torch.tensor(g3, dtype=torch.float32).cuda()).sum().backward()
    print('cuda       : rod fwd %.1e bwd %.1e | mat2aa fwd %.1e bwd %.1e' % (rel(R, out[torch.float64][0]), rel(a.grad, out[torch.float64][1]), rel(x, out[torch.float64][2]), rel(Rm.grad, out[torch.float64][3])))
# LBS reverse (sparse: 43 key verts + 73 joints), betas per sequence
from humor_b200.body_model import KEYPT_VERTS
asset = synth.make_smplh_asset()
B, T = 3, 5; N = B * T
ro = rng.randn(N, 3) * 0.8; pb = rng.randn(N, 63) * 0.4; be = rng.randn(B, 16) * 0.5; tr = rng.randn(N, 3)
gs = rng.randn(N, 43, 3); gj = rng.randn(N, 73, 3)
res = {}
for dt in (torch.float64, torch.float32):
    om = OracleBodyModel(asset, use_vtx_selector=True, dtype=dt)
    t = [torch.tensor(x, dtype=dt, requires_grad=True) for x in (ro, pb, be, tr)]
    o = om(root_orient=t[0], pose_body=t[1], betas=t[2][:, None].expand(B, T, 16).reshape(N, 16), trans=t[3])
    ((o.v[:, KEYPT_VERTS] * torch.tensor(gs, dtype=dt)).sum() + (o.Jtr * torch.tensor(gj, dtype=dt)).sum()).backward()
    res[dt] = [x.grad for x in t] + [o.v, o.Jtr]
print('fp32 port  : lbs bwd ro %.1e pb %.1e betas %.1e trans %.1e | v %.1e J %.1e' % tuple(rel(res[torch.float32][i], res[torch.float64][i]) for i in range(6)))
if gpu:
    from humor_b200.body_model import BodyModel, lbs
    bm = BodyModel(asset, num_betas=16, batch_size=N, use_vtx_selector=True).to('cuda')
    t = [torch.tensor(x, dtype=torch.float32, device='cuda', requires_grad=True) for x in (ro, pb, be, tr)]
    v, vs, J = lbs(bm.lbs_model, t[0], t[1], t[2], t[3], frames_per_beta=T, sel_ids=KEYPT_VERTS, want_dense=True, dense_grad=False, num_joints_out=73)
    ((vs * torch.tensor(gs, dtype=torch.float32).cuda()).sum() + (J * torch.tensor(gj, dtype=torch.float32).cuda()).sum()).backward()
    r = res[torch.float64]
    print('cuda       : lbs bwd ro %.1e pb %.1e betas %.1e trans %.1e | v %.1e J %.1e' % (rel(t[0].grad, r[0]), rel(t[1].grad, r[1]), rel(t[2].grad, r[2]), rel(t[3].grad, r[3]), rel(v, r[4]), rel(J, r[5])))
