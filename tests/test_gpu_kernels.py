"""GPU parity of the individual kernels against the CPU oracle (through the C-ABI).
Tolerances: vertices L-inf 1e-4 m (north_star); everything else compared relative to the oracle's scale."""
import numpy as np
import pytest
import torch

from humor_b200 import synth

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(scope='module')
def asset():
    return synth.make_smplh_asset()


@pytest.fixture(scope='module')
def bm(asset):
    from humor_b200.body_model import BodyModel
    return BodyModel(asset, num_betas=16, batch_size=1, use_vtx_selector=True).to('cuda')


@pytest.fixture(scope='module')
def oracle_bm(asset):
    from oracle.smplh_lbs import OracleBodyModel
    return OracleBodyModel(asset, use_vtx_selector=True)


def rand_pose(n, seed, scale=0.4):
    rng = np.random.RandomState(seed)
    return (rng.randn(n, 3).astype(np.float32) * 0.8, (rng.randn(n, 63) * scale).astype(np.float32),
            (rng.randn(n, 16) * 0.7).astype(np.float32), rng.randn(n, 3).astype(np.float32))


@pytest.mark.parametrize('n', [1, 3, 70, 200])
def test_lbs_forward_matches_oracle(bm, oracle_bm, n):
    ro, pb, be, tr = rand_pose(n, n)
    o = oracle_bm(root_orient=torch.tensor(ro), pose_body=torch.tensor(pb), betas=torch.tensor(be), trans=torch.tensor(tr))
    g = bm(root_orient=torch.tensor(ro).cuda(), pose_body=torch.tensor(pb).cuda(), betas=torch.tensor(be).cuda(),
           trans=torch.tensor(tr).cuda())
    assert g.v.shape == (n, 6890, 3) and g.Jtr.shape == (n, 73, 3)
    assert float((g.v.cpu() - o.v).abs().max()) < 1e-4          # north_star: SMPL vertices within 1e-4 m
    assert float((g.Jtr.cpu() - o.Jtr).abs().max()) < 1e-4
    assert float((g.v.cpu() - o.v).abs().max()) < 2e-5          # what exact-fp32 actually delivers
    assert torch.equal(g.f.cpu(), o.f)


def test_lbs_known_answers(bm, asset):
    """zero pose + zero betas -> template + trans exactly; hands are identity -> no dependence on posedirs 189.."""
    n = 2
    z = lambda d: torch.zeros(n, d, device='cuda')
    tr = torch.tensor([[0.1, -0.2, 0.3], [1.0, 2.0, 3.0]], device='cuda')
    g = bm(root_orient=z(3), pose_body=z(63), betas=z(16), trans=tr)
    vt = torch.tensor(asset['v_template'], device='cuda')
    assert float((g.v - (vt[None] + tr[:, None])).abs().max()) < 1e-6
    Jt = torch.tensor(asset['J_regressor'] @ asset['v_template'], device='cuda')
    assert float((g.Jtr[:, :52] - (Jt[None] + tr[:, None])).abs().max()) < 1e-6


def test_lbs_backward_dense_matches_autograd(bm, oracle_bm):
    n = 5
    ro, pb, be, tr = rand_pose(n, 77)
    rng = np.random.RandomState(1)
    gv = rng.randn(n, 6890, 3).astype(np.float32)
    gj = rng.randn(n, 73, 3).astype(np.float32)
    cpu = [torch.tensor(x, requires_grad=True) for x in (ro, pb, be, tr)]
    o = oracle_bm(root_orient=cpu[0], pose_body=cpu[1], betas=cpu[2], trans=cpu[3])
    ((o.v * torch.tensor(gv)).sum() + (o.Jtr * torch.tensor(gj)).sum()).backward()
    gpu = [torch.tensor(x, device='cuda', requires_grad=True) for x in (ro, pb, be, tr)]
    g = bm(root_orient=gpu[0], pose_body=gpu[1], betas=gpu[2], trans=gpu[3])
    ((g.v * torch.tensor(gv).cuda()).sum() + (g.Jtr * torch.tensor(gj).cuda()).sum()).backward()
    for a, b, name in zip(gpu, cpu, ('root_orient', 'pose_body', 'betas', 'trans')):
        assert rel_err(a.grad, b.grad) < 2e-4, name


def test_lbs_selected_vertices_and_sparse_backward(bm, oracle_bm):
    """the Stage-III path: dense forward, gradient only through 43 key vertices + 73 joints, betas per sequence."""
    from humor_b200.body_model import lbs, KEYPT_VERTS
    B, T = 3, 5
    n = B * T
    ro, pb, _, tr = rand_pose(n, 9)
    be = (np.random.RandomState(3).randn(B, 16) * 0.5).astype(np.float32)
    rng = np.random.RandomState(2)
    gs = rng.randn(n, 43, 3).astype(np.float32)
    gj = rng.randn(n, 73, 3).astype(np.float32)
    cpu = [torch.tensor(x, requires_grad=True) for x in (ro, pb, be, tr)]
    bexp = cpu[2][:, None].expand(B, T, 16).reshape(n, 16)
    o = oracle_bm(root_orient=cpu[0], pose_body=cpu[1], betas=bexp, trans=cpu[3])
    ((o.v[:, KEYPT_VERTS] * torch.tensor(gs)).sum() + (o.Jtr * torch.tensor(gj)).sum()).backward()
    gpu = [torch.tensor(x, device='cuda', requires_grad=True) for x in (ro, pb, be, tr)]
    v, vs, J = lbs(bm.lbs_model, gpu[0], gpu[1], gpu[2], gpu[3], frames_per_beta=T, sel_ids=KEYPT_VERTS,
                   want_dense=True, dense_grad=False, num_joints_out=73)
    assert float((v.cpu() - o.v).abs().max()) < 2e-5
    assert float((vs.cpu() - o.v[:, KEYPT_VERTS]).abs().max()) < 2e-5
    ((vs * torch.tensor(gs).cuda()).sum() + (J * torch.tensor(gj).cuda()).sum()).backward()
    for a, b, name in zip(gpu, cpu, ('root_orient', 'pose_body', 'betas', 'trans')):
        assert rel_err(a.grad, b.grad) < 2e-4, name
    # list-mode forward (no dense output) gives the same selected vertices
    _, vs2, J2 = lbs(bm.lbs_model, gpu[0].detach(), gpu[1].detach(), gpu[2].detach(), gpu[3].detach(), frames_per_beta=T,
                     sel_ids=KEYPT_VERTS, want_dense=False, dense_grad=False, num_joints_out=52)
    assert float((vs2 - vs).abs().max()) < 1e-6 and float((J2 - J[:, :52]).abs().max()) < 1e-6


def test_rotation_kernels_match_oracle():
    from humor_b200.transforms import batch_rodrigues, rotation_matrix_to_angle_axis
    from oracle.smplh_lbs import rodrigues
    from oracle.stage3_port import mat2aa
    rng = np.random.RandomState(0)
    n = 5000
    aa = rng.randn(n, 3).astype(np.float32)
    aa *= (rng.uniform(0.01, 3.1, (n, 1)) / np.linalg.norm(aa, axis=1, keepdims=True)).astype(np.float32)
    a_c = torch.tensor(aa, requires_grad=True)
    a_g = torch.tensor(aa, device='cuda', requires_grad=True)
    R_c, R_g = rodrigues(a_c), batch_rodrigues(a_g)
    assert float((R_g.cpu() - R_c).abs().max()) < 2e-6
    G = rng.randn(n, 3, 3).astype(np.float32)
    (R_c * torch.tensor(G)).sum().backward()
    (R_g * torch.tensor(G).cuda()).sum().backward()
    assert rel_err(a_g.grad, a_c.grad) < 1e-4
    Rm_c = R_c.detach().clone().requires_grad_(True)
    Rm_g = R_c.detach().clone().cuda().requires_grad_(True)
    x_c, x_g = mat2aa(Rm_c), rotation_matrix_to_angle_axis(Rm_g)
    assert float((x_g.cpu() - x_c).abs().max()) < 5e-6
    g = rng.randn(n, 3).astype(np.float32)
    (x_c * torch.tensor(g)).sum().backward()
    (x_g * torch.tensor(g).cuda()).sum().backward()
    assert rel_err(Rm_g.grad, Rm_c.grad) < 1e-4
    # identity: the reference yields aa = 0 (NaN -> 0); ours too, with a finite gradient
    eye = torch.eye(3, device='cuda')[None].clone().requires_grad_(True)
    z = rotation_matrix_to_angle_axis(eye)
    z.sum().backward()
    assert float(z.abs().max()) == 0.0 and torch.isfinite(eye.grad).all()


def make_state(B, seed):
    """valid rollout input state (B,339) as the reference builds it (rotations as matrices)."""
    from oracle.smplh_lbs import rodrigues
    rng = np.random.RandomState(seed)
    x = np.zeros((B, 339), np.float32)
    x[:, 0:3] = rng.randn(B, 3) * 0.05 + [0, 0, 0.95]
    x[:, 3:6] = rng.randn(B, 3) * 0.1
    R = rodrigues(torch.tensor((rng.randn(B * 22, 3) * 0.4).astype(np.float32))).numpy().reshape(B, 22, 9)
    x[:, 6:15], x[:, 15:18], x[:, 18:207] = R[:, 0], rng.randn(B, 3) * 0.1, R[:, 1:].reshape(B, 189)
    x[:, 207:273], x[:, 273:339] = rng.randn(B, 66) * 0.3, rng.randn(B, 66) * 0.1
    return x


@pytest.fixture(scope='module')
def humor():
    from humor_b200.humor_model import HumorModel
    m = HumorModel(in_rot_rep='mat', out_rot_rep='aa', latent_size=48, model_data_config='smpl+joints+contacts', steps_in=1)
    m.load_state_dict(synth.make_humor_state_dict())
    return m.to('cuda').eval()


def port_rollout(x0, z):
    from oracle import stage3_port as sp
    sd = {k: v.to(x0.dtype) for k, v in synth.make_humor_state_dict().items()}
    names = ['trans', 'trans_vel', 'root_orient', 'root_orient_vel', 'pose_body', 'joints', 'joints_vel']
    dims = [3, 3, 9, 3, 189, 66, 66]
    init, s = {}, 0
    for nme, d in zip(names, dims):
        init[nme] = x0[:, s:s + d]
        s += d
    out, (pm, pv) = sp.roll_out(sd, init, z)
    world = torch.cat([out[k] for k in names + ['contacts']], -1)       # (B,S,348)
    return world, pm, pv


@pytest.mark.parametrize('B,S', [(2, 3), (5, 9), (33, 4)])
def test_rollout_forward_matches_oracle(humor, B, S):
    x0 = make_state(B, B)
    z = (np.random.RandomState(S).randn(B, S, 48) * 0.5).astype(np.float32)
    # oracle in fp64: both fp32 implementations carry ~1e-6 of their own rounding, the bound is on ours
    world_c, pm_c, pv_c = port_rollout(torch.tensor(x0).double(), torch.tensor(z).double())
    world_g, prior_g = humor.roll_out_raw(torch.tensor(x0).cuda(), torch.tensor(z).cuda(), True)
    world_g = world_g.permute(1, 0, 2).cpu()
    # north_star: decoder states within 1e-5 relative
    assert rel_err(world_g, world_c) < 1e-5
    assert rel_err(prior_g[..., :48].permute(1, 0, 2), pm_c) < 1e-5
    assert rel_err(torch.exp(prior_g[..., 48:]).permute(1, 0, 2), pv_c) < 1e-5


def test_decoder_step_and_prior_logprob_config2(humor):
    """BASELINE config 2: decoder single step + prior log-prob on 256 synthetic states."""
    from oracle import stage3_port as sp
    import math
    B = 256
    x0 = make_state(B, 5)
    z = (np.random.RandomState(6).randn(B, 48) * 0.5).astype(np.float32)
    sd = synth.make_humor_state_dict()
    dec_c = sp.decode(sd, torch.tensor(z), torch.tensor(x0))
    pm_c, pv_c = sp.prior_net(sd, torch.tensor(x0))
    lp_c = (-torch.log(torch.sqrt(pv_c)) - math.log(math.sqrt(2 * math.pi)) - (torch.tensor(z) - pm_c) ** 2 / (2 * pv_c)).sum(-1)
    xg, zg = torch.tensor(x0).cuda(), torch.tensor(z).cuda()
    dec_g = humor.decode(zg, xg)
    pm_g, pv_g = humor.prior(xg)
    lp_g = (-torch.log(torch.sqrt(pv_g)) - math.log(math.sqrt(2 * math.pi)) - (zg - pm_g) ** 2 / (2 * pv_g)).sum(-1)
    assert rel_err(dec_g, dec_c) < 1e-5
    assert float(((lp_g.cpu() - lp_c).abs() / lp_c.abs()).max()) < 1e-5


def test_rollout_backward_matches_autograd(humor):
    B, S = 4, 6
    x0 = make_state(B, 3)
    z = (np.random.RandomState(4).randn(B, S, 48) * 0.5).astype(np.float32)
    rng = np.random.RandomState(8)
    gw = rng.randn(B, S, 348).astype(np.float32)
    gp = rng.randn(B, S, 96).astype(np.float32)
    xc, zc = torch.tensor(x0, requires_grad=True), torch.tensor(z, requires_grad=True)
    world_c, pm_c, pv_c = port_rollout(xc, zc)
    pr_c = torch.cat([pm_c, torch.log(pv_c)], -1)
    ((world_c * torch.tensor(gw)).sum() + (pr_c * torch.tensor(gp)).sum()).backward()
    xg, zg = torch.tensor(x0, device='cuda', requires_grad=True), torch.tensor(z, device='cuda', requires_grad=True)
    world_g, prior_g = humor.roll_out_raw(xg, zg, True)
    ((world_g.permute(1, 0, 2) * torch.tensor(gw).cuda()).sum() + (prior_g.permute(1, 0, 2) * torch.tensor(gp).cuda()).sum()).backward()
    assert rel_err(xg.grad, xc.grad) < 1e-4
    assert rel_err(zg.grad, zc.grad) < 1e-4


def test_roll_out_api_dict(humor):
    """HumorModel.roll_out keeps the reference's call surface and output dict."""
    B, S = 3, 4
    x0 = torch.tensor(make_state(B, 1)).cuda()
    z = (torch.randn(B, S, 48) * 0.3).cuda()
    names, dims = ['trans', 'trans_vel', 'root_orient', 'root_orient_vel', 'pose_body', 'joints', 'joints_vel'], [3, 3, 9, 3, 189, 66, 66]
    init, s = {}, 0
    for n, d in zip(names, dims):
        init[n] = x0[:, None, s:s + d]
        s += d
    out, (pm, pv) = humor.roll_out(None, init, S, z_seq=z, return_prior=True)
    assert out['trans'].shape == (B, S, 3) and out['pose_body'].shape == (B, S, 189) and out['contacts'].shape == (B, S, 9)
    assert pm.shape == (B, S, 48) and bool((pv > 0).all())
    sd = humor.state_dict()
    assert 'prior_net.net.12.weight' in sd and 'decoder.net.9.bias' in sd and 'encoder.net.0.weight' in sd


def test_gmm_matches_torch_distributions():
    from humor_b200.fitting_loss import build_gmm, _GmmFn
    from torch.distributions import MixtureSameFamily, Categorical, MultivariateNormal
    w, m, c = synth.make_gmm()
    x = (torch.randn(7, 138) * 0.3)
    xc = x.clone().requires_grad_(True)
    ref = -MixtureSameFamily(Categorical(w), MultivariateNormal(m, covariance_matrix=c)).log_prob(xc)
    ref.sum().backward()
    gmm = build_gmm(w.cuda(), m.cuda(), c.cuda())
    xg = x.clone().cuda().requires_grad_(True)
    nll = _GmmFn.apply(gmm, xg)
    nll.sum().backward()
    assert rel_err(nll, ref) < 1e-5
    assert rel_err(xg.grad, xc.grad) < 1e-4
