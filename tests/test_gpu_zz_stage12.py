"""GPU parity of the Stage-I (root_fit) / Stage-II (smpl_fit) closures (MotionOptimizer.stage12_forward) against fixtures of
the unmodified reference.  The same closures pass on the CPU through the emulated kernels (tests/test_emul_product.py); this
is the on-device twin."""
import numpy as np
import pytest
import torch

from humor_b200 import synth

pytestmark = pytest.mark.gpu


class NoMotionPrior:
    latent_size, use_conditional_prior = 48, True


@pytest.mark.parametrize('name', ['stage1_rgb', 'stage2_rgb', 'stage2_amass', 'stage2_proxd'])
def test_stage12_closure_matches_reference_golden(name):
    from humor_b200.body_model import BodyModel
    from humor_b200.motion_optimizer import MotionOptimizer
    from tests.golden_util import load_case12, check_against_golden
    g, c = load_case12(name)
    B, T = c['B'], c['T']
    dev = torch.device('cuda')
    bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=B * T, use_vtx_selector=c['optim_floor']).to(dev)
    obs = {k: torch.as_tensor(v).to(dev) for k, v in c['obs'].items()}
    mo = MotionOptimizer(dev, bm, 16, B, T, list(obs.keys()), [dict(c['W12']), dict(c['W12']), dict(c['W3'])], synth.FakeVPoser().to(dev),
                         NoMotionPrior(), {'gmm': tuple(x.to(dev) for x in synth.make_gmm())}, c['optim_floor'],
                         torch.as_tensor(c['cam_mat']).to(dev) if c['optim_floor'] else None, 'bisquare', 4.6851, 100.0,
                         use_chamfer='points3d' in obs)
    names = ['trans', 'root_orient'] + (['betas', 'latent_pose'] if c['stage'] == 1 else [])
    for k, v in c['params'].items():
        setattr(mo, k, torch.as_tensor(v).to(dev).clone().requires_grad_(k in names))
    mo.fitting_loss.set_stage(c['stage'])
    loss, stats, pred = mo.stage12_forward(obs, c['stage'])
    loss.backward()
    check_against_golden(g, float(loss.detach()), {k: float(v.detach()) for k, v in stats.items()},
                         {n: getattr(mo, n).grad for n in names}, loss_tol=1e-5, stat_tol=1e-4, grad_tol=1e-4)
    assert np.abs(pred['verts3d'].detach().cpu().numpy() - g['pred_verts3d']).max() < 1e-5
