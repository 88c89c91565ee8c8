"""GPU twin of tests/test_lbfgs.py: humor_b200.lbfgs.LBFGS (Gram-space two-loop recursion, one packed device->host read per
closure evaluation) on CUDA tensors against torch.optim.LBFGS - the optimiser the reference builds
(motion_optimizer.py:228-231,281-284,461-478) - on the same objective, and through MotionOptimizer.run."""
import os

import numpy as np
import pytest
import torch

from humor_b200.lbfgs import LBFGS

pytestmark = pytest.mark.gpu


def objective(n, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, generator=g)
    A = (A @ A.t() / n + 0.1 * torch.eye(n)).to(dtype).cuda()
    b = torch.randn(n, generator=g).to(dtype).cuda()

    def f(z):
        return 0.5 * z @ A @ z - b @ z + 0.05 * (z ** 4).sum() + 0.3 * torch.sin(z).sum()
    return f


def run(cls, f, shapes, steps, dtype=torch.float32, **kw):
    params = [torch.full(s, 0.05 * (i + 1), dtype=dtype, device='cuda', requires_grad=True) for i, s in enumerate(shapes)]
    opt = cls(params, lr=1.0, line_search_fn='strong_wolfe', **kw)
    losses = []
    for _ in range(steps):
        def closure():
            opt.zero_grad()
            loss = f(torch.cat([p.reshape(-1) for p in params]))
            loss.backward()
            return loss
        losses.append(float(opt.step(closure).detach()))
    return losses, torch.cat([p.detach().reshape(-1) for p in params]), opt


@pytest.mark.parametrize('history,max_iter', [(100, 20), (4, 3)])
def test_native_lbfgs_matches_torch_lbfgs_on_device_fp64(history, max_iter):
    """fp64 on the device: rounding cannot steer the line search, so iterates, losses and (far from convergence) evaluation
    counts agree with the library optimiser call after call; at most one packed host read per evaluation and per iteration."""
    f = objective(60, dtype=torch.float64)
    shapes = [(4, 10), (20,)]
    l_ref, x_ref, o_ref = run(torch.optim.LBFGS, f, shapes, 5, dtype=torch.float64, max_iter=max_iter, history_size=history)
    l_new, x_new, o_new = run(LBFGS, f, shapes, 5, dtype=torch.float64, max_iter=max_iter, history_size=history)
    for a, b in zip(l_ref, l_new):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a)), (l_ref, l_new)
    assert float((x_ref - x_new).abs().max()) < 1e-7
    if max_iter == 3:
        assert o_ref.state[o_ref._params[0]]['func_evals'] == o_new._st['func_evals']
    assert o_new.syncs <= o_new._st['func_evals'] + o_new._st['n_iter']


def test_native_lbfgs_fp32_on_device():
    """fp32 (what the fitting runs in): both optimisers stop where the loss cannot be resolved further - equally good end points."""
    f = objective(60)
    shapes = [(4, 10), (20,)]
    l_ref, x_ref, _ = run(torch.optim.LBFGS, f, shapes, 3, max_iter=20)
    l_new, x_new, _ = run(LBFGS, f, shapes, 3, max_iter=20)
    for a, b in zip(l_ref, l_new):
        assert abs(a - b) <= 5e-5 * max(1.0, abs(a)), (l_ref, l_new)
    assert abs(float(f(x_ref)) - float(f(x_new))) <= 5e-5 * abs(float(f(x_ref)))


def test_motion_optimizer_run_native_vs_library_lbfgs():
    """MotionOptimizer.run (three stages) with the native optimiser against the same run with torch.optim.LBFGS: the results
    agree to the optimisation's own noise floor (tolerances of tests/test_emul_product.py::RUN_TOL x 10: two fp32 line searches)."""
    from humor_b200 import synth
    from tests import util_stage3 as U
    B, T = 2, 8
    outs = {}
    for mode in ('library', 'native'):
        prob = synth.make_stage3_problem(B, T, seed=3, overlap=3)
        mo = U.build_product(B, T, synth.RGB_STAGE3_WEIGHTS, True, prob, contact_refine_only=True)
        mo.set_precision('exact')
        mo.use_cuda_graph = False
        mo.stage3_tune_init_num_frames, mo.stage3_tune_init_freeze_start, mo.stage3_tune_init_freeze_end = 4, 1, 2
        obs = {k: torch.as_tensor(v).cuda() for k, v in prob['obs'].items() if k in U.obs_keys(True)}
        mo.lbfgs_impl = 'native' if mode == 'native' else 'torch'
        res, _ = mo.run(obs, num_iter=[2, 2, 3], lbfgs_max_iter=5)
        outs[mode] = {k: v.detach().cpu().numpy() for k, v in res.items()}
    for k in ('trans', 'root_orient', 'pose_body', 'betas'):
        d = float(np.abs(outs['native'][k] - outs['library'][k]).max())
        assert d < 5e-3, (k, d)
    for v in outs['native'].values():
        assert np.isfinite(v).all()
