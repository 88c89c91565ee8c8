"""humor_b200.lbfgs.LBFGS (Gram-space two-loop recursion, packed scalar reads, optional process group) against
torch.optim.LBFGS — the optimiser the reference builds (motion_optimizer.py:228-231,281-284,461-478) — and, sharded over
two gloo ranks, against the single-process run of the same coupled objective (joint L-BFGS).  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humor_b200.lbfgs import LBFGS


def objective(n, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, generator=g)
    A = (A @ A.t() / n + 0.1 * torch.eye(n)).to(dtype)
    b = torch.randn(n, generator=g).to(dtype)

    def f(z):
        return 0.5 * z @ A @ z - b @ z + 0.05 * (z ** 4).sum() + 0.3 * torch.sin(z).sum()
    return f


def run(cls, f, shapes, steps, dtype=torch.float32, **kw):
    params = [torch.full(s, 0.05 * (i + 1), dtype=dtype, requires_grad=True) for i, s in enumerate(shapes)]
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


@pytest.mark.parametrize('history', [100, 4])
@pytest.mark.parametrize('max_iter', [20, 3])
def test_matches_torch_lbfgs_fp64(history, max_iter):
    """In double precision rounding cannot steer the line search: same evaluations, same iterates, call after call
    (the state carried across .step() calls and the eviction of old pairs at history 4 included)."""
    f = objective(60, dtype=torch.float64)
    shapes = [(4, 10), (20,)]
    l_ref, x_ref, o_ref = run(torch.optim.LBFGS, f, shapes, 5, dtype=torch.float64, max_iter=max_iter, history_size=history)
    l_new, x_new, o_new = run(LBFGS, f, shapes, 5, dtype=torch.float64, max_iter=max_iter, history_size=history)
    for a, b in zip(l_ref, l_new):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a)), (l_ref, l_new)
    assert float((x_ref - x_new).abs().max()) < 1e-7
    if max_iter == 3:       # far from convergence: the evaluation counts are identical too
        assert o_ref.state[o_ref._params[0]]['func_evals'] == o_new._st['func_evals']
    # one packed host read per closure evaluation + one per outer iteration
    assert o_new.syncs <= o_new._st['func_evals'] + o_new._st['n_iter']


def test_matches_torch_lbfgs_fp32():
    """fp32 (what the fitting runs in): both optimisers stop where the loss cannot be resolved any further; the end
    points are then equally good rather than equal (the valley floor is ~1e-3 wide in x at this resolution)."""
    f = objective(60)
    shapes = [(4, 10), (20,)]
    l_ref, x_ref, _ = run(torch.optim.LBFGS, f, shapes, 3, max_iter=20)
    l_new, x_new, _ = run(LBFGS, f, shapes, 3, max_iter=20)
    for a, b in zip(l_ref, l_new):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(a)), (l_ref, l_new)
    assert abs(float(f(x_ref)) - float(f(x_new))) <= 2e-5 * abs(float(f(x_ref)))
    assert float((x_ref - x_new).abs().max()) < 5e-3


def test_first_iterations_track_torch_closely():
    """Before rounding differences can steer the line search: identical evaluation count and iterates after one step."""
    f = objective(40, seed=3)
    l_ref, x_ref, o_ref = run(torch.optim.LBFGS, f, [(40,)], 1, max_iter=6)
    l_new, x_new, o_new = run(LBFGS, f, [(40,)], 1, max_iter=6)
    assert o_ref.state[o_ref._params[0]]['func_evals'] == o_new._st['func_evals']
    assert float((x_ref - x_new).abs().max()) < 1e-5


def test_frozen_parameter_without_grad():
    """A parameter whose .grad stays None contributes a zero gradient (what freezing the init state does,
    motion_optimizer.py:545-563) and is left untouched."""
    f = objective(30, seed=1)
    a = torch.zeros(20, requires_grad=True)
    frozen = torch.full((10,), 0.3)
    opt = LBFGS([a, frozen], max_iter=10)

    def closure():
        opt.zero_grad()
        loss = f(torch.cat([a, frozen]))
        loss.backward()
        return loss
    l0 = float(opt.step(closure))
    l1 = float(closure())
    assert l1 < l0 and torch.equal(frozen, torch.full((10,), 0.3))


def test_converged_start_returns_immediately():
    x = torch.zeros(5, requires_grad=True)
    opt = LBFGS([x])
    calls = []

    def closure():
        opt.zero_grad()
        loss = (x ** 2).sum()
        loss.backward()
        calls.append(1)
        return loss
    opt.step(closure)
    assert len(calls) == 1 and opt._st['n_iter'] == 0


# ------------------------------------------------------------------------------------------------ sharded (gloo, world 2)
def coupled_loss(x, f_local, nb, rank, world, gather):
    """sum of per-rank energies + a coupling between the last variable of rank r and the first of rank r+1,
    evaluated on the receiving rank only (the layout of parallel.boundary_overlap_energy)."""
    loss = f_local(x)
    if world > 1:
        tails = gather(x[-1:].clone())                    # (world, 1), differentiable
        if rank > 0:
            loss = loss + 2.0 * ((tails[rank - 1, 0] - x[0]) ** 2) + 0.0 * tails.sum()
        else:
            loss = loss + 0.0 * tails.sum()
    return loss


def _worker(rank, world, port, n, steps, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from humor_b200.parallel import Shard, _GatherPacks
    shard = Shard(rank, world, None)
    f_local = objective(n, seed=10 + rank, dtype=torch.float64)
    x = torch.full((n,), 0.05, dtype=torch.float64, requires_grad=True)
    opt = LBFGS([x], max_iter=8, group=True)
    losses = []
    for _ in range(steps):
        def closure():
            opt.zero_grad()
            loss = coupled_loss(x, f_local, n, rank, world, lambda t: _GatherPacks.apply(shard, t))
            loss.backward()
            return loss
        losses.append(float(opt.step(closure).detach()))
    out[rank] = (losses, x.detach().clone(), opt._st['func_evals'])
    dist.destroy_process_group()


def test_sharded_joint_lbfgs_matches_single_process():
    world, n, steps = 2, 24, 3
    fs = [objective(n, seed=10 + r, dtype=torch.float64) for r in range(world)]
    x = torch.full((world * n,), 0.05, dtype=torch.float64, requires_grad=True)
    opt = LBFGS([x], max_iter=8)
    ref_losses = []
    for _ in range(steps):
        def closure():
            opt.zero_grad()
            loss = fs[0](x[:n]) + fs[1](x[n:]) + 2.0 * (x[n - 1] - x[n]) ** 2
            loss.backward()
            return loss
        ref_losses.append(float(opt.step(closure).detach()))
    mgr = mp.Manager()
    out = mgr.dict()
    port = 23000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, steps, out), nprocs=world, join=True)
    xs = torch.cat([out[0][1], out[1][1]])
    # every rank returns ITS share of the loss; the shares add up to the single-process loss
    for i in range(steps):
        tot = out[0][0][i] + out[1][0][i]
        assert abs(tot - ref_losses[i]) <= 1e-9 * max(1.0, abs(ref_losses[i])), (i, tot, ref_losses[i])
    assert out[0][2] == out[1][2] == opt._st['func_evals']   # lock-step, and the same evaluations as one process
    assert float((xs - x.detach()).abs().max()) < 1e-7


def test_fixed_step_mode_matches_torch():
    """line_search_fn=None (fixed step lr): the other branch of the library optimiser, iterate for iterate in fp64."""
    f = objective(30, seed=5, dtype=torch.float64)

    def go(cls):
        x = torch.full((30,), 0.05, dtype=torch.float64, requires_grad=True)
        opt = cls([x], lr=0.5, max_iter=6, line_search_fn=None)
        out = []
        for _ in range(3):
            def closure():
                opt.zero_grad()
                loss = f(x)
                loss.backward()
                return loss
            out.append(float(opt.step(closure).detach()))
        return out, x.detach().clone()
    l_ref, x_ref = go(torch.optim.LBFGS)
    l_new, x_new = go(LBFGS)
    for a, b in zip(l_ref, l_new):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(a)), (l_ref, l_new)
    assert float((x_ref - x_new).abs().max()) < 1e-8
