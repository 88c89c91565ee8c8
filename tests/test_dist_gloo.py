"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): the halo exchange of parallel.py must
reproduce — energy and gradients — the overlap-consistency coupling a single process sees between the two
sequences that straddle the rank boundary."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _single_process(verts, betas, floor, iv, T):
    """the coupling between global sequences B-1 and B (fitting_loss.py:142-157,211-215,296-300)."""
    B2 = verts.shape[0]
    b = B2 // 2
    ov = int(iv[b - 1, 1] - iv[b, 0])
    a, c = verts[b - 1, T - ov:], verts[b, :ov]
    d = a - c
    e = 0.5 * (d ** 2).sum() + 0.5 * ((d[1:] - d[:-1]) ** 2).sum()
    e = e + 0.5 * ((betas[b - 1] - betas[b]) ** 2).sum() + 0.5 * ((floor[b - 1] - floor[b]) ** 2).sum()
    return e


def _worker(rank, world, port, data, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from humor_b200.parallel import Shard, boundary_overlap_energy, allreduce_scalars
    verts, betas, floor, iv, T = data
    B = verts.shape[0] // world
    sl = slice(rank * B, (rank + 1) * B)
    v = verts[sl].clone().requires_grad_(True)
    be = betas[sl].clone().requires_grad_(True)
    fl = floor[sl].clone().requires_grad_(True)
    shard = Shard(rank, world, None, ov_max=8)
    e, stats = boundary_overlap_energy(shard, v, be, fl, iv[sl], T)
    e.backward()
    tot = allreduce_scalars(shard, e.detach().clone().reshape(1))
    out[rank] = (float(tot), v.grad.clone(), be.grad.clone(), fl.grad.clone(), sorted(stats))
    dist.destroy_process_group()


def test_halo_exchange_matches_single_process():
    torch.manual_seed(0)
    world, B, T, ov = 2, 3, 9, 4
    verts = torch.randn(world * B, T, 43, 3)
    betas = torch.randn(world * B, 16)
    floor = torch.randn(world * B, 3)
    step = T - ov
    iv = torch.tensor([[i * step, i * step + T] for i in range(world * B)], dtype=torch.int32)
    v0, b0, f0 = (x.clone().requires_grad_(True) for x in (verts, betas, floor))
    e_ref = _single_process(v0, b0, f0, iv, T)
    e_ref.backward()
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, (verts, betas, floor, iv, T), out), nprocs=world, join=True)
    assert abs(out[0][0] - float(e_ref)) < 1e-3 * abs(float(e_ref)) and abs(out[1][0] - float(e_ref)) < 1e-3 * abs(float(e_ref))
    gv = torch.cat([out[0][1], out[1][1]], 0)
    gb = torch.cat([out[0][2], out[1][2]], 0)
    gf = torch.cat([out[0][3], out[1][3]], 0)
    assert torch.allclose(gv, v0.grad, atol=1e-5)
    assert torch.allclose(gb, b0.grad, atol=1e-5)
    assert torch.allclose(gf, f0.grad, atol=1e-5)
    assert out[0][4] == [] and 'rgb_overlap_consist_verts3d_pos' in out[1][4]
