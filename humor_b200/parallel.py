"""Multi-GPU sharding of the Stage-III batch.

One process per GPU; rank r owns the contiguous sub-sequences [r*B_local, (r+1)*B_local) of one video and
their 2 961 optimisation variables each.  Every energy is per-sequence except the overlap-consistency terms
(fitting_loss.py:136-157 key-vertex positions/velocities, :211-215 betas, :296-300 floor), which couple
ADJACENT sequences only.  Inside a rank they are evaluated by the fused kernel; across the rank boundary the
last sequence of rank r and the first of rank r+1 exchange a halo:

    tail pack of rank r  = [ verts3d[-1, T-ov_max:T] (ov_max*43*3) | betas[-1] (16) | floor[-1] (3) ]

Forward: rank r sends its pack to r+1 and receives r-1's (a few KB, latency-bound; NCCL send/recv over NVLink on
GPUs, gloo in the CPU tests); the receiving rank evaluates the boundary energy.  Backward: the gradient w.r.t. the
received pack is sent back to r-1, which adds it to its tail (HB_HALO=allgather selects the round-1 world-wide
all_gather / all_reduce pair instead).  Scalars shared by a joint
L-BFGS (loss, directional derivatives) go through `allreduce_scalars` — one collective per evaluation.
"""
import torch
import torch.distributed as dist


class Shard:
    def __init__(self, rank=0, world=1, group=None, ov_max=16):
        self.rank, self.world, self.group, self.ov_max = rank, world, group, ov_max
        self.ov_prev = None          # frames shared with the previous rank's last sequence (host int, set by prepare)

    def prepare(self, seq_interval):
        """One-time host exchange of the interval ends so that the per-evaluation halo code has no host
        synchronisation (and is CUDA-graph capturable together with its collectives)."""
        iv = seq_interval.detach().to('cpu', torch.int64)
        mine = torch.tensor([int(iv[0, 0]), int(iv[-1, 1])], dtype=torch.int64)
        if self.world > 1:
            dev = seq_interval.device if seq_interval.is_cuda else torch.device('cpu')
            buf = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(self.world)]
            dist.all_gather(buf, mine.to(dev), group=self.group)
            ends = [int(b[1]) for b in buf]
        else:
            ends = [int(mine[1])]
        self.ov_prev = 0 if self.rank == 0 else max(0, ends[self.rank - 1] - int(mine[0]))
        return self

    @staticmethod
    def from_env(ov_max=16):
        if dist.is_available() and dist.is_initialized():
            return Shard(dist.get_rank(), dist.get_world_size(), None, ov_max)
        return Shard()


class _GatherPacks(torch.autograd.Function):
    """all_gather with the matching reverse: d(pack_r) = sum over ranks of their gradient w.r.t. slot r."""

    @staticmethod
    def forward(ctx, shard, pack):
        out = [torch.empty_like(pack) for _ in range(shard.world)]
        dist.all_gather(out, pack.contiguous(), group=shard.group)
        ctx.shard = shard
        return torch.stack(out, 0)

    @staticmethod
    def backward(ctx, d_all):
        shard = ctx.shard
        d_all = d_all.contiguous()
        dist.all_reduce(d_all, op=dist.ReduceOp.SUM, group=shard.group)
        return None, d_all[shard.rank]


class _NeighbourPack(torch.autograd.Function):
    """Halo as a +-1 neighbour exchange: rank r sends its tail pack to r+1 and receives r-1's (zeros on rank 0); in reverse the
    gradient w.r.t. the received pack travels back to r-1.  Same values as the all_gather / all_reduce pair (only rank r+1 ever
    reads rank r's slot), but no rank waits for ranks it shares no frames with: with the world-wide collectives un-synchronised
    back-to-back graph replays lost 21 % at 8 ranks (SCALE_r01: 0.789)."""

    @staticmethod
    def forward(ctx, shard, pack):
        pack = pack.contiguous()
        prev = torch.zeros_like(pack)
        ops = []
        if shard.rank + 1 < shard.world:
            ops.append(dist.P2POp(dist.isend, pack, shard.rank + 1, shard.group))
        if shard.rank > 0:
            ops.append(dist.P2POp(dist.irecv, prev, shard.rank - 1, shard.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ctx.shard = shard
        return prev

    @staticmethod
    def backward(ctx, d_prev):
        shard = ctx.shard
        d_prev = d_prev.contiguous()
        d_pack = torch.zeros_like(d_prev)
        ops = []
        if shard.rank > 0:
            ops.append(dist.P2POp(dist.isend, d_prev, shard.rank - 1, shard.group))
        if shard.rank + 1 < shard.world:
            ops.append(dist.P2POp(dist.irecv, d_pack, shard.rank + 1, shard.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return None, d_pack


def previous_pack(shard, pack):
    """The tail pack of rank r-1 (zeros on rank 0) and a handle that keeps the reverse exchange symmetric on every rank."""
    import os
    if os.environ.get('HB_HALO', 'neighbour') == 'allgather':       # round-1 form, kept for A/B runs
        allp = _GatherPacks.apply(shard, pack)
        prev = allp[shard.rank - 1] if shard.rank > 0 else allp[0] * 0.0
        return prev, allp.sum() * 0.0
    prev = _NeighbourPack.apply(shard, pack)
    return prev, prev.sum() * 0.0


def tail_pack(shard, verts3d, betas, floor, T):
    """[ov_max*129 + 16 + 3] floats describing the LAST local sequence (zero-padded in front when T < ov_max)."""
    ov = shard.ov_max
    v = verts3d[-1, max(0, T - ov):T].reshape(-1)
    if v.numel() < ov * 129:
        v = torch.cat([torch.zeros(ov * 129 - v.numel(), device=v.device, dtype=v.dtype), v])
    f = floor[-1] if floor is not None else torch.zeros(3, device=v.device, dtype=v.dtype)
    return torch.cat([v, betas[-1, :16], f])


def boundary_overlap_energy(shard, verts3d, betas, floor, seq_interval, T, with_betas=True):
    """Overlap-consistency energy between this rank's FIRST sequence and the previous rank's LAST one
    (unweighted; same formulas as fitting_loss.py:142-157,211-215,296-300).  The overlap length with the previous
    rank comes from Shard.prepare (host, once); nothing here synchronises with the host."""
    ov_max = shard.ov_max
    if shard.ov_prev is None:
        shard.prepare(seq_interval)
    pack = tail_pack(shard, verts3d, betas, floor, T)
    prev, keep = previous_pack(shard, pack)                             # (P,), 0-valued handle on the exchange
    zero = verts3d.sum() * 0.0
    stats = {}
    ov = shard.ov_prev
    if shard.rank == 0 or ov <= 0:
        return zero + keep, stats                                       # keeps the reverse exchange symmetric
    if ov > ov_max or ov > T:
        raise ValueError(f'overlap {ov} exceeds the halo capacity {ov_max} / sequence length {T}')
    a = prev[:ov_max * 129].reshape(ov_max, 43, 3)[ov_max - ov:]        # tail of the previous rank's last sequence
    c = verts3d[0, :ov]
    d = a - c
    pos = 0.5 * (d ** 2).sum()
    vel = 0.5 * ((d[1:] - d[:-1]) ** 2).sum() if ov > 1 else zero
    e = pos + vel
    stats = {'rgb_overlap_consist_verts3d_pos': pos.detach(), 'rgb_overlap_consist_verts3d_vel': vel.detach() if ov > 1 else zero.detach()}
    if with_betas:             # Stage I (root_fit) couples the key vertices only (fitting_loss.py:136-157 vs :211-215)
        bet = 0.5 * ((prev[ov_max * 129:ov_max * 129 + 16] - betas[0, :16]) ** 2).sum()
        e = e + bet
        stats['rgb_overlap_consist_betas'] = bet.detach()
    if floor is not None:
        fl = 0.5 * ((prev[ov_max * 129 + 16:ov_max * 129 + 19] - floor[0]) ** 2).sum()
        e = e + fl
        stats['rgb_overlap_consist_floor'] = fl.detach()
    return e + keep, stats


def allreduce_scalars(shard, values):
    """Sum a small vector of per-rank scalars (loss, g.d, y.s, y.y, ...) in ONE collective."""
    if shard is None or shard.world == 1:
        return values
    dist.all_reduce(values, op=dist.ReduceOp.SUM, group=shard.group)
    return values
