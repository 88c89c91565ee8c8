"""ChamferDistance — drop-in for humor/utils/chamfer_distance/chamfer_distance.py:13-60 on the sm_100a
nearest-neighbour kernels (csrc/chamfer.cu, C-ABI ``humor_chamfer_fwd`` / ``humor_chamfer_bwd``).

Same call surface: ``ChamferDistance()(xyz1, xyz2) -> (dist1, dist2)`` with ``xyz1 (b,n,3)``, ``xyz2 (b,m,3)``;
``dist1[i,j]`` is the SQUARED distance from ``xyz1[i,j]`` to its nearest neighbour in ``xyz2[i]``.  Results equal the
reference's CPU path bit for bit (same rounding, first minimum wins) and the reverse pass is deterministic.
There is no CPU path: the reference JIT-compiles its extension at import; this one needs the prebuilt library.

``one_way=True`` (an extension; FittingLoss.points3d_loss consumes only ``dist1``, fitting_loss.py:385-392)
skips the xyz2 -> xyz1 search and returns ``dist2 = None``.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _ext


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, one_way=False):
        _ext.require_cuda(xyz1, xyz2)
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
            raise ValueError('expected xyz1 (b,n,3) and xyz2 (b,m,3)')
        a, b_ = _ext.f32c(xyz1), _ext.f32c(xyz2)
        b, n, m = a.shape[0], a.shape[1], b_.shape[1]
        dev = a.device
        dist1 = torch.empty(b, n, device=dev, dtype=torch.float32)
        idx1 = torch.empty(b, n, device=dev, dtype=torch.int32)
        dist2 = idx2 = None
        if not one_way:
            dist2 = torch.empty(b, m, device=dev, dtype=torch.float32)
            idx2 = torch.empty(b, m, device=dev, dtype=torch.int32)
        nl = C.c_int64(0)
        _ext.check(_ext.lib().humor_chamfer_fwd(b, n, _ext.ptr(a), m, _ext.ptr(b_), _ext.ptr(dist1), _ext.ptr(idx1),
                                                _ext.ptr(dist2), _ext.ptr(idx2), C.byref(nl), _ext.stream_ptr()),
                   'humor_chamfer_fwd')
        _ext.LaunchCounter.total += nl.value
        ctx.one_way = one_way
        ctx.save_for_backward(a, b_, idx1, idx2 if idx2 is not None else idx1)
        ctx.mark_non_differentiable(idx1)
        if one_way:
            return dist1, None, idx1
        ctx.mark_non_differentiable(idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2=None, *unused):
        a, b_, idx1, idx2 = ctx.saved_tensors
        b, n, m = a.shape[0], a.shape[1], b_.shape[1]
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g1 = torch.empty_like(a) if need1 else None
        g2 = torch.empty_like(b_) if need2 else None
        gd1 = _ext.f32c(graddist1) if graddist1 is not None else None
        gd2 = _ext.f32c(graddist2) if (graddist2 is not None and not ctx.one_way) else None
        nl = C.c_int64(0)
        _ext.check(_ext.lib().humor_chamfer_bwd(b, n, _ext.ptr(a), m, _ext.ptr(b_), _ext.ptr(gd1), _ext.ptr(idx1),
                                                _ext.ptr(gd2), _ext.ptr(idx2) if gd2 is not None else None,
                                                _ext.ptr(g1), _ext.ptr(g2), C.byref(nl), _ext.stream_ptr()),
                   'humor_chamfer_bwd')
        _ext.LaunchCounter.total += nl.value
        return g1, g2, None


def chamfer_nn(xyz1, xyz2, one_way=False):
    """(dist1, dist2, idx1, idx2) — the nearest-neighbour indices are what cd.forward fills into its idx arguments."""
    out = ChamferDistanceFunction.apply(xyz1, xyz2, one_way)
    if one_way:
        return out[0], None, out[2], None
    return out


class ChamferDistance(nn.Module):
    """chamfer_distance.py:58-60."""

    def __init__(self, one_way=False):
        super().__init__()
        self.one_way = one_way

    def forward(self, xyz1, xyz2):
        out = ChamferDistanceFunction.apply(xyz1, xyz2, self.one_way)
        return out[0], out[1]
