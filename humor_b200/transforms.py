"""Rotation conversions with the reference's names and semantics (humor/utils/transforms.py),
running on the sm_100a kernels of csrc/rot.cu with hand-written reverse modes."""
import torch

from . import _ext


class _Rodrigues(torch.autograd.Function):
    @staticmethod
    def forward(ctx, aa):
        _ext.require_cuda(aa)
        aa = _ext.f32c(aa)
        n = aa.shape[0]
        R = torch.empty(n, 3, 3, device=aa.device, dtype=torch.float32)
        _ext.check(_ext.lib().humor_rodrigues_fwd(n, _ext.ptr(aa), _ext.ptr(R), _ext.stream_ptr()), 'humor_rodrigues_fwd')
        _ext.LaunchCounter.total += 1
        ctx.save_for_backward(aa)
        return R

    @staticmethod
    def backward(ctx, dR):
        (aa,) = ctx.saved_tensors
        dR = _ext.f32c(dR)
        daa = torch.empty_like(aa)
        _ext.check(_ext.lib().humor_rodrigues_bwd(aa.shape[0], _ext.ptr(aa), _ext.ptr(dR), _ext.ptr(daa), _ext.stream_ptr()),
                   'humor_rodrigues_bwd')
        _ext.LaunchCounter.total += 1
        return daa


class _Mat2AA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, R):
        _ext.require_cuda(R)
        R = _ext.f32c(R)
        n = R.shape[0]
        aa = torch.empty(n, 3, device=R.device, dtype=torch.float32)
        _ext.check(_ext.lib().humor_mat2aa_fwd(n, _ext.ptr(R), _ext.ptr(aa), _ext.stream_ptr()), 'humor_mat2aa_fwd')
        _ext.LaunchCounter.total += 1
        ctx.save_for_backward(R)
        return aa

    @staticmethod
    def backward(ctx, daa):
        (R,) = ctx.saved_tensors
        daa = _ext.f32c(daa)
        dR = torch.empty_like(R)
        _ext.check(_ext.lib().humor_mat2aa_bwd(R.shape[0], _ext.ptr(R), _ext.ptr(daa), _ext.ptr(dR), _ext.stream_ptr()),
                   'humor_mat2aa_bwd')
        _ext.LaunchCounter.total += 1
        return dR


def batch_rodrigues(rot_vecs):
    """(N,3) axis-angle -> (N,3,3).  transforms.py:139-170."""
    return _Rodrigues.apply(rot_vecs.reshape(-1, 3))


def rotation_matrix_to_angle_axis(rotation_matrix):
    """(N,3,3) -> (N,3).  transforms.py:243-267 (quaternion route, NaN -> 0)."""
    return _Mat2AA.apply(rotation_matrix.reshape(-1, 3, 3))


def compute_world2aligned_mat(rot_pos):
    """transforms.py:17-42 — yaw aligning the body-right axis with +x.  Only used outside the fused
    rollout (canonicalize_input / Stage-III initialisation), so it stays a few torch ops + our Rodrigues."""
    right = -rot_pos[:, :, 0]
    xproj = right[:, 0:1] / (torch.norm(right[:, :2], dim=1, keepdim=True) + 1e-6)
    ang = torch.acos(torch.clamp(xproj, -1.0, 1.0))
    az = -right[:, 1:2]
    zeros = torch.zeros_like(az)
    axis = torch.cat([zeros, zeros, az], 1)
    aa = axis / (torch.norm(axis, dim=1, keepdim=True) + 1e-6) * ang
    return batch_rodrigues(aa)
