"""Camera <-> prior-frame geometry of the fitting path (humor/fitting/fitting_utils.py:61-103,
:149-190, :678-682).  Per-sequence (B rows, once per closure): compute_cam2prior on the optimised (B,3) floor is ONE kernel
forward and one in reverse (csrc/rot.cu: as torch ops it was ~60 launches forward and ~150 autograd nodes in reverse, the largest
group of tiny launches in the Stage-III graph); the torch form below remains for the already-parsed (B,4) plane of the cold paths."""
import torch

from . import _ext
from .transforms import batch_rodrigues

OP_NUM_JOINTS = 25
OP_IGNORE_JOINTS = [1, 9, 12]
OP_EDGE_LIST = [[1, 8], [1, 2], [1, 5], [2, 3], [3, 4], [5, 6], [6, 7], [8, 9], [9, 10], [10, 11], [8, 12], [12, 13],
                [13, 14], [1, 0], [0, 15], [15, 17], [0, 16], [16, 18], [14, 19], [19, 20], [14, 21], [11, 22],
                [22, 23], [11, 24]]
NSTAGES = 3


def parse_floor_plane(floor_plane):
    """(B,3) normal*offset -> (B,4) (a,b,c,d) with the normal pointing up in camera space (-y)."""
    d = floor_plane.norm(dim=1, keepdim=True)
    n = floor_plane / d
    flip = n[:, 1:2] > 0.0
    sgn = torch.where(flip, -torch.ones_like(d), torch.ones_like(d))
    return torch.cat([n * sgn, d * sgn], 1)


def compute_plane_intersection(point, direction, plane):
    """s with point + s*direction on the plane n.x = d (either sign of s)."""
    n, d = plane[:, :3], plane[:, 3]
    s = (d - (n * point).sum(-1)) / (n * direction).sum(-1)
    return point + s[:, None] * direction, s


class _Cam2Prior(torch.autograd.Function):
    """humor_cam2prior_fwd / _bwd: (floor (B,3), trans0 (B,3), root_orient0 (B,3), root joint (B,3)) -> (R (B,3,3), t (B,3), h (B,1))."""

    @staticmethod
    def forward(ctx, floor, trans0, orient0, joint0):
        _ext.require_cuda(floor, trans0, orient0, joint0)
        floor, trans0, orient0, joint0 = (_ext.f32c(x) for x in (floor, trans0, orient0, joint0))
        B = floor.shape[0]
        R = torch.empty(B, 3, 3, device=floor.device, dtype=torch.float32)
        t = torch.empty(B, 3, device=floor.device, dtype=torch.float32)
        h = torch.empty(B, 1, device=floor.device, dtype=torch.float32)
        _ext.check(_ext.lib().humor_cam2prior_fwd(B, _ext.ptr(floor), _ext.ptr(trans0), 3, _ext.ptr(orient0), 3, _ext.ptr(joint0), 3,
                                                  _ext.ptr(R), _ext.ptr(t), _ext.ptr(h), _ext.stream_ptr()), 'humor_cam2prior_fwd')
        _ext.LaunchCounter.total += 1
        ctx.save_for_backward(floor, trans0, orient0, joint0)
        return R, t, h

    @staticmethod
    def backward(ctx, gR, gt, gh):
        floor, trans0, orient0, joint0 = ctx.saved_tensors
        B = floor.shape[0]
        gR, gt, gh = (None if g is None else _ext.f32c(g) for g in (gR, gt, gh))
        out = torch.empty(4, B, 3, device=floor.device, dtype=torch.float32)
        _ext.check(_ext.lib().humor_cam2prior_bwd(B, _ext.ptr(floor), _ext.ptr(trans0), 3, _ext.ptr(orient0), 3, _ext.ptr(joint0), 3,
                                                  _ext.ptr(gR), _ext.ptr(gt), _ext.ptr(gh), _ext.ptr(out[0]), _ext.ptr(out[1]),
                                                  _ext.ptr(out[2]), _ext.ptr(out[3]), _ext.stream_ptr()), 'humor_cam2prior_bwd')
        _ext.LaunchCounter.total += 1
        return out[0], out[1], out[2], out[3]


def compute_cam2prior(floor_plane, trans, root_orient, joints):
    """Rotation/translation from the camera frame to the prior's canonical frame and the root height
    above the floor (fitting_utils.py:149-190): up = floor normal, right = body -x projected on the floor."""
    if floor_plane.size(1) == 3:
        return _Cam2Prior.apply(floor_plane, trans, root_orient, joints[:, 0])
    return compute_cam2prior_torch(floor_plane, trans, root_orient, joints)


def compute_cam2prior_torch(floor_plane, trans, root_orient, joints):
    """The same in torch ops (the form round 1 ran everywhere); kept for a plane that arrives parsed, (B,4), and as the
    autograd cross-check of the kernel pair in the tests."""
    plane = parse_floor_plane(floor_plane) if floor_plane.size(1) == 3 else floor_plane
    up = plane[:, :3]
    foot, _ = compute_plane_intersection(trans, -up, plane)
    body_right = -batch_rodrigues(root_orient)[:, :, 0]
    hit, s = compute_plane_intersection(trans, body_right, plane)
    right = (hit - foot) * torch.where(s < 0, -1.0, 1.0)[:, None]
    right = right / right.norm(dim=1, keepdim=True)
    fwd = torch.linalg.cross(up, right, dim=1)
    fwd = fwd / fwd.norm(dim=1, keepdim=True)
    R = torch.stack([right, fwd, up], 1)           # rows = prior axes expressed in the camera frame
    _, s_root = compute_plane_intersection(joints[:, 0], -up, plane)
    return R, -trans, s_root[:, None]


# ------------------------------------------------------------------------------------------------
# robust weighting of the point-cloud residuals (fitting_utils.py:192-248)
# ------------------------------------------------------------------------------------------------
def robust_std(res):
    """MAD / 0.67449 per row of res (B,N) (fitting_utils.py:213-228; torch.median = lower median)."""
    B = res.size(0)
    med = torch.median(res, dim=-1)[0].reshape((B, 1))
    mad = torch.median(torch.abs(res - med), dim=-1)[0].reshape((B, 1))
    return mad / 0.67449


def bisquare_robust_weights(res, tune_const=4.6851):
    """Tukey bisquare weights (fitting_utils.py:230-248).  `torch.where` instead of the reference's boolean-mask
    assignment: same values, no host synchronisation (CUDA-graph capturable)."""
    norm_res = res / (robust_std(res) * tune_const)
    w = (1.0 - norm_res ** 2) ** 2
    return torch.where(norm_res >= 1.0, torch.zeros_like(w), w)


class _SqrtSquare(torch.autograd.Function):
    """sqrt(d)**2 with the exact derivative 1.  The reference forms `obs2pred_sqr_dist.sqrt()` and squares it again
    (fitting_loss.py:389-391, fitting_utils.py:210): autograd then gives 2*sqrt(d)*(1/(2*sqrt(d))), NaN at d == 0
    (an observed point lying exactly on a vertex).  Forward value identical, gradient finite."""

    @staticmethod
    def forward(ctx, d):
        r = torch.sqrt(d)
        return r * r

    @staticmethod
    def backward(ctx, g):
        return g


def apply_robust_weighting_sq(sqr_res, robust_loss_type='bisquare', robust_tuning_const=4.6851):
    """apply_robust_weighting (fitting_utils.py:192-211) taking the SQUARED residuals the chamfer search returns:
    (w * sqrt(d)**2, w), the weights detached exactly as in the reference."""
    if robust_loss_type not in ('none', 'bisquare'):
        raise ValueError('Not a valid robust loss: %s' % robust_loss_type)
    with torch.no_grad():
        res = torch.sqrt(sqr_res)
        w = torch.ones_like(res) if robust_loss_type == 'none' else bisquare_robust_weights(res, robust_tuning_const)
    return w * _SqrtSquare.apply(sqr_res), w
