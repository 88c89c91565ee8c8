"""Loader for the C-ABI shared library (include/humor_b200.h) — ctypes, no torch types cross it.

The library is built in-tree (``humor_b200/libhumor_b200.so``) by ``build()`` /
``__graft_entry__.build()`` with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libhumor_b200.so')
SOURCES = ['rollout.cu', 'lbs.cu', 'rot.cu', 'losses.cu', 'umma_gemm.cu', 'chamfer.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']

HB_NUM_TERMS = 24
TERM_NAMES = ['joints2d', 'joints3d', 'verts3d', 'rgb_overlap_consist_verts3d_pos',
              'rgb_overlap_consist_verts3d_vel', 'pose_prior', 'shape_prior', 'joints3d_smooth',
              'rgb_overlap_consist_betas', 'motion_prior', 'init_motion_prior', 'joint_consistency',
              'bone_length', 'joints3d_rollout', 'contact_vel', 'contact_height', 'floor_reg',
              'rgb_overlap_consist_floor']
TERM = {n: i for i, n in enumerate(TERM_NAMES)}


def _sources_newer_than_lib():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + \
           [os.path.join(_HERE, '..', 'include', 'humor_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library (cross-compiles without a GPU)."""
    if not force and not _sources_newer_than_lib():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, 'build'), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(_HERE, 'build', s.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(_CSRC, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {s}:\n{out}')
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    return LIB_PATH


class HbLbsModel(C.Structure):
    _fields_ = [('num_verts', C.c_int), ('v3_ld', C.c_int), ('wk', C.c_int), ('flags', C.c_int),
                ('v_template', C.c_void_p), ('blend', C.c_void_p), ('blend_t', C.c_void_p),
                ('j_template', C.c_void_p), ('j_dirs', C.c_void_p), ('w_idx', C.c_void_p),
                ('w_val', C.c_void_p), ('parents', C.c_void_p), ('extra_ids', C.c_void_p),
                ('blend_t_hi', C.c_void_p), ('blend_t_lo', C.c_void_p), ('use_umma', C.c_int), ('max_depth', C.c_int),
                ('depth', C.c_void_p), ('child_start', C.c_void_p), ('child_list', C.c_void_p),
                ('g_start', C.c_void_p), ('g_joint', C.c_void_p), ('g_w', C.c_void_p), ('num_groups', C.c_int),
                ('ft_nct', C.c_int), ('g_slot', C.c_void_p), ('ft_tab', C.c_void_p),
                ('blend16a_h', C.c_void_p), ('blend16a_l', C.c_void_p),
                ('sel_ids', C.c_void_p), ('sel_blend', C.c_void_p), ('sel_nv', C.c_int), ('ft_rec_stride', C.c_int),
                ('ft_rec', C.c_void_p), ('blend16p_h', C.c_void_p), ('blend16p_l', C.c_void_p)]


class HbHumorWeights(C.Structure):
    _fields_ = [('dec_w', C.c_void_p * 4), ('dec_b', C.c_void_p * 4), ('dec_g', C.c_void_p * 3),
                ('dec_be', C.c_void_p * 3), ('dec_wt', C.c_void_p * 4), ('pri_w', C.c_void_p * 5),
                ('pri_b', C.c_void_p * 5), ('pri_g', C.c_void_p * 4), ('pri_be', C.c_void_p * 4),
                ('pri_wt', C.c_void_p * 5), ('pri_w_hi', C.c_void_p * 5), ('pri_w_lo', C.c_void_p * 5),
                ('pri_wt_hi', C.c_void_p * 5), ('pri_wt_lo', C.c_void_p * 5), ('dec_w_hi', C.c_void_p * 4),
                ('dec_w_lo', C.c_void_p * 4), ('dec_wt_hi', C.c_void_p * 4), ('dec_wt_lo', C.c_void_p * 4),
                ('use_umma', C.c_int), ('reserved', C.c_int), ('dec_w16_h', C.c_void_p * 4), ('dec_w16_l', C.c_void_p * 4),
                ('pri_w16_h', C.c_void_p * 5), ('pri_w16_l', C.c_void_p * 5), ('dec_wz_hi', C.c_void_p), ('dec_wz_lo', C.c_void_p)]


class HbFitArgs(C.Structure):
    _fields_ = [('B', C.c_int), ('T', C.c_int), ('njx', C.c_int), ('coef', C.c_float * HB_NUM_TERMS),
                ('sigma2d', C.c_float),
                ('cam_joints', C.c_void_p), ('cam_verts', C.c_void_p), ('prior_joints', C.c_void_p),
                ('roll_joints', C.c_void_p), ('contact_logits', C.c_void_p), ('betas', C.c_void_p),
                ('floor', C.c_void_p), ('z', C.c_void_p), ('prior_out', C.c_void_p), ('latent_pose', C.c_void_p),
                ('obs_joints2d', C.c_void_p), ('obs_joints3d', C.c_void_p), ('obs_verts3d', C.c_void_p),
                ('obs_floor', C.c_void_p), ('seq_interval', C.c_void_p), ('cam_f', C.c_void_p), ('cam_c', C.c_void_p),
                ('T_obs', C.c_int),
                ('terms', C.c_void_p), ('loss', C.c_void_p),
                ('d_cam_joints', C.c_void_p), ('d_cam_verts', C.c_void_p), ('d_prior_joints', C.c_void_p),
                ('d_roll_joints', C.c_void_p), ('d_contact_logits', C.c_void_p), ('d_betas', C.c_void_p),
                ('d_floor', C.c_void_p), ('d_z', C.c_void_p), ('d_prior_out', C.c_void_p),
                ('d_latent_pose', C.c_void_p), ('partials', C.c_void_p)]


EXPORTS = ['humor_lbs_workspace_bytes', 'humor_lbs_fwd', 'humor_lbs_bwd', 'humor_rollout_workspace_bytes',
           'humor_rollout_fwd', 'humor_rollout_bwd', 'humor_rodrigues_fwd', 'humor_rodrigues_bwd',
           'humor_mat2aa_fwd', 'humor_mat2aa_bwd', 'humor_fit_losses', 'humor_gmm_nll', 'humor_b200_version',
           'humor_umma_gemm', 'humor_umma_gemm_workspace_bytes', 'humor_chamfer_fwd', 'humor_chamfer_bwd', 'humor_lbs_configure', 'humor_lbs_forms_used',
           'humor_umma_gemm16', 'humor_umma_gemm16_workspace_bytes', 'humor_chain_debug', 'humor_cam2prior_fwd', 'humor_cam2prior_bwd', 'humor_rollout_outputs_fwd', 'humor_rollout_outputs_bwd', 'humor_gmm_workspace_bytes', 'humor_gmm_nll_ws', 'humor_rollout_bwd_started_wait', 'humor_lbs_set_fuseg_ctas', 'humor_lbs_fuseg_timing']

_LIB = None


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(humor_b200 has no CPU fallback)')
    L = C.CDLL(LIB_PATH)
    vp, ci, sz, i64p = C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int64)
    L.humor_lbs_workspace_bytes.restype = sz
    L.humor_lbs_workspace_bytes.argtypes = [ci]
    L.humor_lbs_fwd.restype = ci
    L.humor_lbs_fwd.argtypes = [C.POINTER(HbLbsModel), ci, ci, vp, vp, vp, vp, vp, sz, vp, ci, vp, vp, ci, i64p, vp]
    L.humor_lbs_bwd.restype = ci
    L.humor_lbs_bwd.argtypes = [C.POINTER(HbLbsModel), ci, ci, vp, vp, vp, vp, vp, sz, vp, ci, vp, vp, ci,
                                vp, vp, vp, vp, i64p, vp]
    L.humor_rollout_workspace_bytes.restype = sz
    L.humor_rollout_workspace_bytes.argtypes = [ci, ci]
    L.humor_rollout_fwd.restype = ci
    L.humor_rollout_fwd.argtypes = [C.POINTER(HbHumorWeights), ci, ci, vp, vp, vp, sz, vp, vp, i64p, vp]
    L.humor_rollout_bwd.restype = ci
    L.humor_rollout_bwd.argtypes = [C.POINTER(HbHumorWeights), ci, ci, vp, sz, vp, vp, vp, vp, i64p, vp]
    for n in ('humor_rodrigues_fwd', 'humor_mat2aa_fwd'):
        getattr(L, n).restype = ci
        getattr(L, n).argtypes = [ci, vp, vp, vp]
    for n in ('humor_rodrigues_bwd', 'humor_mat2aa_bwd'):
        getattr(L, n).restype = ci
        getattr(L, n).argtypes = [ci, vp, vp, vp, vp]
    L.humor_cam2prior_fwd.restype = ci
    L.humor_cam2prior_fwd.argtypes = [ci, vp, vp, ci, vp, ci, vp, ci, vp, vp, vp, vp]
    L.humor_cam2prior_bwd.restype = ci
    L.humor_cam2prior_bwd.argtypes = [ci, vp, vp, ci, vp, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.humor_rollout_outputs_fwd.restype = ci
    L.humor_rollout_outputs_fwd.argtypes = [ci, ci] + [vp] * 8 + [C.c_float] + [vp] * 10
    L.humor_rollout_outputs_bwd.restype = ci
    L.humor_rollout_outputs_bwd.argtypes = [ci, ci] + [vp] * 19
    L.humor_fit_losses.restype = ci
    L.humor_fit_losses.argtypes = [C.POINTER(HbFitArgs), i64p, vp]
    L.humor_gmm_nll.restype = ci
    L.humor_gmm_nll.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.humor_gmm_workspace_bytes.restype = sz
    L.humor_gmm_workspace_bytes.argtypes = [ci, ci, ci]
    L.humor_gmm_nll_ws.restype = ci
    L.humor_gmm_nll_ws.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.humor_umma_gemm_workspace_bytes.restype = sz
    L.humor_umma_gemm_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.humor_umma_gemm.restype = ci
    L.humor_umma_gemm.argtypes = [vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, vp, sz, vp]
    L.humor_umma_gemm16_workspace_bytes.restype = sz
    L.humor_umma_gemm16_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.humor_umma_gemm16.restype = ci
    L.humor_umma_gemm16.argtypes = [vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, vp, sz, vp]
    L.humor_rollout_bwd_started_wait.restype = ci
    L.humor_rollout_bwd_started_wait.argtypes = [vp]
    L.humor_lbs_set_fuseg_ctas.restype = ci
    L.humor_lbs_set_fuseg_ctas.argtypes = [ci]
    L.humor_lbs_fuseg_timing.restype = ci
    L.humor_lbs_fuseg_timing.argtypes = [ci, C.POINTER(C.c_float), C.POINTER(ci)]
    L.humor_lbs_configure.restype = ci
    L.humor_lbs_configure.argtypes = [ci, ci, ci]
    L.humor_lbs_forms_used.restype = ci
    L.humor_lbs_forms_used.argtypes = [C.POINTER(ci), C.POINTER(ci)]
    L.humor_chamfer_fwd.restype = ci
    L.humor_chamfer_fwd.argtypes = [ci, ci, vp, ci, vp, vp, vp, vp, vp, i64p, vp]
    L.humor_chamfer_bwd.restype = ci
    L.humor_chamfer_bwd.argtypes = [ci, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp, i64p, vp]
    L.humor_chain_debug.restype = ci
    L.humor_chain_debug.argtypes = [vp, sz]
    L.humor_b200_version.restype = C.c_char_p
    _LIB = L
    return L


class LaunchCounter:
    """Counts the kernels our C-ABI calls enqueue (bench.py reports it as gpu_launches)."""
    total = 0


def check(rc, what):
    if rc != 0:
        msg = f'{what} failed with code {rc}'
        if 0 < rc < 1000:
            msg += ' (cudaError_t)'
        raise RuntimeError(msg)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('humor_b200 kernels need CUDA tensors (there is no CPU path)')


def f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
