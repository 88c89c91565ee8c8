"""Host dispatch of the C-ABI under a fake CUDA runtime (tests/host/mock/cudart_stub.cpp, LD_PRELOADed into a child
process): which kernels humor_lbs_fwd launches, with which grids, for each form of humor_lbs_configure — on the REAL packed
model constants.  (A layout guard once made a requested form fall back silently on the B200; this is the test for that.)"""
import glob
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='module')
def mock(built_lib, tmp_path_factory):
    d = tmp_path_factory.mktemp('mockrt')
    stub = str(d / 'libcudart_stub.so')
    lib = str(d / 'libhumor_b200_mock.so')
    subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', os.path.join(HERE, 'host', 'mock', 'cudart_stub.cpp'), '-o', stub, '-ldl'])
    objs = sorted(glob.glob(os.path.join(ROOT, 'humor_b200', 'build', '*.o')))
    assert objs, 'object files of the library are missing (build() keeps them in humor_b200/build)'
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    subprocess.check_call([nvcc, '-shared', '-cudart', 'shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return stub, lib


def probe(mock, skin, blend, N, *extra):
    stub, lib = mock
    env = dict(os.environ, LD_PRELOAD=stub, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'mock', 'dispatch_probe.py'), ROOT, lib, str(skin), str(blend), str(N)] +
                       [str(e) for e in extra],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.startswith('LAUNCH') or l.startswith('{')]
    assert lines and lines[-1].startswith('{'), r.stdout[-2000:]
    info = json.loads(lines[-1])
    kernels = [l.split()[1] if not l.split()[1] == 'void' else l.split()[2] for l in lines[:-1]]
    grids = [l[l.index('grid='):] for l in lines[:-1]]
    return info, kernels, grids


def count(kernels, name):
    return sum(name in k for k in kernels)


def test_two_kernel_form_runs_slab_by_slab(mock):
    """forms (1, 1): blend GEMM into a v_posed slab that stays in L2 + lane = vertex skin pass, per 512-frame slab."""
    N = 1100                                                   # 2 full 512-frame slabs + 76 frames
    info, k, grids = probe(mock, 1, 1, N)
    assert info['rc_cfg'] == 0 and info['rc'] == 0
    assert info['used'] == [1, 1], (info, k[:8])
    assert count(k, 'lbs_pose_warp_kernel') == 1 and count(k, 'lbs_gather_extra_kernel') == 1
    assert count(k, 'lbs_skin_apply_kernel') == 3 and count(k, 'umma_gemm3p_kernel<0>') == 3      # persistent 128x128-tile blend GEMM
    assert info['launches'] == len(k) == 8


@pytest.mark.parametrize('skin,blend', [(2, 1), (1, 2), (3, 3), (3, 4), (4, 1), (1, 6)])
def test_removed_forms_are_refused(mock, skin, blend):
    """forms 2 (skin, blend), blend 3 and 4 were measured and removed in round 2: humor_lbs_configure says so instead of silently
    running something else."""
    info, _, _ = probe(mock, skin, blend, 1100)
    assert info['rc_cfg'] != 0


@pytest.mark.parametrize('blend', [1, 5])
def test_skin_form_3_is_one_persistent_kernel_for_all_frames(mock, blend):
    """fused blend + group skinning (lbs_fuseg.cuh): no slabs, no v_posed round trip - pose kernel, ONE persistent launch of
    min(tiles, SMs) CTAs x 608 threads (two TMA warps, MMA and 16 skinning warps) with 229 120 B of shared memory, joint gather."""
    info, k, grids = probe(mock, 3, blend, 1100)
    assert info['rc_cfg'] == 0 and info['rc'] == 0
    assert info['used'] == [3, blend], (info, k)
    # form 5: + the fp16 planes of the features + (60 frames per shape) the shaped template of every sequence
    assert info['launches'] == len(k) == (5 if blend == 5 else 3)
    assert count(k, 'lbs_pose_warp_kernel') == 1 and count(k, 'lbs_fuseg_kernel') == 1 and count(k, 'lbs_gather_extra_kernel') == 1
    assert count(k, 'feat_f16_kernel') == count(k, 'lbs_shape_rows_kernel') == (1 if blend == 5 else 0)
    fg = [g for kk, g in zip(k, grids) if 'lbs_fuseg_kernel' in kk][0]
    assert fg.startswith('grid=(148,1,1)') and 'block=608' in fg and 'smem=229120' in fg, fg


def test_per_frame_shapes_keep_every_column_in_the_product(mock):
    """frames_per_beta < 32: no shaped-template rows (a 128-frame tile could span more sequences than the kernel stages) - the fused
    kernel runs on the K = 256 planes, four launches."""
    info, k, _ = probe(mock, 3, 5, 1100, 1)
    assert info['rc'] == 0 and info['used'] == [3, 5] and info['launches'] == len(k) == 4
    assert count(k, 'lbs_shape_rows_kernel') == 0 and count(k, 'feat_f16_kernel') == 1 and count(k, 'lbs_fuseg_kernel') == 1


def test_model_without_skinning_records_takes_the_two_kernel_form(mock):
    """HbLbsModel.ft_rec absent (body_model.fuseg_records returns None for a mesh whose column tiles need more entries than a record
    buffer holds): the fused kernel is not eligible and humor_lbs_fwd runs skin form 1 - and says so."""
    info, k, _ = probe(mock, 3, 5, 1100, 60, 'norec')
    assert info['rc'] == 0 and info['used'] == [1, 1]
    assert count(k, 'lbs_fuseg_kernel') == 0 and count(k, 'lbs_skin_apply_kernel') == 3


def test_short_batches_stay_on_the_ffma_path(mock):
    info, k, _ = probe(mock, 3, 5, 64)                         # < 128 frames: no tensor-core path, forms irrelevant
    assert info['rc'] == 0 and count(k, 'lbs_skin_fwd_kernel') >= 1 and count(k, 'lbs_fuseg_kernel') == 0


def test_asset_outside_fp16_range_withholds_the_fp16_planes():
    """body_model.LbsModel._build: blend planes scaled by 2^10 must fit fp16; an asset in other units (a template in millimetres) gets no
    fp16 planes - the C-ABI then runs the fused kernel on the tf32 planes (blend form 1) - while SMPL+H in metres gets them."""
    import numpy as np
    import torch
    from humor_b200 import synth
    from humor_b200 import body_model as BM

    class Probe(BM.LbsModel):
        def __init__(self, packed):            # host memory: only the tables / planes / struct are under test
            self._build(packed, 'cpu')
    asset = synth.make_smplh_asset()
    ok = Probe(BM.pack_smplh(asset, 16)).struct
    assert ok.blend16a_h and ok.blend16a_l and ok.blend16p_h and ok.blend16p_l and (ok.flags & 1)
    big = dict(asset)
    big['v_template'] = np.asarray(asset['v_template']) * 1000.0
    far = Probe(BM.pack_smplh(big, 16)).struct
    assert not far.blend16a_h and not far.blend16p_h and far.blend_t_hi and far.ft_rec
