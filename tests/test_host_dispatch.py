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


def probe(mock, skin, blend, N):
    stub, lib = mock
    env = dict(os.environ, LD_PRELOAD=stub, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'mock', 'dispatch_probe.py'), ROOT, lib, str(skin), str(blend), str(N)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.startswith('LAUNCH') or l.startswith('{')]
    assert lines and lines[-1].startswith('{'), r.stdout[-2000:]
    info = json.loads(lines[-1])
    kernels = [l.split()[1] if not l.split()[1] == 'void' else l.split()[2] for l in lines[:-1]]
    grids = [l[l.index('grid='):] for l in lines[:-1]]
    return info, kernels, grids


def count(kernels, name):
    return sum(name in k for k in kernels)


@pytest.mark.parametrize('skin,blend', [(1, 1), (2, 1), (1, 2), (2, 2)])
def test_dense_forward_launches_the_requested_forms(mock, skin, blend):
    N = 1100                                                   # 2 full 512-frame slabs + 76 frames
    info, k, grids = probe(mock, skin, blend, N)
    assert info['rc_cfg'] == 0 and info['rc'] == 0
    assert info['used'] == [skin, blend], (info, k[:8])
    assert count(k, 'lbs_pose_warp_kernel') == 1 and count(k, 'lbs_gather_extra_kernel') == 1
    assert count(k, 'lbs_skin_group_kernel') == (3 if skin == 2 else 0)
    assert count(k, 'lbs_skin_apply_kernel') == (3 if skin == 1 else 0)
    assert count(k, 'lbs_blend_kernel') == (3 if blend == 2 else 0)
    assert count(k, 'umma_gemm3_kernel<128') == (3 if blend == 1 else 0)
    assert info['launches'] == len(k) == 8
    if blend == 2:                                             # persistent: min(tiles, SMs) CTAs, 81 column tiles of 256
        bg = [g for kk, g in zip(k, grids) if 'lbs_blend_kernel' in kk]
        assert bg[0].startswith('grid=(148,1,1)') and bg[2].startswith('grid=(81,1,1)')
    if skin == 2:                                              # 16 frame blocks x ~2 blocks per SM
        sg = [g for kk, g in zip(k, grids) if 'lbs_skin_group_kernel' in kk]
        assert sg[0].startswith('grid=(18,16,1)') and 'smem=80384' in sg[0]


@pytest.mark.parametrize('blend,used_blend', [(1, 1), (2, 1), (3, 3), (4, 4), (5, 5)])
def test_skin_form_3_is_one_persistent_kernel_for_all_frames(mock, blend, used_blend):
    """fused blend + group skinning (lbs_fuseg.cuh): no slabs, no v_posed round trip - pose kernel, ONE persistent launch of
    min(tiles, SMs) CTAs x 576 threads (TMA, MMA and 16 skinning warps) with 208 000 B of shared memory, joint gather."""
    info, k, grids = probe(mock, 3, blend, 1100)
    assert info['rc_cfg'] == 0 and info['rc'] == 0
    assert info['used'] == [3, used_blend], (info, k)
    assert info['launches'] == len(k) == (4 if blend >= 4 else 3)           # forms 4, 5: + the fp16 plane(s) of the features
    assert count(k, 'lbs_pose_warp_kernel') == 1 and count(k, 'lbs_fuseg_kernel') == 1 and count(k, 'lbs_gather_extra_kernel') == 1
    assert count(k, 'feat_f16_kernel') == (1 if blend >= 4 else 0)
    fg = [g for kk, g in zip(k, grids) if 'lbs_fuseg_kernel' in kk][0]
    assert fg.startswith('grid=(148,1,1)') and 'block=576' in fg and 'smem=208000' in fg, fg


def test_short_batches_stay_on_the_ffma_path(mock):
    info, k, _ = probe(mock, 2, 2, 64)                         # < 128 frames: no tensor-core path, forms irrelevant
    assert info['rc'] == 0 and count(k, 'lbs_skin_fwd_kernel') >= 1 and count(k, 'lbs_blend_kernel') == 0
