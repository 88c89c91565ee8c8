"""GPU parity of the alternative kernel forms of the dense tensor-core LBS forward (humor_lbs_configure):
skin_form 2 = lane-per-frame over vertex groups (csrc/lbs_skin_group.cuh; the same source runs on the CPU through the
SIMT shim in tests/test_host_lbs_skin.py), blend_form 2 = persistent 128x256-tile tcgen05 kernel (csrc/lbs_blend.cuh).
Every combination must agree with the default forms to fp32 rounding and with the CPU oracle to the 1e-4 m bound."""
import numpy as np
import pytest
import torch

from humor_b200 import synth, _ext

import os

# First hardware run: round 2, call r02a (profiles/r02a_gpu_tests_ungated.txt): all 18 cases green on the B200.
pytestmark = pytest.mark.gpu


def forms_used():
    import ctypes as C
    a, b = C.c_int(0), C.c_int(0)
    _ext.lib().humor_lbs_forms_used(C.byref(a), C.byref(b))
    return a.value, b.value


@pytest.fixture(scope='module')
def bm():
    from humor_b200.body_model import BodyModel
    return BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=1, use_vtx_selector=True).to('cuda')


def rand_pose(n, seed):
    rng = np.random.RandomState(seed)
    return [torch.tensor(a).cuda() for a in (rng.randn(n, 3).astype(np.float32) * 0.8, (rng.randn(n, 63) * 0.4).astype(np.float32),
                                             (rng.randn(n, 16) * 0.7).astype(np.float32), rng.randn(n, 3).astype(np.float32))]


def configure(skin, blend, slab=512):
    _ext.check(_ext.lib().humor_lbs_configure(skin, blend, slab), 'humor_lbs_configure')


@pytest.mark.parametrize('skin,blend,slab', [(2, 1, 512), (1, 2, 512), (2, 2, 512), (2, 2, 256), (3, 1, 512), (3, 5, 512)])   # skin 3 = fused (lbs_fuseg.cuh); blend 5 = fp16 hi/lo planes
@pytest.mark.parametrize('n', [300, 1100])
def test_forms_agree_with_default(bm, skin, blend, slab, n):
    ro, pb, be, tr = rand_pose(n, n)                     # 300: ragged row tile / frame block; 1100: 3 slabs
    configure(1, 1)
    ref = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
    try:
        configure(skin, blend, slab)
        got = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (skin, blend)             # the requested kernels really ran (no silent fall-back)
    finally:
        configure(1, 1)
    assert torch.isfinite(got.v).all()
    assert not torch.equal(got.v, ref.v)                 # a different summation order must show in the last bits
    assert float((got.v - ref.v).abs().max()) < 5e-6
    assert float((got.Jtr - ref.Jtr).abs().max()) < 5e-6


def test_forms_match_oracle(bm):
    from oracle.smplh_lbs import OracleBodyModel
    ob = OracleBodyModel(synth.make_smplh_asset(), use_vtx_selector=True)
    n = 200
    ro, pb, be, tr = rand_pose(n, 5)
    o = ob(root_orient=ro.cpu(), pose_body=pb.cpu(), betas=be.cpu(), trans=tr.cpu())
    try:
        configure(2, 2)
        g = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (2, 2)
    finally:
        configure(1, 1)
    assert float((g.v.cpu() - o.v).abs().max()) < 2e-5
    assert float((g.Jtr.cpu() - o.Jtr).abs().max()) < 2e-5


@pytest.mark.parametrize('skin,blend', [(2, 3), (3, 3), (3, 4)])          # (3, 4): the pose columns as fp16 planes
def test_mixed_precision_blend_stays_inside_the_vertex_bound(bm, skin, blend):
    """blend form 3 (one TF32 pass on the pose-offset k-blocks): <= 1e-4 m against the oracle, visibly different from form 1."""
    from oracle.smplh_lbs import OracleBodyModel
    ob = OracleBodyModel(synth.make_smplh_asset(), use_vtx_selector=True)
    n = 300
    ro, pb, be, tr = rand_pose(n, 9)
    o = ob(root_orient=ro.cpu(), pose_body=pb.cpu(), betas=be.cpu(), trans=tr.cpu())
    try:
        configure(skin, blend)
        g = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (skin, blend)
    finally:
        configure(1, 1)
    err = float((g.v.cpu() - o.v).abs().max())
    assert 1e-6 < err < 1e-4, err


def test_fused_group_form_over_many_row_tiles(bm):
    """skin form 3 at a size where every persistent CTA walks several column tiles and most cross a row-tile boundary."""
    n = 128 * 9 + 5
    ro, pb, be, tr = rand_pose(n, 21)
    configure(1, 1)
    ref = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
    try:
        configure(3, 1)
        got = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        again = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (3, 1)
    finally:
        configure(1, 1)
    assert torch.isfinite(got.v).all() and torch.equal(got.v, again.v)       # deterministic run to run
    assert float((got.v - ref.v).abs().max()) < 5e-6


def test_configure_rejects_bad_values():
    L = _ext.lib()
    assert L.humor_lbs_configure(4, 0, 0) != 0 and L.humor_lbs_configure(0, 7, 0) != 0 and L.humor_lbs_configure(0, 6, 0) != 0 and L.humor_lbs_configure(0, 0, 64) != 0
    assert L.humor_lbs_configure(0, 0, 0) == 0
