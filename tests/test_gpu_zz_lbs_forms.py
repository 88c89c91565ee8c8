"""GPU parity of the kernel forms of the dense tensor-core LBS forward (humor_lbs_configure): the default (3, 5) = fused blend +
group skinning on fp16 hi/lo planes (csrc/lbs_fuseg.cuh), the same kernel on 3xTF32 planes (3, 1), and the two-kernel form (1, 1)
of round 1.  Every form must agree with (1, 1) to fp32 rounding and with the CPU oracle far inside the 1e-4 m bound.  (Forms 2,
blend 3 and blend 4 were measured on the B200 in rounds 1-2 and removed: profiles/r01*, r02a*, r02f*.)"""
import numpy as np
import pytest
import torch

from humor_b200 import synth, _ext

# First hardware run: round 2, call r02a (profiles/r02a_gpu_tests_ungated.txt).
pytestmark = pytest.mark.gpu
DEFAULT = (3, 5)


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


@pytest.mark.parametrize('skin,blend,slab', [(3, 1, 512), (3, 5, 512), (1, 1, 256)])
@pytest.mark.parametrize('n', [300, 1100])
def test_forms_agree_with_the_two_kernel_form(bm, skin, blend, slab, n):
    ro, pb, be, tr = rand_pose(n, n)                     # 300: ragged row tile / frame block; 1100: 3 slabs
    try:
        configure(1, 1)
        ref = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        configure(skin, blend, slab)
        got = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (skin, blend)             # the requested kernels really ran (no silent fall-back)
    finally:
        configure(*DEFAULT)
    assert torch.isfinite(got.v).all()
    if skin == 3:
        assert not torch.equal(got.v, ref.v)             # a different summation order must show in the last bits
    assert float((got.v - ref.v).abs().max()) < 5e-6
    assert float((got.Jtr - ref.Jtr).abs().max()) < 5e-6


@pytest.mark.parametrize('skin,blend', [(1, 1), (3, 1), (3, 5)])
def test_forms_match_oracle(bm, skin, blend):
    from oracle.smplh_lbs import OracleBodyModel
    ob = OracleBodyModel(synth.make_smplh_asset(), use_vtx_selector=True)
    n = 200
    ro, pb, be, tr = rand_pose(n, 5)
    o = ob(root_orient=ro.cpu(), pose_body=pb.cpu(), betas=be.cpu(), trans=tr.cpu())
    try:
        configure(skin, blend)
        g = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (skin, blend)
    finally:
        configure(*DEFAULT)
    assert float((g.v.cpu() - o.v).abs().max()) < 2e-5
    assert float((g.Jtr.cpu() - o.Jtr).abs().max()) < 2e-5


@pytest.mark.parametrize('blend', [1, 5])
def test_fused_group_form_over_many_row_tiles(bm, blend):
    """skin form 3 at a size where every persistent CTA walks several column tiles and most cross a row-tile boundary."""
    n = 128 * 9 + 5
    ro, pb, be, tr = rand_pose(n, 21)
    try:
        configure(1, 1)
        ref = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        configure(3, blend)
        got = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        again = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
        torch.cuda.synchronize()
        assert forms_used() == (3, blend)
    finally:
        configure(*DEFAULT)
    assert torch.isfinite(got.v).all() and torch.equal(got.v, again.v)       # deterministic run to run
    assert float((got.v - ref.v).abs().max()) < 5e-6


def test_configure_rejects_bad_and_removed_values():
    L = _ext.lib()
    for skin, blend, slab in [(4, 0, 0), (2, 0, 0), (0, 2, 0), (0, 3, 0), (0, 4, 0), (0, 6, 0), (0, 0, 64)]:
        assert L.humor_lbs_configure(skin, blend, slab) != 0, (skin, blend, slab)
    assert L.humor_lbs_configure(0, 0, 0) == 0
    a = forms_used()
    assert L.humor_lbs_configure(*DEFAULT, 0) == 0 and a is not None


def test_shaped_template_path_matches_oracle_and_the_all_columns_path(bm):
    """One shape per T >= 32 frames (what MotionOptimizer passes: frames_per_beta = T): template + shape blend are added per SEQUENCE
    by the fused kernel's epilogue and the GEMM carries the 189 pose columns only (lbs_shape_rows_kernel + K = 192; five launches).
    Same vertices as the oracle, and as the all-columns path (K = 256, four launches) that a per-frame betas call takes - a ragged
    last row tile, a 128-frame tile that spans four sequences."""
    from humor_b200.body_model import lbs
    from oracle.smplh_lbs import OracleBodyModel
    B, T = 7, 40
    N = B * T
    ro, pb, _, tr = rand_pose(N, 9)
    be = torch.tensor((np.random.RandomState(10).randn(B, 16) * 0.7).astype(np.float32)).cuda()
    be_rep = be.repeat_interleave(T, 0).contiguous()
    ob = OracleBodyModel(synth.make_smplh_asset(), use_vtx_selector=True)
    o = ob(root_orient=ro.cpu(), pose_body=pb.cpu(), betas=be_rep.cpu(), trans=tr.cpu())
    configure(*DEFAULT)
    with torch.no_grad():
        n0 = _ext.LaunchCounter.total
        v, _, J = lbs(bm.lbs_model, ro, pb, be, tr, T, None, True, False, 73)
        n1 = _ext.LaunchCounter.total
        v2, _, J2 = lbs(bm.lbs_model, ro, pb, be_rep, tr, 1, None, True, False, 73)
        n2 = _ext.LaunchCounter.total
    torch.cuda.synchronize()
    assert forms_used() == DEFAULT
    assert (n1 - n0, n2 - n1) == (5, 4)                  # pose, [shaped templates,] fp16 feature planes, fused kernel, joint gather
    assert torch.isfinite(v).all() and float((v.cpu() - o.v).abs().max()) < 2e-5 and float((J.cpu() - o.Jtr).abs().max()) < 2e-5
    assert not torch.equal(v, v2) and float((v - v2).abs().max()) < 5e-6
