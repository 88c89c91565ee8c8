"""The PRODUCT on CPU tensors: the library's own nvcc-compiled host code (argument checks, workspace carving, kernel dispatch)
linked with `-cudart shared`, run in a child process whose CUDA runtime is tests/host/emul/cudart_emul.cpp — every kernel
launch executes the same kernel source compiled with g++ against the SIMT shim.  Python side: the product's classes with the
"CUDA tensors only" guards patched out for that child process (tests/host/emul/cpu_backend.py).

* BodyModel forward / reverse against the torch oracle.
* Stage-I (root_fit) and Stage-II (smpl_fit) closures of MotionOptimizer against fixtures of the UNMODIFIED reference
  (RGB, AMASS key-vertex and PROX-RGBD/point-cloud configurations): loss, every energy term, every gradient."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='module')
def emul(built_lib, tmp_path_factory):
    d = tmp_path_factory.mktemp('emul')
    rt = str(d / 'libcudart_emul.so')
    lib = str(d / 'libhumor_b200_emul.so')
    subprocess.check_call(['g++', '-O1', '-std=c++20', '-pthread', '-shared', '-fPIC', '-I' + os.path.join(HERE, 'host', 'shim'),
                           '-DHB_HOST_SHIM', os.path.join(HERE, 'host', 'emul', 'cudart_emul.cpp'), '-o', rt, '-ldl'])
    objs = sorted(glob.glob(os.path.join(ROOT, 'humor_b200', 'build', '*.o')))
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    subprocess.check_call([nvcc, '-shared', '-cudart', 'shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return rt, lib


def run_probe(emul, script, *args, tensor=False, extra_env=None):
    rt, lib = emul
    env = dict(os.environ, LD_PRELOAD=rt, CUDA_VISIBLE_DEVICES='')
    env.setdefault('HB_CHAIN_CLUSTERS', '2')      # persistent decoder chain: 2 clusters x 4 CTAs x 192 threads run at once
    env.update(extra_env or {})
    if tensor:          # tcgen05 kernels on the functional emulation, incl. the 4-CTA-cluster split-K GEMMs of the decoder chain
        env.update(HB_EMUL_TENSOR='1')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'emul', script), ROOT, lib] + list(args),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert r.returncode == 0 and lines, (r.stdout[-1500:], r.stderr[-3000:])
    return json.loads(lines[-1])


def test_body_model_forward_and_reverse(emul):
    out = run_probe(emul, 'probe_lbs.py')
    assert out['v_err'] < 2e-5 and out['J_err'] < 2e-5          # metres (bound 1e-4)
    for k in ('g_root_orient', 'g_pose_body', 'g_betas', 'g_trans'):
        assert out[k] < 1e-4, (k, out[k])


@pytest.mark.parametrize('name', ['stage1_rgb', 'stage2_rgb', 'stage2_amass', 'stage2_proxd'])
def test_stage12_closure_matches_reference_golden(emul, name):
    out = run_probe(emul, 'probe_stage12.py', name)
    g = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    ref_loss = float(g['loss'])
    assert abs(out['loss'] - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    stats = {k[5:]: float(g[k]) for k in g.files if k.startswith('stat_')}
    assert set(stats) == set(out['stats']), (sorted(stats), sorted(out['stats']))
    for k, v in stats.items():
        assert abs(out['stats'][k] - v) <= 1e-4 * max(1.0, abs(v)), (k, out['stats'][k], v)
    assert out['verts_err'] < 1e-5
    for k, e in out['grad_err'].items():
        assert e < 1e-4, (k, e)


@pytest.mark.parametrize('name', ['stage3_rgb_phase1', 'stage3_rgb_xbatch',          # xbatch: prev_batch_overlap_res terms
                                  pytest.param('stage3_proxd', marks=pytest.mark.skipif(not os.environ.get('HB_SLOW_TESTS'),
                                               reason='1 min on the emulation (chamfer at PROX size): set HB_SLOW_TESTS=1'))])
def test_stage3_closure_matches_reference_golden(emul, name):
    """The WHOLE Stage-III closure of the product (VPoser decode, cam->prior, CVAE rollout forward + BPTT on the exact-fp32
    kernels, SMPL+H LBS, fused energies, GMM prior, chamfer/points3d for the PROX-RGBD case) on the CPU through the emulated
    kernels, against fixtures of the unmodified reference.  (stage3_rgb / stage3_amass / stage3_rgb_refine pass the same way;
    two cases keep the CPU suite short: `python tests/host/emul/probe_stage3.py` runs any of them.)"""
    out = run_probe(emul, 'probe_stage3.py', name)
    g = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    ref_loss = float(g['loss'])
    assert abs(out['loss'] - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    for k in g.files:
        if k.startswith('stat_'):
            v = float(g[k])
            assert abs(out['stats'][k[5:]] - v) <= 1e-4 * max(1.0, abs(v)), (k, out['stats'][k[5:]], v)
    assert out['verts_err'] < 1e-5 and out['trans_err'] < 1e-5 and out['prior_mean_err'] < 1e-5
    for k, e in out['grad_err'].items():
        assert e < 1e-4, (k, e)


# Tolerances of the END-TO-END comparison = ~10x the rounding-noise floor of the optimisation itself, measured by rebuilding the
# emulated kernels with FMA contraction (g++ -mfma -ffp-contract=fast, what nvcc does on the device): the same run then moves by
# trans 3e-6, pose_body 3e-6, stage3_verts3d 5e-6 and latent_motion 2.3e-4 (the latent is weakly determined: 7 L-BFGS iterations
# amplify last-bit differences).  Measured against the reference on the plain build: trans 6e-7 ... latent_motion 2e-5.
RUN_TOL = {'trans': 3e-5, 'root_orient': 3e-5, 'pose_body': 3e-5, 'betas': 2e-5, 'latent_pose': 3e-5, 'latent_motion': 2e-3,
           'floor_plane': 5e-5, 'stage3_verts3d': 5e-5, 'stage1_joints3d': 2e-5, 'stage2_joints3d': 2e-5}


def check_run_result(got, g):
    for k, tol in RUN_TOL.items():
        assert np.abs(got[k] - g[k]).max() <= tol, (k, float(np.abs(got[k] - g[k]).max()))
    assert np.array_equal(got['contacts'], g['contacts'])
    # Stage-III initialisation outputs and the per-stage npz dumps of run() (motion_optimizer.py:406-455, 650-674): same files,
    # same keys, same values as the reference wrote
    assert np.abs(got['stage3_init_joints3d'] - g['stage3_init_joints3d']).max() <= 3e-5
    names = [k for k in g.files if k.startswith('file_')]
    assert sorted(names) == sorted(k for k in got.keys() if k.startswith('file_'))
    for k in names:
        assert got[k].shape == g[k].shape, k
        if k.endswith('_contacts'):
            assert np.array_equal(got[k], g[k]), k
        else:
            assert np.abs(got[k] - g[k]).max() <= 5e-5, (k, float(np.abs(got[k] - g[k]).max()))


@pytest.mark.skipif(not os.environ.get('HB_SLOW_TESTS'), reason='~6 min on the CPU emulation: set HB_SLOW_TESTS=1')
def test_run_end_to_end_matches_reference_run(emul, tmp_path):
    """MotionOptimizer.run — Stage I, Stage II, Stage-III initialisation (posterior encoder, finite-difference velocities),
    Stage III with its three phases, L-BFGS with strong-Wolfe — of the product on the CPU emulation against the result of the
    UNMODIFIED reference's run() on the same problem (tests/golden/run_rgb.npz, oracle/make_golden_run.py).
    Measured: trans 6e-7, root_orient 1e-6, pose_body 1e-6, betas 3e-7, latent_motion 2e-5, contacts identical."""
    from oracle.make_golden_run import CFG
    out = str(tmp_path / 'run.npz')
    rt, lib = emul
    env = dict(os.environ, LD_PRELOAD=rt, CUDA_VISIBLE_DEVICES='')
    args = [str(x) for x in (CFG['B'], CFG['T'], CFG['seed'], *CFG['num_iter'], CFG['lbfgs_max_iter'])]
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'emul', 'probe_run.py'), ROOT, lib, out] + args,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=3000)
    assert r.returncode == 0, r.stderr[-3000:]
    check_run_result(np.load(out), np.load(os.path.join(HERE, 'golden', 'run_rgb.npz')))


@pytest.mark.skipif(not os.environ.get('HB_SLOW_TESTS'), reason='~5 min on the CPU emulation: set HB_SLOW_TESTS=1')
def test_sharded_run_matches_single_process_run(emul, tmp_path):
    """SURVEY.md 8(e) end to end: MotionOptimizer.run with the sub-sequences sharded over 2 gloo ranks — neighbour halo
    exchange of the overlap energies in all three stages + the joint L-BFGS (all-reduced inner products) — against the same
    run in one process.  Measured: every output within 8e-6 (trans 1e-6, latent_motion 1e-6, contacts identical)."""
    rt, lib = emul
    env = dict(os.environ, LD_PRELOAD=rt, CUDA_VISIBLE_DEVICES='')
    probe = os.path.join(HERE, 'host', 'emul', 'probe_run_dist.py')
    port = 24000 + (os.getpid() % 2000)
    tail = ['4', '6', '61', '1', '1', '2', '2']                  # B_total T seed num_iter x3 lbfgs_max_iter
    outs = [str(tmp_path / f'r{r}.npz') for r in range(2)] + [str(tmp_path / 'single.npz')]
    procs = [subprocess.Popen([sys.executable, probe, ROOT, lib, outs[r], str(r), '2', str(port)] + tail, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    procs.append(subprocess.Popen([sys.executable, probe, ROOT, lib, outs[2], '0', '1', '0'] + tail, env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        so, se = p.communicate(timeout=3000)
        assert p.returncode == 0, se[-3000:]
    r0, r1, single = (np.load(o) for o in outs)
    for k in single.files:
        both = np.concatenate([r0[k], r1[k]], 0)
        if k == 'contacts':
            assert np.array_equal(both, single[k])
        else:
            assert np.abs(both - single[k]).max() < 5e-5, (k, float(np.abs(both - single[k]).max()))


def test_cam2prior_kernel_pair_matches_the_torch_form(emul):
    """humor_cam2prior_fwd / _bwd (one launch each instead of ~60 torch ops forward and ~150 autograd nodes in reverse) against
    fitting_utils.compute_cam2prior_torch in float64 with autograd: both signs of the floor normal's y (the flip of
    parse_floor_plane), a near-identity root orientation, gradient of the root joint only.  Measured: forward 9e-7, gradients 1.4e-5
    relative (fp32 Rodrigues near the identity)."""
    out = run_probe(emul, 'probe_cam2prior.py')
    assert out['finite'] and out['joint_grad_only_root'] and out['orthonormal'] < 3e-6
    assert max(out['fwd']) < 5e-6 and max(out['bwd_rel']) < 1e-4, out
    assert out['parsed_plane_vs_kernel'] < 5e-6


def test_rollout_outputs_kernel_pair_matches_the_torch_form(emul):
    """humor_rollout_outputs_fwd / _bwd (world rows + frame-0 state -> the (B,T,.) tensors of the energies, prior and camera frame; one
    launch each way) against MotionOptimizer._rollout_outputs_torch with autograd: values identical (same device functions, same
    order), gradients to rounding; without a camera frame R / t receive no gradient."""
    out = run_probe(emul, 'probe_rollout_outputs.py')
    for tag in ('cam', 'nocam'):
        assert max(out[tag + '_fwd'].values()) < 1e-6, out[tag + '_fwd']
        assert max(v for v in out[tag + '_bwd_rel'].values() if v is not None) < 1e-5, out[tag + '_bwd_rel']
        assert out[tag + '_unused_grads_none']


def test_dense_lbs_kernel_forms_through_the_real_dispatch(emul):
    """humor_lbs_fwd's tensor-core path on the emulated tcgen05 kernels: the library's own dispatch AND its own TMA-descriptor
    code (cuTensorMapEncodeTiled is emulated) for the two-kernel form (1, 1) and the fused blend + group-skinning kernel on 3xTF32
    planes (3, 1) and on all-column fp16 planes (3, 5: one shape per FRAME here, so no shaped-template rows; the per-sequence form
    of the default goes through test_forms_verification_tool below).  All within 2e-5 m of the fp64 oracle."""
    out = run_probe(emul, 'probe_lbs_forms.py', '140', '11;31;35', tensor=True)
    assert out['exact_vs_oracle'] < 2e-5
    for key, want in (('forms_11', [1, 1]), ('forms_31', [3, 1]), ('forms_35', [3, 5])):
        f = out[key]
        assert f['used'] == want and f['finite'], (key, f)
        assert f['v_vs_oracle'] < 2e-5 and f['J_vs_oracle'] < 2e-5 and f['v_vs_exact'] < 5e-6, (key, f)


def test_umma_gemm_single_cta_and_split_k_cluster(emul):
    """humor_umma_gemm through the library's dispatch: few-tile shapes with K >= 256 take the split-K path (a 4-CTA thread-block
    cluster whose partial accumulators meet in the leader's shared memory through DSMEM stores between two cluster barriers;
    the emulation runs the four CTAs concurrently), the others the single-CTA tiles; ragged M / N / K-per-rank."""
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'emul', 'probe_umma.py'), ROOT, emul[1]], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900,
                       env=dict(os.environ, LD_PRELOAD=emul[0], CUDA_VISIBLE_DEVICES='', HB_EMUL_TENSOR='1', HB_EMUL_TRACE='1'))
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
    for c in out['cases']:
        assert c['rc'] == 0 and c['finite'] and c['rel_err'] < 3e-6, c
    launched = [l.split()[1] + ' ' + ' '.join(l.split()[2:4]) for l in r.stderr.splitlines() if l.startswith('EMUL hb::umma_gemm3_kernel')]
    assert sum(', 4>' in k for k in launched) == 1 and sum(', 1>' in k for k in launched) == 1, launched
    # fp16 hi + scaled lo operand planes (umma_gemm16.cuh): 22-bit significands -> errors of a few 1e-7 of sum |a||b|, also for
    # operands spread over decades and for weights around / below fp16's normal range; single-CTA, split-K cluster and 128-wide tiles
    for c in out['cases16']:
        assert c['rc'] == 0 and c['finite'] and c['rel_err'] < 1e-6, c
    l16 = [l for l in r.stderr.splitlines() if l.startswith('EMUL hb::umma_gemm16_kernel')]
    assert sum('<64, 0, 4>' in k for k in l16) == 1 and sum('<64, 0, 1>' in k for k in l16) == 1 and sum('<128, 0, 1>' in k for k in l16) == 1, l16


def test_forms_verification_tool(emul):
    """tools/lbs_forms_time.measure (what bench.py's `roofline_candidates` children run on the device) on the emulation: every
    form reports the kernels it really launched, agrees with form (1, 1) inside its tolerance, is deterministic, and differs
    from form (1, 1) in the last bits."""
    out = run_probe(emul, 'probe_forms_tool.py', '3,5', tensor=True)  # (3, 1) goes through the dispatch test above
    recs = {(r['skin'], r['blend']): r for r in out['recs']}
    assert set(recs) == {(3, 5)}
    for key, r in recs.items():
        assert r['verified'] and r['used'] == list(key) and r['deterministic'] and r['finite'] and r['frames'] == 129, r
        assert r['ms'] > 0 and r['GBps'] > 0 and 0 < r['frac'] < 1
    # form 5 (fp16 hi + lo planes, three products): back at the accuracy of the three-pass forms, from 4-byte operand elements
    assert not recs[(3, 5)]['bitwise_equal_to_11'] and recs[(3, 5)]['max_abs_diff_vs_11'] < 5e-6


def test_persistent_chain_matches_launch_per_layer_chain(emul):
    """csrc/chain_persist.cuh on the emulation (ALL clusters of the grid run concurrently: data-flow flags, remote mbarrier
    hand-shakes, DSMEM reduce-scatter) against the launch-per-layer chain: a batch of two ragged 128-row tiles, three resident
    clusters (tiles strided unevenly over clusters), forward states / prior and the reverse pass incl. the batched d z GEMM."""
    rt, lib = emul
    env = dict(os.environ, LD_PRELOAD=rt, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'host', 'emul', 'probe_chain.py'), ROOT, lib, '131', '3', '3'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=1500)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert r.returncode == 0 and lines, (r.stdout[-1500:], r.stderr[-3000:])
    out = json.loads(lines[-1])
    assert out['finite'] and out['world'] < 5e-6 and out['prior'] < 1e-5 and out['d_init'] < 1e-4 and out['d_z'] < 1e-4, out


def test_stage3_closure_tensor16_precision(emul):
    """precision 'tensor16' (opt-in): the FORWARD decoder chain on fp16 hi + scaled lo operand planes (4 bytes per element;
    csrc/umma_gemm16.cuh with its GroupNorm epilogues, chain16_pack_kernel for the step inputs), the tape, the batched prior and
    the whole reverse pass as in 'tensor' - against the same fixture of the unmodified reference, same tolerances."""
    out = run_probe(emul, 'probe_stage3.py', 'stage3_rgb_phase1', tensor=True, extra_env={'HB_EMUL_PRECISION': 'tensor16'})
    g = np.load(os.path.join(HERE, 'golden', 'stage3_rgb_phase1.npz'))
    assert abs(out['loss'] - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert out['trans_err'] < 1e-5 and out['prior_mean_err'] < 1e-5
    for k, e in out['grad_err'].items():
        assert e < 1e-3, (k, e)


@pytest.mark.parametrize('prefetch_b', [False, pytest.param(True, marks=pytest.mark.skipif(
    not os.environ.get('HB_SLOW_TESTS'), reason='second producer order of the same kernels: set HB_SLOW_TESTS=1'))])
def test_stage3_closure_tensor_precision(emul, prefetch_b):
    """The DEFAULT precision mode: every rollout GEMM on the (emulated) tcgen05 3xTF32 kernels with descriptors built by the
    library's own host code - the decoder chain as the PERSISTENT kernel of csrc/chain_persist.cuh (two resident clusters), the
    batched prior on umma_gemm3_kernel - against the fixture of the unmodified reference.  prefetch_b: the opt-in producer order of
    HB_UMMA_PREFETCH_B=1 (weight tiles requested before the programmatic-dependent-launch wait) - same numbers."""
    out = run_probe(emul, 'probe_stage3.py', 'stage3_rgb_phase1', tensor=True,
                    extra_env={'HB_UMMA_PREFETCH_B': '1'} if prefetch_b else None)
    g = np.load(os.path.join(HERE, 'golden', 'stage3_rgb_phase1.npz'))
    assert abs(out['loss'] - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert out['trans_err'] < 1e-5 and out['prior_mean_err'] < 1e-5
    for k, e in out['grad_err'].items():
        assert e < 1e-3, (k, e)          # hardware tolerance for this mode is 1e-2 (truncating accumulator, BPTT amplification)
