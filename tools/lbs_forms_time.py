"""Times the dense LBS forward (B x T = 256 x 60 frames) for the kernel forms of humor_lbs_configure and VERIFIES each one
against the default forms on the same inputs: the forms a call actually launched (humor_lbs_forms_used), max |dv| against
form (1, 1), bitwise difference (a different summation order must show in the last bits - evidence the form really ran), and
run-to-run determinism.  One JSON line per form, flushed as soon as it is measured, so that a form that faults does not take
the earlier results with it; bench.py runs this in bounded child processes for its `roofline_candidates`.
  python tools/lbs_forms_time.py [--forms "3,1;3,5"] [--reps R] [--slab S] [--frames-per-seq T] [--seqs B]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humor_b200 import synth, _ext  # noqa: E402
from humor_b200.body_model import BodyModel, lbs  # noqa: E402

LBS_BYTES_FWD = 83896            # algorithmic bytes per frame of the dense forward, as in bench.py
ALL = [(1, 1), (3, 1), (3, 5)]
# vertex tolerance against form (1, 1): three-pass / three-product forms differ by summation order only
TOL = {1: 5e-6, 5: 5e-6}


def measure(forms, B, T, reps=5, slab=512, device='cuda', peak_gbs=0.0, emit=None):
    """Runs every form on the same seeded inputs; yields one record per form through `emit` as soon as it is measured.
    device='cpu' only works inside the emulation harness of tests/host/emul (the product itself has no CPU path)."""
    import time
    cuda = device == 'cuda'
    N = B * T
    bm = BodyModel(synth.make_smplh_asset(), num_betas=16, batch_size=N, use_vtx_selector=True)
    if cuda:
        bm = bm.to('cuda')
    m = bm.lbs_model
    g = torch.Generator(device='cpu').manual_seed(0)
    ro = (torch.randn(N, 3, generator=g) * 0.5).to(device)
    pb = (torch.randn(N, 63, generator=g) * 0.3).to(device)
    be = (torch.randn(B, 16, generator=g) * 0.5).to(device)
    tr = torch.randn(N, 3, generator=g).to(device)
    L = _ext.lib()
    sync = torch.cuda.synchronize if cuda else (lambda: None)

    def run():
        with torch.no_grad():
            v, _, J = lbs(m, ro, pb, be, tr, T, None, True, False, 73)
        return v, J

    _ext.check(L.humor_lbs_configure(1, 1, slab), 'configure')
    ref, refJ = run()
    sync()
    ref, refJ = ref.clone(), refJ.clone()
    out = []
    try:
        for skin, blend in forms:
            _ext.check(L.humor_lbs_configure(skin, blend, slab), 'configure')
            v, J = run()
            v2, _ = run()
            sync()
            us, ub = C.c_int(0), C.c_int(0)
            L.humor_lbs_forms_used(C.byref(us), C.byref(ub))
            if cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                sync()
                ms = e0.elapsed_time(e1) / reps
            else:
                t0 = time.perf_counter()
                for _ in range(reps):
                    run()
                ms = (time.perf_counter() - t0) * 1e3 / max(reps, 1)
            dv, dj = float((v - ref).abs().max()), float((J - refJ).abs().max())
            finite = bool(torch.isfinite(v).all())
            ran = (us.value, ub.value) == (skin, blend)
            same = bool(torch.equal(v, v2))
            rec = {'skin': skin, 'blend': blend, 'used': [us.value, ub.value], 'slab': slab, 'frames': N, 'ms': ms,
                   'GBps': N * LBS_BYTES_FWD / (ms * 1e-3) / 1e9 if ms > 0 else None, 'max_abs_diff_vs_11': dv,
                   'max_abs_diff_joints_vs_11': dj, 'bitwise_equal_to_11': bool(torch.equal(v, ref)), 'deterministic': same,
                   'finite': finite, 'verified': bool(ran and finite and dv < TOL[blend] and dj < TOL[blend] and same)}
            if peak_gbs > 0 and rec['GBps']:
                rec['frac'] = rec['GBps'] / peak_gbs
            out.append(rec)
            if emit:
                emit(rec)
            del v, v2, J
    finally:
        L.humor_lbs_configure(3, 5, slab)              # the library's defaults
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--forms', default='', help='"skin,blend;skin,blend;..." (default: all)')
    ap.add_argument('--only', default='', help='one form "skin,blend" (kept for older scripts)')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--slab', type=int, default=512)
    ap.add_argument('--seqs', type=int, default=256)
    ap.add_argument('--frames-per-seq', type=int, default=60)
    ap.add_argument('--peak-gbs', type=float, default=0.0)
    args = ap.parse_args()
    forms = ALL
    if args.forms or args.only:
        forms = [tuple(int(x) for x in f.split(',')) for f in (args.forms or args.only).split(';') if f]
    measure(forms, args.seqs, args.frames_per_seq, args.reps, args.slab, 'cuda', args.peak_gbs,
            emit=lambda rec: print(json.dumps(rec), flush=True))


if __name__ == '__main__':
    main()
