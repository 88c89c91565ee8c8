"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/humor_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

from humor_b200 import _ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'humor_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(humor_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(built_lib):
    decl = declared_functions()
    assert len(decl) >= 12
    for name in decl:
        assert hasattr(built_lib, name), f'{name} declared in include/humor_b200.h but not exported'
    assert sorted(_ext.EXPORTS) == decl


def test_version_string(built_lib):
    assert b'sm_100a' in built_lib.humor_b200_version()


def test_workspace_queries(built_lib):
    assert built_lib.humor_lbs_workspace_bytes(64) > 64 * (208 + 624) * 4
    assert built_lib.humor_rollout_workspace_bytes(4, 7) > 4 * 7 * 7000 * 4


def test_struct_sizes_match_header():
    import ctypes as C
    # pointers-only structs: one slot per array entry
    assert C.sizeof(_ext.HbHumorWeights) == 8 * (4 + 4 + 3 + 3 + 4 + 5 + 5 + 4 + 4 + 5 + 20 + 16) + 8 + 8 * 8 + 8 * 10 + 8 * 2
    assert C.sizeof(_ext.HbLbsModel) == 16 + 8 * 9 + 8 * 2 + 8 + 8 * 3 + 8 * 3 + 8 + 8 * 2 + 8 * 2 + 8 * 2 + 8 + 8 + 8 * 2


def test_lbs_model_layout_matches_the_compiled_header(tmp_path):
    """sizeof / offsetof of HbLbsModel as gcc sees include/humor_b200.h against the ctypes mirror (every field)."""
    import ctypes as C
    import subprocess
    fields = [n for n, _ in _ext.HbLbsModel._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "humor_b200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(HbLbsModel));\n' +
                   ''.join(f'  printf("%zu\\n", offsetof(HbLbsModel, {n}));\n' for n in fields) + '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I' + os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert got[0] == C.sizeof(_ext.HbLbsModel)
    assert got[1:] == [getattr(_ext.HbLbsModel, n).offset for n in fields]


def test_product_has_no_oracle_import():
    """The product package must never import the oracle (no CPU fallback)."""
    pkg = os.path.join(ROOT, 'humor_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            txt = open(os.path.join(pkg, f)).read()
            assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_argument_errors_are_reported_before_any_launch(built_lib):
    """Error behaviour of the boundary (include/humor_b200.h): bad arguments and short workspaces come back as HB_ERR_* codes
    from the argument checks, i.e. before any CUDA call - which is why this runs without a GPU.  (The reference's only native
    boundary, the chamfer module, merely printf's launch errors: chamfer_distance.cu:155-157.)"""
    import ctypes as C
    L = built_lib
    ARG, WS = 1001, 1002
    hdr = open(os.path.join(ROOT, 'include', 'humor_b200.h')).read()
    assert '#define HB_ERR_ARG 1001' in hdr and '#define HB_ERR_WORKSPACE 1002' in hdr
    buf = (C.c_float * 4096)()
    p = C.cast(buf, C.c_void_p)
    m = _ext.HbLbsModel()
    nl = C.c_int64(0)
    # LBS: NULL model / inputs, bad joint count, N <= 0, workspace too small
    assert L.humor_lbs_fwd(None, 4, 1, p, p, p, p, p, 1 << 30, None, 0, p, p, 73, C.byref(nl), None) == ARG
    assert L.humor_lbs_fwd(C.byref(m), 0, 1, p, p, p, p, p, 1 << 30, None, 0, p, p, 73, C.byref(nl), None) == ARG
    assert L.humor_lbs_fwd(C.byref(m), 4, 1, None, p, p, p, p, 1 << 30, None, 0, p, p, 73, C.byref(nl), None) == ARG
    assert L.humor_lbs_fwd(C.byref(m), 4, 1, p, p, p, p, p, 1 << 30, None, 0, p, p, 60, C.byref(nl), None) == ARG
    assert L.humor_lbs_fwd(C.byref(m), 4, 0, p, p, p, p, p, 1 << 30, None, 0, p, p, 73, C.byref(nl), None) == ARG
    assert L.humor_lbs_fwd(C.byref(m), 4, 1, p, p, p, p, p, 16, None, 0, p, p, 73, C.byref(nl), None) == WS
    assert L.humor_lbs_bwd(None, 4, 1, p, p, p, p, p, 1 << 30, None, 0, p, p, 73, p, p, p, p, C.byref(nl), None) == ARG
    assert L.humor_lbs_bwd(C.byref(m), 4, 1, p, p, p, p, p, 16, None, 0, p, p, 73, p, p, p, p, C.byref(nl), None) == WS
    # rollout: NULL weights / state, short workspace
    w = _ext.HbHumorWeights()
    assert L.humor_rollout_fwd(None, 2, 3, p, p, p, 1 << 30, p, p, C.byref(nl), None) == ARG
    assert L.humor_rollout_fwd(C.byref(w), 2, 3, None, p, p, 1 << 30, p, p, C.byref(nl), None) == ARG
    assert L.humor_rollout_fwd(C.byref(w), 0, 3, p, p, p, 1 << 30, p, p, C.byref(nl), None) == ARG
    assert L.humor_rollout_fwd(C.byref(w), 2, 3, p, p, p, 16, p, p, C.byref(nl), None) == WS
    assert L.humor_rollout_bwd(C.byref(w), 2, 3, p, 16, p, p, p, p, C.byref(nl), None) == WS
    assert L.humor_rollout_bwd(C.byref(w), 2, 3, None, 1 << 30, p, p, p, p, C.byref(nl), None) == ARG
    # energies, GMM, rotations, chamfer, kernel-form switches
    assert L.humor_fit_losses(None, C.byref(nl), None) == ARG
    a = _ext.HbFitArgs()
    a.B, a.T, a.njx = 2, 3, 60
    assert L.humor_fit_losses(C.byref(a), C.byref(nl), None) == ARG
    assert L.humor_gmm_nll(0, 4, 2, p, p, p, p, p, p, p, None) == ARG
    assert L.humor_gmm_nll(2, 4, 2, None, p, p, p, p, p, p, None) == ARG
    assert L.humor_rodrigues_fwd(0, p, p, None) == ARG and L.humor_rodrigues_fwd(4, None, p, None) == ARG
    assert L.humor_mat2aa_bwd(4, p, None, p, None) == ARG
    assert L.humor_chamfer_fwd(-1, 4, p, 4, p, p, p, p, p, C.byref(nl), None) == ARG
    assert L.humor_chamfer_fwd(1, 4, None, 4, p, p, p, p, p, C.byref(nl), None) == ARG
    assert L.humor_chamfer_fwd(1, 4, p, 4, p, p, None, p, p, C.byref(nl), None) == ARG        # dist1 without idx1
    assert L.humor_umma_gemm16(None, 64, p, 64, p, p, 64, 4, 4, 64, p, 1 << 30, None) == ARG
    assert L.humor_umma_gemm16(p, 64, p, 64, p, p, 64, 4, 4, 64, p, 16, None) == WS
    assert L.humor_lbs_configure(4, 0, 0) == ARG and L.humor_lbs_configure(0, 6, 0) == ARG and L.humor_lbs_configure(0, 0, 64) == ARG
    assert L.humor_lbs_configure(2, 0, 0) == ARG and L.humor_lbs_configure(0, 3, 0) == ARG          # forms removed in round 2
    assert L.humor_lbs_configure(0, 0, 0) == 0
    assert nl.value == 0


def test_integration_doc_stub_mirrors_the_struct():
    """The ctypes stub INTEGRATION.md shows a maintainer must list every field of HbLbsModel (the library reads all of them)."""
    import ctypes as C
    src = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = src[src.index('P, I = C.c_void_p, C.c_int'):src.index('# the first 13 fields')]
    ns = {'C': C}
    exec(code, ns)
    assert [n for n, _ in ns['HbLbsModel']._fields_] == [n for n, _ in _ext.HbLbsModel._fields_]
    assert C.sizeof(ns['HbLbsModel']) == C.sizeof(_ext.HbLbsModel)
