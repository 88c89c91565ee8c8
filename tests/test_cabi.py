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
    assert C.sizeof(_ext.HbHumorWeights) == 8 * (4 + 4 + 3 + 3 + 4 + 5 + 5 + 4 + 4 + 5 + 20 + 16) + 8
    assert C.sizeof(_ext.HbLbsModel) == 16 + 8 * 9 + 8 * 2 + 8 + 8 * 3 + 8 * 4 + 8 + 8 * 3 + 8 + 8 * 2 + 8 * 3


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
