"""ORACLE (test infrastructure) — build recipe for the parts of the reference that compile from their own sources.

The only compiled code of davrempe/humor on (or next to) the Stage-III path is the chamfer module
(humor/utils/chamfer_distance/chamfer_distance.cpp; its CPU entry points `forward` / `backward`, pybind names
chamfer_distance.cpp:180-185).  This script compiles that file WHERE IT LIES under /root/reference with g++ —
no source is copied into the repo, the reference's own JIT build (torch.utils.cpp_extension.load) is not run —
and writes ONLY oracle/_ref/cd_ref.so (git-ignored; it travels to the GPU box with the snapshot).

chamfer_distance.cpp also declares two CUDA launchers that live in chamfer_distance.cu; the CPU oracle never calls
them, so oracle/cd_ref_stub.cpp (ours) defines them as failing stubs to satisfy the linker.

    python oracle/build_ref.py            # build if missing / stale
    from oracle.build_ref import load_cd_ref; cd = load_cd_ref()   # None when neither built nor buildable
"""
import importlib.util
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/humor/utils/chamfer_distance/chamfer_distance.cpp'
OUT_DIR = os.path.join(HERE, '_ref')
OUT_SO = os.path.join(OUT_DIR, 'cd_ref.so')
STUB = os.path.join(HERE, 'cd_ref_stub.cpp')


def build(force=False, verbose=False):
    """Returns the path of the built module, or None when the reference sources are not present (GPU box)."""
    if not os.path.exists(REF_SRC):
        return OUT_SO if os.path.exists(OUT_SO) else None
    if not force and os.path.exists(OUT_SO) and os.path.getmtime(OUT_SO) >= max(os.path.getmtime(REF_SRC),
                                                                               os.path.getmtime(STUB)):
        return OUT_SO
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cmd = ['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-DTORCH_EXTENSION_NAME=cd_ref',
           '-DTORCH_API_INCLUDE_EXTENSION_H', f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
    cmd += [f'-I{p}' for p in ce.include_paths()] + [f'-I{sysconfig.get_paths()["include"]}']
    cmd += [REF_SRC, STUB, '-o', OUT_SO, f'-L{tlib}', '-ltorch', '-ltorch_cpu', '-lc10', '-ltorch_python',
            f'-Wl,-rpath,{tlib}']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('building the reference chamfer module failed:\n' + r.stdout[-4000:])
    if verbose:
        print('built', OUT_SO)
    return OUT_SO


def load_cd_ref():
    """The compiled reference module (functions forward / backward on CPU tensors) or None."""
    if not os.path.exists(OUT_SO):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location('cd_ref', OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose=True)
    print(p if p else 'reference sources not present and no prebuilt oracle/_ref/cd_ref.so')
