"""Child process (HB_EMUL_TENSOR=1): the dense tensor-core LBS forward through the library's real dispatch and TMA-descriptor
code on the emulated tcgen05 kernels, for every kernel form of humor_lbs_configure; compared with the
fp64 oracle and with the exact-fp32 FFMA path.  Prints JSON."""
import ctypes as C
import json
import sys

import numpy as np
import torch

root, lib, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
import cpu_backend  # noqa: E402

L = cpu_backend.install(lib)
from humor_b200 import synth  # noqa: E402
from humor_b200.body_model import BodyModel  # noqa: E402
from oracle.smplh_lbs import OracleBodyModel  # noqa: E402

asset = synth.make_smplh_asset()
bm = BodyModel(asset, num_betas=16, batch_size=n, use_vtx_selector=True)
ob = OracleBodyModel(asset, use_vtx_selector=True)
rng = np.random.RandomState(n)
ro, pb, be, tr = (torch.tensor(a) for a in (rng.randn(n, 3).astype(np.float32) * 0.8, (rng.randn(n, 63) * 0.4).astype(np.float32),
                                            (rng.randn(n, 16) * 0.7).astype(np.float32), rng.randn(n, 3).astype(np.float32)))
o = ob(root_orient=ro, pose_body=pb, betas=be, trans=tr)
m = bm.lbs_model
out = {}


def used():
    a, b = C.c_int(0), C.c_int(0)
    L.humor_lbs_forms_used(C.byref(a), C.byref(b))
    return [a.value, b.value]


m.struct.use_umma = 0
exact = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
out['exact_vs_oracle'] = float((exact.v - o.v).abs().max())
m.struct.use_umma = 1
for skin, blend in [(int(f[0]), int(f[1])) for f in (sys.argv[4].split(';') if len(sys.argv) > 4 else ['11', '31', '35'])]:
    assert L.humor_lbs_configure(skin, blend, 512) == 0
    g = bm(root_orient=ro, pose_body=pb, betas=be, trans=tr)
    out[f'forms_{skin}{blend}'] = {'used': used(), 'v_vs_oracle': float((g.v - o.v).abs().max()), 'J_vs_oracle': float((g.Jtr - o.Jtr).abs().max()),
                                   'v_vs_exact': float((g.v - exact.v).abs().max()), 'finite': bool(torch.isfinite(g.v).all())}
print(json.dumps(out))
