"""Child process of tests/test_host_dispatch.py: runs humor_lbs_fwd (dense, tensor-core path) on the REAL packed SMPL+H
constants through a library linked against a fake CUDA runtime and prints the kernels it launches (JSON on the last line)."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, sys.argv[1])
lib_path, skin, blend, N = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
fpb, norec = (int(sys.argv[6]) if len(sys.argv) > 6 else 60), (len(sys.argv) > 7 and sys.argv[7] == 'norec')
from humor_b200 import synth, _ext  # noqa: E402
from humor_b200 import body_model as BM  # noqa: E402

_ext.LIB_PATH = lib_path
packed = BM.pack_smplh(synth.make_smplh_asset(), 16)
t = {k: torch.as_tensor(v).contiguous() for k, v in packed.items() if isinstance(v, np.ndarray)}
s = _ext.HbLbsModel()
s.num_verts, s.v3_ld, s.wk = packed['num_verts'], packed['v3_ld'], packed['wk']
for k in ('v_template', 'blend', 'blend_t', 'j_template', 'j_dirs', 'w_idx', 'w_val', 'parents', 'extra_ids', 'depth',
          'child_start', 'child_list', 'g_start', 'g_joint', 'g_w', 'g_slot', 'ft_tab', 'ft_rec'):
    setattr(s, k, t[k].data_ptr())
s.ft_nct, s.ft_rec_stride = packed['ft_nct'], packed['ft_rec'].shape[1]
s.flags = 1 | (2 if packed['w_rows_sum_to_one'] else 0)      # as body_model.LbsModel: template in column 205 of the planes
planes = torch.zeros(packed['v3_ld'], 224)
s.blend_t_hi = s.blend_t_lo = s.blend16a_h = s.blend16a_l = s.blend16p_h = s.blend16p_l = planes.data_ptr()
s.use_umma, s.max_depth, s.num_groups = 1, packed['max_depth'], packed['num_groups']
if norec:              # a mesh whose tiles need more entries than a record buffer holds gets no records (body_model.fuseg_records)
    s.ft_rec, s.ft_rec_stride = None, 0
L = _ext.lib()
ws = torch.empty(L.humor_lbs_workspace_bytes(N) // 4)
x = torch.zeros(N * 64)
out = torch.empty(8)
rc_cfg = L.humor_lbs_configure(skin, blend, 0)
nl = C.c_int64(0)
sys.stdout.flush()
rc = L.humor_lbs_fwd(C.byref(s), N, fpb, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), ws.data_ptr(), ws.numel() * 4, None, 0,
                     out.data_ptr(), out.data_ptr(), 73, C.byref(nl), None)
a, b = C.c_int(0), C.c_int(0)
L.humor_lbs_forms_used(C.byref(a), C.byref(b))
sys.stdout.flush()
print(json.dumps({'rc_cfg': rc_cfg, 'rc': rc, 'launches': nl.value, 'used': [a.value, b.value], 'v3_ld': packed['v3_ld']}))
