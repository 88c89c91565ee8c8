"""Test infrastructure: run the product's Python surfaces on CPU tensors against the kernel-executing CUDA runtime stand-in
(tests/host/emul/cudart_emul.cpp).  Must be imported in a process started with LD_PRELOAD=<libcudart_emul.so>; `install`
points the C-ABI loader at the `-cudart shared` link of the product objects and removes the product's "CUDA tensors only"
guards FOR THIS TEST PROCESS ONLY.  The product has no CPU path of its own: without this harness every wrapper raises."""
import os

import numpy as np
import torch


def install(lib_path):
    from humor_b200 import _ext, body_model
    assert 'cudart_emul' in os.environ.get('LD_PRELOAD', ''), 'start the process with the emulating runtime preloaded'
    _ext.LIB_PATH = lib_path
    _ext._LIB = None
    _ext.require_cuda = lambda *t: None
    _ext.stream_ptr = lambda: None

    def cpu_lbs_model(self, packed, device):
        """LbsModel.__init__ for host memory (same layouts as the product's constructor, body_model.py:LbsModel)."""
        self.device = torch.device('cpu')
        self.t = {k: torch.as_tensor(v).contiguous() for k, v in packed.items() if isinstance(v, np.ndarray)}
        s = _ext.HbLbsModel()
        s.num_verts, s.v3_ld, s.wk, s.reserved = packed['num_verts'], packed['v3_ld'], packed['wk'], 0
        for k in ('v_template', 'blend', 'blend_t', 'j_template', 'j_dirs', 'w_idx', 'w_val', 'parents', 'extra_ids'):
            setattr(s, k, self.t[k].data_ptr())
        split = lambda x: (body_model._tf32_rn(x).contiguous(),)
        bt = torch.zeros(packed['v3_ld'], 224)
        bt[:, :208] = self.t['blend_t']
        hi = split(bt)[0]
        self.t['blend_t_hi'], self.t['blend_t_lo'] = hi, (bt - hi).contiguous()
        s.blend_t_hi, s.blend_t_lo = self.t['blend_t_hi'].data_ptr(), self.t['blend_t_lo'].data_ptr()
        s.use_umma = 1 if os.environ.get('HB_EMUL_TENSOR') else 0
        s.g_start, s.g_joint, s.g_w = (self.t[k].data_ptr() for k in ('g_start', 'g_joint', 'g_w'))
        s.num_groups, s.max_depth = packed['num_groups'], packed['max_depth']
        s.g_slot, s.ft_tab, s.ft_nct = self.t['g_slot'].data_ptr(), self.t['ft_tab'].data_ptr(), packed['ft_nct']
        bs = torch.zeros(packed['v3_ld'], 256)
        bs[:, :224] = bt * 1024.0
        bh = bs.to(torch.float16)
        self.t['blend16a_h'], self.t['blend16a_l'] = bh.contiguous(), (bs - bh.float()).to(torch.float16).contiguous()
        s.blend16a_h, s.blend16a_l = self.t['blend16a_h'].data_ptr(), self.t['blend16a_l'].data_ptr()
        s.depth, s.child_start, s.child_list = (self.t[k].data_ptr() for k in ('depth', 'child_start', 'child_list'))
        self.ws_slot = 0
        self.struct = s
        self._ws, self._vlists = {}, {}

    body_model.LbsModel.__init__ = cpu_lbs_model

    def lbs_model(self):
        if self._model is None:
            self._model = body_model.LbsModel(self._packed, 'cpu')
        return self._model

    body_model.BodyModel.lbs_model = property(lbs_model)
    body_model.BodyModel.set_precision = lambda self, mode: setattr(self.lbs_model.struct, 'use_umma', 0 if mode == 'exact' else 1)

    from humor_b200 import humor_model

    def packed(self):
        """HumorModel.packed for host memory: exact-fp32 kernels only."""
        dev = self.decoder.net[0].weight.device
        if self._packed is None or self._packed.device != dev:
            self._packed = humor_model.PackedWeights(self.decoder, self.prior_net, dev)
        self._packed.struct.use_umma = self._UMMA_MODE[self.precision] if os.environ.get('HB_EMUL_TENSOR') else 0
        return self._packed

    humor_model.HumorModel.packed = packed
    return _ext.lib()
