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
        """LbsModel.__init__ for host memory: the product's own table / plane construction (body_model.LbsModel._build)."""
        self._build(packed, 'cpu')
        self.struct.use_umma = 1 if os.environ.get('HB_EMUL_TENSOR') else 0

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
