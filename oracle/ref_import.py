"""ORACLE (test infrastructure) — import the UNMODIFIED reference from /root/reference.

Only works in the build container (the GPU box has no /root/reference).  The reference's
``body_model/body_model.py:7-9`` imports ``smplx`` at module load; smplx is absent, so a stub
is installed in ``sys.modules`` whose ``SMPLH`` is backed by the restated LBS
(oracle/smplh_lbs.py).  Everything else (HumorModel, FittingLoss, MotionOptimizer,
transforms, fitting_utils) is the reference's own code, executed as is.
"""
import os
import sys
import types
import numpy as np
import torch

REF_ROOT = '/root/reference/humor'


def available():
    return os.path.isdir(REF_ROOT)


def _install_smplx_stub():
    if 'smplx' in sys.modules and getattr(sys.modules['smplx'], '_humor_b200_stub', False):
        return
    from oracle.smplh_lbs import SMPLHOracle, EXTRA_VERTEX_IDS

    class Struct:
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

    class _Out:
        pass

    class SMPL(torch.nn.Module):
        SHAPE_SPACE_DIM = 300
        NUM_JOINTS = 23
        NUM_BODY_JOINTS = 23

    class SMPLH(torch.nn.Module):
        NUM_BODY_JOINTS = 21
        NUM_HAND_JOINTS = 15
        NUM_JOINTS = NUM_BODY_JOINTS + 2 * NUM_HAND_JOINTS
        SHAPE_SPACE_DIM = 300

        def __init__(self, model_path, data_struct=None, num_betas=10, batch_size=1, **kw):
            super().__init__()
            assert kw.get('use_pca', False) is False and kw.get('flat_hand_mean', True) is True
            asset = {k: getattr(data_struct, k) for k in
                     ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'weights', 'kintree_table', 'f')}
            self._core = SMPLHOracle(asset, num_betas=num_betas)
            self.batch_size = batch_size
            self.register_buffer('faces_tensor', self._core.faces)
            self.register_buffer('_dummy', torch.zeros(1))

        def _apply(self, fn, *a, **k):
            super()._apply(fn, *a, **k)
            c = self._core
            for n in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights'):
                setattr(c, n, fn(getattr(c, n)))
            return self

        def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                    right_hand_pose=None, transl=None, return_full_pose=False, **kw):
            v, J, fp = self._core.forward(betas, global_orient, body_pose, transl,
                                          left_hand_pose, right_hand_pose)
            o = _Out()
            o.vertices, o.joints, o.betas, o.body_pose, o.full_pose = v, J, betas, body_pose, fp
            o.left_hand_pose, o.right_hand_pose = fp[:, 66:111], fp[:, 111:156]
            return o

    class SMPLX(SMPLH):
        NUM_JOINTS = 54

    m = types.ModuleType('smplx')
    m._humor_b200_stub = True
    m.SMPL, m.SMPLH, m.SMPLX = SMPL, SMPLH, SMPLX
    vid = types.ModuleType('smplx.vertex_ids')
    names = ['nose', 'reye', 'leye', 'rear', 'lear', 'LBigToe', 'LSmallToe', 'LHeel', 'RBigToe',
             'RSmallToe', 'RHeel', 'lthumb', 'lindex', 'lmiddle', 'lring', 'lpinky',
             'rthumb', 'rindex', 'rmiddle', 'rring', 'rpinky']
    vid.vertex_ids = {'smplh': dict(zip(names, EXTRA_VERTEX_IDS))}
    ut = types.ModuleType('smplx.utils')
    ut.Struct = Struct
    m.vertex_ids, m.utils = vid, ut
    sys.modules['smplx'] = m
    sys.modules['smplx.vertex_ids'] = vid
    sys.modules['smplx.utils'] = ut


_MODS = {}


def load():
    """Returns a namespace with the reference modules (imported once)."""
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')
    if _MODS:
        return types.SimpleNamespace(**_MODS)
    _install_smplx_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        _MODS['transforms'] = importlib.import_module('utils.transforms')
        _MODS['body_model'] = importlib.import_module('body_model.body_model')
        _MODS['bm_utils'] = importlib.import_module('body_model.utils')
        _MODS['humor_model'] = importlib.import_module('models.humor_model')
        _MODS['fitting_utils'] = importlib.import_module('fitting.fitting_utils')
        _MODS['fitting_loss'] = importlib.import_module('fitting.fitting_loss')
        _MODS['motion_optimizer'] = importlib.import_module('fitting.motion_optimizer')
        _MODS['logging'] = importlib.import_module('utils.logging')
    return types.SimpleNamespace(**_MODS)


def quiet_logger(ref):
    """The reference logs every loss term every closure (fitting_utils.py:261-272)."""
    ref.logging.Logger.log = staticmethod(lambda *a, **k: None)
    ref.motion_optimizer.log_cur_stats = lambda *a, **k: None
    ref.fitting_utils.log_cur_stats = lambda *a, **k: None
