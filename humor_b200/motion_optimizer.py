"""MotionOptimizer — drop-in for humor/fitting/motion_optimizer.py:29-1120.

Same constructor and ``run`` contract as the reference (so ``run_fitting.py:385-416`` is unchanged);
the Stage-III closure (motion_optimizer.py:514-608) is re-planned around the native kernels:

  reference closure                                  here
  -------------------------------------------------  ---------------------------------------------
  5 full SMPL evaluations on B*T rows                1 dense + 1 joints-only on B*T rows, 3 joints-only
   (3 of them a T=1 state expanded to T)              on B rows (results identical: rows are independent)
  59 sequential python steps x ~200 eager kernels    decoder: 5 launches/step; prior: one batched MLP
  python loop over B for the overlap energies        one fused energy+gradient kernel
  autograd through (N,6890,4,4) transforms           reverse LBS over the 43+21 vertices that carry gradient

L-BFGS itself is ``torch.optim.LBFGS`` with strong-Wolfe line search exactly as in the reference
(:24,:461-478) — moving it onto the device is the scope table's next row.
"""
import os

import numpy as np
import torch

from . import _ext
from .body_model import KEYPT_VERTS, lbs
from .fitting_loss import FittingLoss, SMPL2OP
from .fitting_utils import OP_EDGE_LIST, OP_IGNORE_JOINTS, compute_cam2prior, parse_floor_plane
from .transforms import batch_rodrigues, rotation_matrix_to_angle_axis

# CTAs per SM of the dense vertex pass queued behind the reverse decoder chain: short chunks of the tile list that the
# hardware scheduler slots into the SMs the chain leaves idle.  8 was measured (profiles/r02x_*); HB_DENSE_CTAS_PER_SM for A/B runs.
_DENSE_CTAS_PER_SM = int(os.environ.get('HB_DENSE_CTAS_PER_SM', '8'))

LINE_SEARCH = 'strong_wolfe'
J_BODY = 21
CONTACT_THRESH = 0.5
CONTACT_INDS = [0, 4, 5, 7, 8, 10, 11, 20, 21]
NUM_JOINTS = 22


class _RolloutOutputs(torch.autograd.Function):
    """humor_rollout_outputs_fwd / _bwd (csrc/rot.cu): world rows of the CVAE chain + the frame-0 state -> the (B,T,.) tensors the
    energies read, in the prior frame and (R, t given) in the camera frame.  Returns (trans, root_orient, pose_body, joints (B,T,22,3),
    contacts_logits (B,S,9), contacts_conf, contacts, cam_trans | None, cam_root_orient | None)."""

    @staticmethod
    def forward(ctx, world, trans0, orient0, pose0, joints0, R, t, contact_idx):
        from . import _ext
        _ext.require_cuda(world, trans0, orient0, pose0, joints0)
        world, trans0, orient0, pose0, joints0 = (_ext.f32c(x) for x in (world, trans0, orient0, pose0, joints0))
        S, B = world.shape[0], world.shape[1]
        T = S + 1
        if R is not None:
            R, t = _ext.f32c(R), _ext.f32c(t)
        new = lambda *sh: torch.empty(*sh, device=world.device, dtype=torch.float32)
        trans, orient, pose, joints = new(B, T, 3), new(B, T, 3), new(B, T, 63), new(B, T, NUM_JOINTS, 3)
        logits, conf, labels = new(B, S, 9), new(B, T, NUM_JOINTS), new(B, T, NUM_JOINTS)
        cam_t = new(B, T, 3) if R is not None else None
        cam_o = new(B, T, 3) if R is not None else None
        _ext.check(_ext.lib().humor_rollout_outputs_fwd(
            B, S, _ext.ptr(world), _ext.ptr(trans0), _ext.ptr(orient0), _ext.ptr(pose0), _ext.ptr(joints0), _ext.ptr(R), _ext.ptr(t),
            _ext.ptr(contact_idx), CONTACT_THRESH, _ext.ptr(trans), _ext.ptr(orient), _ext.ptr(pose), _ext.ptr(joints), _ext.ptr(logits),
            _ext.ptr(conf), _ext.ptr(labels), _ext.ptr(cam_t), _ext.ptr(cam_o), _ext.stream_ptr()), 'humor_rollout_outputs_fwd')
        _ext.LaunchCounter.total += 1
        ctx.save_for_backward(world, trans0, orient0, R)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(conf, labels)
        if R is None:
            return trans, orient, pose, joints, logits, conf, labels
        return trans, orient, pose, joints, logits, conf, labels, cam_t, cam_o

    @staticmethod
    def backward(ctx, g_trans, g_orient, g_pose, g_joints, g_logits, g_conf, g_labels, g_cam_t=None, g_cam_o=None):
        from . import _ext
        world, trans0, orient0, R = ctx.saved_tensors
        S, B = world.shape[0], world.shape[1]
        c = lambda g: None if g is None else _ext.f32c(g)
        g_trans, g_orient, g_pose, g_joints, g_logits, g_cam_t, g_cam_o = (c(g) for g in (g_trans, g_orient, g_pose, g_joints, g_logits,
                                                                                       g_cam_t, g_cam_o))
        new = lambda *sh: torch.empty(*sh, device=world.device, dtype=torch.float32)
        d_world, d_t0, d_o0, d_p0, d_j0 = new(S, B, 348), new(B, 3), new(B, 3), new(B, 63), new(B, 66)
        d_R = new(B, 3, 3) if R is not None else None
        d_t = new(B, 3) if R is not None else None
        _ext.check(_ext.lib().humor_rollout_outputs_bwd(
            B, S, _ext.ptr(world), _ext.ptr(trans0), _ext.ptr(orient0), _ext.ptr(R), _ext.ptr(g_trans), _ext.ptr(g_orient), _ext.ptr(g_pose),
            _ext.ptr(g_joints), _ext.ptr(g_logits), _ext.ptr(g_cam_t), _ext.ptr(g_cam_o), _ext.ptr(d_world), _ext.ptr(d_t0), _ext.ptr(d_o0),
            _ext.ptr(d_p0), _ext.ptr(d_j0), _ext.ptr(d_R), _ext.ptr(d_t), _ext.stream_ptr()), 'humor_rollout_outputs_bwd')
        _ext.LaunchCounter.total += 1
        return d_world, d_t0, d_o0, d_p0, d_j0, d_R, d_t, None


class MotionOptimizer():
    """Fits SMPL shape and motion to an observation sequence (3 stages of L-BFGS)."""

    def __init__(self, device, body_model, num_betas, batch_size, seq_len, observed_modalities, loss_weights,
                 pose_prior, motion_prior=None, init_motion_prior=None, optim_floor=False, camera_matrix=None,
                 robust_loss_type='none', robust_tuning_const=4.6851, joint2d_sigma=100,
                 stage3_tune_init_state=True, stage3_tune_init_num_frames=15, stage3_tune_init_freeze_start=30,
                 stage3_tune_init_freeze_end=50, stage3_contact_refine_only=False, use_chamfer=False,
                 im_dim=(1080, 1080)):
        B, T = batch_size, seq_len
        self.device = device
        self.batch_size, self.seq_len = B, T
        self.body_model = body_model
        self.num_betas = num_betas
        self.optim_floor = optim_floor
        self.stage3_tune_init_state = stage3_tune_init_state
        self.stage3_tune_init_num_frames = stage3_tune_init_num_frames
        self.stage3_tune_init_freeze_start = stage3_tune_init_freeze_start
        self.stage3_tune_init_freeze_end = stage3_tune_init_freeze_end
        self.stage3_contact_refine_only = stage3_contact_refine_only
        self.im_dim = im_dim
        if motion_prior is None:
            raise ValueError('Need the motion prior to use all-implicit parameterization!')
        self.pose_prior = pose_prior
        self.latent_pose_dim = pose_prior.latentD
        self.latent_pose = torch.zeros((B, T, self.latent_pose_dim), device=device)
        self.trans = torch.zeros((B, T, 3), device=device)
        self.root_orient = torch.zeros((B, T, 3), device=device)
        self.root_orient[:, :, 0] = np.pi
        self.betas = torch.zeros((B, num_betas), device=device)
        self.motion_prior = motion_prior
        self.init_motion_prior = init_motion_prior
        self.latent_motion = None
        self.latent_motion_dim = motion_prior.latent_size
        self.cond_prior = motion_prior.use_conditional_prior
        self.trans_vel = self.root_orient_vel = self.joints_vel = None
        self.init_fidx = np.zeros((B))
        self.cam_f = self.cam_center = None
        if optim_floor:
            if camera_matrix is None:
                raise ValueError('Must have camera intrinsics (camera_matrix) to optimize the floor plane!')
            self.floor_plane = torch.zeros((B, 3), device=device)
            self.floor_plane[:, 2] = 1.0
            self.cam2prior_R = torch.eye(3, device=device).reshape(1, 3, 3).expand(B, 3, 3)
            self.cam2prior_t = torch.zeros((B, 3), device=device)
            self.cam2prior_root_height = torch.zeros((B, 1), device=device)
            cm = camera_matrix.to(device)
            self.cam_f = torch.stack([cm[:, 0, 0], cm[:, 1, 1]], 1)
            self.cam_center = torch.stack([cm[:, 0, 2], cm[:, 1, 2]], 1)
        self.use_camera = self.cam_f is not None
        self.njo = 73 if getattr(body_model, 'use_vtx_selector', False) else 52
        self.smpl2op_map = np.asarray(SMPL2OP, dtype=np.int32)
        self.fitting_loss = FittingLoss(loss_weights, init_motion_prior, self.smpl2op_map, OP_IGNORE_JOINTS, self.cam_f,
                                        self.cam_center, robust_loss_type, robust_tuning_const,
                                        joints2d_sigma=joint2d_sigma, use_chamfer=use_chamfer).to(device)
        self.return_points3d = True      # dense vertices in every camera-frame SMPL evaluation (reference behaviour)
        # Nothing in the Stage-III energy reads the dense vertices (they are an OUTPUT of the closure), so their LBS pass
        # runs on a side stream under the latency-bound reverse decoder chain and is joined when the closure ends.
        self.overlap_dense = True
        self._dense_stream = None
        self._capture_stream = None
        self._dense_pending = False
        self._defer_dense_join = False
        # the gradient-free dense LBS pass of a closure evaluation is queued BEHIND the launch of the reverse decoder chain (which
        # occupies 128 of 148 SMs for 3 ms without loading them) in short CTAs that fill the idle SMs; HB_DENSE_EARLY=1: right after
        # the roll-out as in round 1
        import os
        self.dense_under_reverse_chain = os.environ.get('HB_DENSE_EARLY') is None
        self._dense_deferred = None
        # multi-GPU: this process owns a contiguous block of the sub-sequences; the overlap energies that couple
        # the last sequence of rank r with the first of rank r+1 are exchanged as small halos (parallel.py)
        self.shard = None
        # CUDA-graph capture of the Stage-III closure (forward + backward): one replay per L-BFGS evaluation
        import os
        self.lbfgs_impl = os.environ.get('HB_LBFGS', 'torch')    # 'torch' | 'native' (humor_b200.lbfgs); sharded runs use native
        self.use_cuda_graph = True       # falls back to eager launches (with a warning) if the closure cannot be captured
        self._graphs = {}
        self._contact_idx = torch.tensor(CONTACT_INDS, dtype=torch.long, device=device)
        self._contact_idx32 = self._contact_idx.to(torch.int32)

    def set_precision(self, mode):
        """'tensor' (default) or 'exact' for the GEMM-shaped kernels of the motion prior and the body model
        (see HumorModel.set_precision for the measured accuracy of each)."""
        self.motion_prior.set_precision(mode)
        self.body_model.set_precision(mode)
        self.precision = mode

    # ------------------------------------------------------------------------------------------------
    # SMPL
    # ------------------------------------------------------------------------------------------------
    def points3d_active(self, observed_data):
        """True when the point-cloud energy is live (fitting_loss.py:114): the dense vertices then carry gradient."""
        return 'points3d' in observed_data and self.fitting_loss.loss_weights.get('points3d', 0.0) > 0.0

    def smpl_results(self, trans, root_orient, body_pose, beta, dense=True, sel=True, njo=None, dense_grad=False):
        """motion_optimizer.py:1065-1110.  The reference expands a T=1 state to T rows (and pads short
        ones) because smplx needs a fixed batch; rows are independent, so only the rows given are run.
        ``dense_grad``: the dense vertices feed an energy (points3d), so they stay on the main stream inside autograd."""
        B, T, _ = trans.size()
        njo = self.njo if njo is None else njo
        dense = dense or dense_grad
        side_dense = dense and not dense_grad and self.overlap_dense and sel and trans.is_cuda
        v, vs, J = lbs(self.body_model.lbs_model, root_orient.reshape(B * T, 3), body_pose.reshape(B * T, 63), beta,
                       trans.reshape(B * T, 3), frames_per_beta=T, sel_ids=KEYPT_VERTS if sel else None,
                       want_dense=dense and not side_dense, dense_grad=dense_grad, num_joints_out=njo)
        if side_dense:
            v = self._dense_on_side_stream(trans, root_orient, body_pose, beta)
        J = J.reshape(B, T, njo, 3)
        pred = {'Jtr': J, 'joints3d': J[:, :, :NUM_JOINTS], 'joints3d_extra': J[:, :, NUM_JOINTS:],
                'faces': self.body_model.bm.faces_tensor}
        if sel:
            pred['verts3d'] = vs.reshape(B, T, len(KEYPT_VERTS), 3)
        if dense:
            pred['points3d'] = v.reshape(B, T, -1, 3)
        return pred, None

    def _dense_on_side_stream(self, trans, root_orient, body_pose, beta):
        """Dense vertices (no gradient) on a second stream with its own scratch; joined by ``join_dense``."""
        B, T, _ = trans.size()
        model = self.body_model.lbs_model
        main = torch.cuda.current_stream()
        if self._dense_stream is None:
            self._dense_stream = torch.cuda.Stream(device=trans.device)
        side = self._dense_stream
        if self._defer_dense_join and self.dense_under_reverse_chain and beta.requires_grad and torch.is_grad_enabled():
            # a reverse pass follows: allocate the output now, queue the kernels when the reverse roll-out has been launched
            from . import humor_model
            side.wait_stream(main)                   # (inside a capture: the side stream joins it, so the block below is graph-private)
            with torch.cuda.stream(side):            # the side stream's allocator pool: its blocks are only ever written there
                v = torch.empty(B * T, model.struct.num_verts, 3, device=trans.device, dtype=torch.float32)
            self._dense_deferred = (v, root_orient.detach().reshape(B * T, 3), body_pose.detach().reshape(B * T, 63), beta.detach(),
                                    trans.detach().reshape(B * T, 3), T)
            humor_model.AFTER_ROLLOUT_BWD.append(self._launch_deferred_dense)
            self._dense_pending = True
            return v
        side.wait_stream(main)
        with torch.cuda.stream(side), torch.no_grad():
            model.ws_slot = 1
            try:
                v, _, _ = lbs(model, root_orient.detach().reshape(B * T, 3), body_pose.detach().reshape(B * T, 63), beta.detach(),
                              trans.detach().reshape(B * T, 3), frames_per_beta=T, sel_ids=None, want_dense=True,
                              dense_grad=False, num_joints_out=52)
            finally:
                model.ws_slot = 0
        if not torch.cuda.is_current_stream_capturing():
            v.record_stream(main)
        self._dense_pending = True
        if not self._defer_dense_join:
            self.join_dense()
        return v

    def _launch_deferred_dense(self, after_reverse_chain_launch=True):
        """Queue the deferred dense pass on the side stream: behind the point where the reverse decoder chain was launched (called from
        humor_model.AFTER_ROLLOUT_BWD), or - no reverse pass came - behind everything queued so far."""
        if self._dense_deferred is None:
            return
        from . import _ext, humor_model
        from .body_model import lbs_dense_into
        v, ro, pb, be, tr, T = self._dense_deferred
        self._dense_deferred = None
        if self._launch_deferred_dense in humor_model.AFTER_ROLLOUT_BWD:
            humor_model.AFTER_ROLLOUT_BWD.remove(self._launch_deferred_dense)
        model = self.body_model.lbs_model
        side = self._dense_stream
        L = _ext.lib()
        if after_reverse_chain_launch:
            _ext.check(L.humor_rollout_bwd_started_wait(_ext.C.c_void_p(side.cuda_stream)), 'humor_rollout_bwd_started_wait')
        else:
            side.wait_stream(torch.cuda.current_stream())
        sms = torch.cuda.get_device_properties(v.device).multi_processor_count if v.is_cuda else 148
        with torch.cuda.stream(side), torch.no_grad():
            model.ws_slot = 1
            L.humor_lbs_set_fuseg_ctas(_DENSE_CTAS_PER_SM * sms if after_reverse_chain_launch else 0)
            try:
                lbs_dense_into(model, ro, pb, be, tr, T, v)
            finally:
                L.humor_lbs_set_fuseg_ctas(0)
                model.ws_slot = 0
        if not torch.cuda.is_current_stream_capturing():
            v.record_stream(torch.cuda.current_stream())      # read on the caller's stream after join_dense

    def join_dense(self):
        if self._dense_deferred is not None:           # no reverse roll-out was launched after the forward: queue the pass now
            self._launch_deferred_dense(after_reverse_chain_launch=False)
        if self._dense_pending:
            torch.cuda.current_stream().wait_stream(self._dense_stream)
            self._dense_pending = False

    def joints_only(self, trans, root_orient, body_pose, beta):
        B, T, _ = trans.size()
        _, _, J = lbs(self.body_model.lbs_model, root_orient.reshape(B * T, 3), body_pose.reshape(B * T, 63), beta,
                      trans.reshape(B * T, 3), frames_per_beta=T, sel_ids=None, want_dense=False, dense_grad=False,
                      num_joints_out=52)
        return J.reshape(B, T, 52, 3)[:, :, :NUM_JOINTS]

    def latent2pose(self, latent_pose):
        """VPoser decode (third-party object) -> axis-angle (motion_optimizer.py:1041-1051)."""
        B, T, _ = latent_pose.size()
        R = self.pose_prior.decode(latent_pose.reshape(-1, self.latent_pose_dim), output_type='matrot')
        return rotation_matrix_to_angle_axis(R.reshape(B * T * J_BODY, 3, 3)).reshape(B, T, J_BODY * 3)

    def pose2latent(self, body_pose):
        B, T, _ = body_pose.size()
        return self.pose_prior.encode(body_pose.reshape(-1, J_BODY * 3)).mean.reshape(B, T, self.latent_pose_dim)

    # ------------------------------------------------------------------------------------------------
    # camera <-> prior frame
    # ------------------------------------------------------------------------------------------------
    def apply_cam2prior(self, data_dict, R, t, root_height, body_pose, betas, key_frame_idx, inverse=False):
        """motion_optimizer.py:678-741 (key frame = frame 0 of each sequence, as everywhere in the reference)."""
        trans, root_orient = data_dict['trans'], data_dict['root_orient']
        B, T, _ = root_orient.size()
        Rt = R[:, None].expand(B, T, 3, 3)
        Rm = batch_rodrigues(root_orient.reshape(-1, 3)).reshape(B, T, 3, 3)
        Rm = torch.matmul(Rt.transpose(3, 2) if inverse else Rt, Rm)
        out = {'root_orient': rotation_matrix_to_angle_axis(Rm.reshape(B * T, 3, 3)).reshape(B, T, 3)}
        if inverse:
            tr = trans - trans[:, 0:1]
            tr = torch.matmul(Rt.transpose(3, 2), tr[..., None])[..., 0] - t[:, None]
        else:
            tr = torch.matmul(Rt, (trans + t[:, None])[..., None])[..., 0]
            cur_h = self.joints_only(tr, out['root_orient'], body_pose, betas)[:, 0, 0, 2:3]
            off = torch.cat([torch.zeros(B, 2, device=tr.device), root_height - cur_h], 1)
            tr = tr + off[:, None]
        out['trans'] = tr
        return out

    # ------------------------------------------------------------------------------------------------
    # rollout
    # ------------------------------------------------------------------------------------------------
    def rollout_latent_motion(self, trans, root_orient, body_pose, betas, prior_opt_params, latent_motion,
                              return_prior=False, return_vel=False, fit_gender='neutral', use_mean=False,
                              num_steps=-1, canonicalize_input=False):
        """motion_optimizer.py:876-1019 with the rollout, rotation conversions and SMPL joints on the kernels."""
        if latent_motion is None or canonicalize_input:
            raise NotImplementedError('sampling rollouts are outside the Stage-III fitting path')
        B, S = latent_motion.shape[0], latent_motion.shape[1]
        if self.optim_floor:
            pd = self.apply_cam2prior({'trans': trans, 'root_orient': root_orient}, self.cam2prior_R, self.cam2prior_t,
                                      self.cam2prior_root_height, body_pose, betas, self.init_fidx)
            trans, root_orient = pd['trans'], pd['root_orient']
        trans_vel, joints_vel, root_orient_vel = prior_opt_params
        joints = self.joints_only(trans, root_orient, body_pose, betas)                       # (B,1,22,3)
        R_all = batch_rodrigues(torch.cat([root_orient, body_pose], 2).reshape(-1, 3)).reshape(B, 22 * 9)
        init_state = torch.cat([trans[:, 0], trans_vel[:, 0], R_all[:, :9], root_orient_vel[:, 0], R_all[:, 9:],
                                joints.reshape(B, 66), joints_vel.reshape(B, 66)], 1)
        world, prior_out = self.motion_prior.roll_out_raw(init_state, latent_motion, return_prior)
        if not return_vel:
            # one kernel: matrix -> axis-angle of the 22 rotations, the frame-0 state in front, contact confidences / labels and the
            # camera-frame root orientation / translation (apply_cam2prior inverse) - csrc/rot.cu, humor_rollout_outputs_*
            res = _RolloutOutputs.apply(world, trans[:, 0], root_orient[:, 0], body_pose[:, 0], joints[:, 0].reshape(B, 66),
                                        self.cam2prior_R if self.optim_floor else None,
                                        self.cam2prior_t if self.optim_floor else None, self._contact_idx32)
            out = {'trans': res[0], 'root_orient': res[1], 'pose_body': res[2], 'joints': res[3], 'contacts_logits': res[4],
                   'contacts_conf': res[5], 'contacts': res[6]}
            if return_prior:
                out['prior_out'] = prior_out                                                  # (S,B,96) mean|logvar
                out['cond_prior'] = (prior_out[..., :48].permute(1, 0, 2), torch.exp(prior_out[..., 48:]).permute(1, 0, 2))
            cam = {'pose_body': out['pose_body']}
            cam['trans'], cam['root_orient'] = (res[7], res[8]) if self.optim_floor else (out['trans'], out['root_orient'])
            return out, cam
        return self._rollout_outputs_torch(world, prior_out, trans, root_orient, body_pose, betas, joints, trans_vel, joints_vel,
                                           root_orient_vel, return_prior, return_vel)

    def _rollout_outputs_torch(self, world, prior_out, trans, root_orient, body_pose, betas, joints, trans_vel, joints_vel,
                               root_orient_vel, return_prior, return_vel):
        """The same in torch ops (round-1 form): the path with velocities (Stage-III initialisation) and the cross-check of the kernel."""
        B, S = world.shape[1], world.shape[0]
        w = world.permute(1, 0, 2)                                                            # (B,S,348)
        rots = torch.cat([w[..., 6:15], w[..., 18:207]], -1).reshape(B * S * 22, 3, 3)
        aa = rotation_matrix_to_angle_axis(rots).reshape(B, S, 66)
        out = {
            'trans': torch.cat([trans, w[..., 0:3]], 1),
            'root_orient': torch.cat([root_orient, aa[..., :3]], 1),
            'pose_body': torch.cat([body_pose, aa[..., 3:]], 1),
            'joints': torch.cat([joints, w[..., 207:273].reshape(B, S, NUM_JOINTS, 3)], 1),
            'contacts_logits': w[..., 339:348],
        }
        if return_vel:
            out['trans_vel'] = torch.cat([trans_vel, w[..., 3:6]], 1)
            out['root_orient_vel'] = torch.cat([root_orient_vel, w[..., 15:18]], 1)
            out['joints_vel'] = torch.cat([joints_vel.reshape(B, 1, NUM_JOINTS, 3),
                                           w[..., 273:339].reshape(B, S, NUM_JOINTS, 3)], 1)
        if return_prior:
            out['prior_out'] = prior_out                                                      # (S,B,96) mean|logvar
            out['cond_prior'] = (prior_out[..., :48].permute(1, 0, 2), torch.exp(prior_out[..., 48:]).permute(1, 0, 2))
        with torch.no_grad():
            conf9 = torch.sigmoid(out['contacts_logits'])
            conf = torch.zeros(B, S, NUM_JOINTS, device=conf9.device)
            conf.index_copy_(2, self._contact_idx, conf9)          # device index: no host copy (graph-capturable)
            conf = torch.cat([conf[:, 0:1], conf], 1)
            out['contacts_conf'] = conf
            out['contacts'] = (conf > CONTACT_THRESH).to(torch.float)
        cam = {'pose_body': out['pose_body']}
        if self.optim_floor:
            cd = self.apply_cam2prior({'trans': out['trans'], 'root_orient': out['root_orient']}, self.cam2prior_R,
                                      self.cam2prior_t, self.cam2prior_root_height, out['pose_body'], betas,
                                      self.init_fidx, inverse=True)
            cam['trans'], cam['root_orient'] = cd['trans'], cd['root_orient']
        else:
            cam['trans'], cam['root_orient'] = out['trans'], out['root_orient']
        return out, cam

    # ------------------------------------------------------------------------------------------------
    # Stage-III closure (the hot path)
    # ------------------------------------------------------------------------------------------------
    def stage3_forward(self, observed_data, nsteps=None, init_motion_scale=1.0, fit_gender='neutral'):
        """Body of the closure of motion_optimizer.py:514-608 up to (loss, stats).
        nsteps=None: full-T phase (iterations >= stage3_tune_init_freeze_start)."""
        prior_opt_params = [self.trans_vel, self.joints_vel, self.root_orient_vel]
        T = self.seq_len
        cur_body_pose = self.latent2pose(self.latent_pose)
        if self.optim_floor:
            j0 = self.joints_only(self.trans, self.root_orient, cur_body_pose, self.betas)
            self.cam2prior_R, self.cam2prior_t, self.cam2prior_root_height = compute_cam2prior(
                self.floor_plane, self.trans[:, 0], self.root_orient[:, 0], j0[:, 0])
        z = self.latent_motion if nsteps is None else self.latent_motion[:, :nsteps - 1]
        roll, cam = self.rollout_latent_motion(self.trans, self.root_orient, cur_body_pose, self.betas, prior_opt_params,
                                               z, return_prior=self.cond_prior, fit_gender=fit_gender)
        need_latent_pose = self.fitting_loss.loss_weights['pose_prior'] > 0.0
        cur_latent_pose = self.pose2latent(roll['pose_body']) if need_latent_pose else None
        if self.optim_floor:
            prior_joints = self.joints_only(roll['trans'], roll['root_orient'], roll['pose_body'], self.betas)
            cam_pred, _ = self.smpl_results(cam['trans'], cam['root_orient'], roll['pose_body'], self.betas,
                                            dense=self.return_points3d, dense_grad=self.points3d_active(observed_data))
            cam_pred['floor_plane'] = self.floor_plane
        else:
            cam_pred, _ = self.smpl_results(roll['trans'], roll['root_orient'], roll['pose_body'], self.betas,
                                            dense=self.return_points3d, dense_grad=self.points3d_active(observed_data))
            prior_joints = cam_pred['joints3d']
        cam_pred['betas'] = self.betas
        cam_pred['latent_pose'] = cur_latent_pose
        pred = {'prior_joints3d': prior_joints, 'joints3d_rollout': roll['joints'], 'contacts_logits': roll['contacts_logits'],
                'latent_motion': z, 'prior_out': roll.get('prior_out'), 'joints_vel': self.joints_vel,
                'trans_vel': self.trans_vel, 'root_orient_vel': self.root_orient_vel}
        loss_obs, loss_nsteps = observed_data, T
        saved_ov = self.fitting_loss.loss_weights['rgb_overlap_consist']
        if nsteps is not None:
            loss_obs = {k: (v[:, :nsteps] if torch.is_tensor(v) else v) for k, v in observed_data.items()
                        if k != 'prev_batch_overlap_res'}
            loss_nsteps = nsteps
            self.fitting_loss.loss_weights['rgb_overlap_consist'] = 0.0
        loss, stats = self.fitting_loss.motion_fit(loss_obs, pred, cam_pred, loss_nsteps,
                                                   init_motion_scale=init_motion_scale)
        loss, stats = self._boundary_energy(loss, stats, cam_pred['verts3d'], loss_obs)
        self.fitting_loss.loss_weights['rgb_overlap_consist'] = saved_ov
        return loss, stats, roll, cam, cam_pred

    # ------------------------------------------------------------------------------------------------
    # CUDA-graphed closure
    # ------------------------------------------------------------------------------------------------
    def stage3_step(self, observed_data, nsteps=None, init_motion_scale=1.0, params=None):
        """One Stage-III closure evaluation: zero grads, forward, backward.  Returns the loss tensor (device);
        gradients land in ``.grad`` of the stage-3 variables that require grad.  With ``use_cuda_graph`` the whole
        evaluation (several hundred launches) is captured once per (phase, weights, trainable-set) and replayed."""
        params = self.stage3_params() if params is None else params
        if not self.use_cuda_graph:
            loss, grads, live = self._eval_on_aliases(observed_data, nsteps, init_motion_scale, params)
            for p in params:
                p.grad = None
            for p, gr in zip(live, grads):
                p.grad = torch.zeros_like(p) if gr is None else gr
            return loss
        from . import _ext
        w = self.fitting_loss.loss_weights
        for p in params:
            if p.requires_grad and p.grad is None:
                p.grad = torch.zeros_like(p)          # static gradient buffers the graph writes into
        packed = self.motion_prior.packed() if hasattr(self.motion_prior, 'packed') else None
        # a captured graph holds raw pointers: variables, their gradient buffers, the packed weights and the observations all
        # belong to its identity (and are kept alive by the cache entry below)
        key = (tuple(0 if p.grad is None else p.grad.data_ptr() for p in params), tuple(p.data_ptr() for p in params), id(packed),
               nsteps, float(init_motion_scale), getattr(self.motion_prior, 'precision', None),
               int(self.body_model.lbs_model.struct.use_umma), tuple(bool(p.requires_grad) for p in params),
               tuple(sorted((k, float(v)) for k, v in w.items())), tuple(id(p) for p in params),
               tuple((k, v.data_ptr()) for k, v in sorted(observed_data.items()) if torch.is_tensor(v)),
               tuple((k, v.data_ptr()) for k, v in sorted(observed_data.get('prev_batch_overlap_res', {}).items()) if torch.is_tensor(v)))
        g = self._graphs.get(key)
        if g is None:
            for p in params:
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)          # static gradient buffers the graph writes into
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                        # warm-up: allocator pools, lazy module state
                    self._eval_into_static(observed_data, nsteps, init_motion_scale, params)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = _ext.LaunchCounter.total
            if self._capture_stream is None:
                # high priority: the sequential decoder chain wins free SMs over the side-stream dense LBS pass
                self._capture_stream = torch.cuda.Stream(priority=-1)
            try:
                with torch.cuda.graph(graph, stream=self._capture_stream):
                    static_loss = self._eval_into_static(observed_data, nsteps, init_motion_scale, params)
            except RuntimeError as e:       # e.g. an injected pose prior that syncs or copies from the host
                import traceback
                import warnings
                where = traceback.format_exc().strip().splitlines()[-6:-1]
                warnings.warn(f'Stage-III closure is not CUDA-graph capturable ({e}); falling back to eager launches. '
                              f'At: {" | ".join(x.strip() for x in where)}')
                self.use_cuda_graph = False
                torch.cuda.synchronize()
                return self.stage3_step(observed_data, nsteps, init_motion_scale, params)
            g = (graph, static_loss, _ext.LaunchCounter.total - l0, list(params), [p.grad for p in params], packed,
                 {k: v for k, v in observed_data.items()})
            self._graphs[key] = g
        g[0].replay()
        _ext.LaunchCounter.total += g[2]
        return g[1]

    _STAGE3_NAMES = ('trans', 'root_orient', 'latent_pose', 'betas', 'latent_motion', 'trans_vel', 'joints_vel',
                     'root_orient_vel', 'floor_plane')

    def _eval_on_aliases(self, observed_data, nsteps, scale, params):
        """Forward + reverse on fresh detached ALIASES of the variables (same storage, new autograd leaves).
        The variables' own lazily created AccumulateGrad nodes stay out of the graph: they are bound to whatever
        stream first used them (the legacy default stream after an eager Stage I/II), and the engine would try to
        synchronise that stream inside a CUDA-graph capture (cudaErrorStreamCaptureImplicit)."""
        saved = {}
        alias_of = {}
        for n in self._STAGE3_NAMES:
            t = getattr(self, n, None)
            if torch.is_tensor(t):
                a = t.detach().requires_grad_(t.requires_grad)
                saved[n] = t
                alias_of[id(t)] = a
                setattr(self, n, a)
        try:
            self._defer_dense_join = True          # dense vertices overlap the reverse pass; joined below
            loss, _, _, _, _ = self.stage3_forward(observed_data, nsteps, scale)
            live = [p for p in params if p.requires_grad]
            grads = torch.autograd.grad(loss, [alias_of.get(id(p), p) for p in live], allow_unused=True)
        finally:
            self._defer_dense_join = False
            self.join_dense()
            for n, t in saved.items():
                setattr(self, n, t)
        return loss.detach(), grads, live

    def _eval_into_static(self, observed_data, nsteps, scale, params):
        loss, grads, live = self._eval_on_aliases(observed_data, nsteps, scale, params)
        for p, gr in zip(live, grads):
            if gr is None:
                p.grad.zero_()
            else:
                p.grad.copy_(gr)
        return loss

    def stage3_params(self):
        p = [self.trans, self.root_orient, self.latent_pose, self.betas, self.latent_motion,
             self.trans_vel, self.joints_vel, self.root_orient_vel]
        if self.optim_floor:
            p.append(self.floor_plane)
        return p

    def set_stage3_state(self, params):
        """Install stage-3 variables directly (dict of (B,1,3) trans/root_orient, (B,1,32) latent_pose, (B,16)
        betas, (B,T-1,48) latent_motion, (B,1,3|66|3) velocities, (B,3) floor_plane) as leaf tensors."""
        names = ['trans', 'root_orient', 'latent_pose', 'betas', 'latent_motion', 'trans_vel', 'joints_vel', 'root_orient_vel']
        if self.optim_floor:
            names.append('floor_plane')
        for n in names:
            setattr(self, n, torch.as_tensor(params[n], dtype=torch.float32, device=self.device).clone().requires_grad_(True))
        self.fitting_loss.set_stage(2)
        self._graphs.clear()                      # graphs captured for the previous variables point at their storage
        return names

    # ------------------------------------------------------------------------------------------------
    # initialisation helpers (once per batch; warm path)
    # ------------------------------------------------------------------------------------------------
    def initialize(self, observed_data):
        """motion_optimizer.py:141-199: floor from the observation, depth from bone-length ratios."""
        if not self.optim_floor:
            return
        fp = observed_data['floor_plane']
        self.floor_plane = (fp[:, :3] * fp[:, 3:]).to(torch.float).clone().detach().requires_grad_(True)
        if 'points3d' in observed_data:
            self.trans = torch.mean(observed_data['points3d'], dim=2).clone().detach()
        elif 'joints2d' in observed_data:
            with torch.no_grad():
                body_pose = self.latent2pose(self.latent_pose)
                pred, _ = self.smpl_results(self.trans, self.root_orient, body_pose, self.betas, dense=False, sel=False)
                j_op = pred['Jtr'][:, 0][:, self.smpl2op_map.tolist()]                  # (B,25,3)
                obs = observed_data['joints2d']
                xy, conf = obs[..., :2], obs[..., 2]
                best = (conf > 0.0).sum(2).max(1)[1]
                bidx = torch.arange(self.batch_size, device=obs.device)
                e0 = [e[0] for e in OP_EDGE_LIST]
                e1 = [e[1] for e in OP_EDGE_LIST]
                bone3d = (j_op[:, e0] - j_op[:, e1]).norm(dim=-1)                        # (B,E)
                bone2d = (xy[:, :, e0] - xy[:, :, e1]).norm(dim=-1)[bidx, best]          # (B,E)
                c2d = torch.min(conf[:, :, e0], conf[:, :, e1])[bidx, best]
                init_z = self.cam_f[:, 0] * (bone3d.mean(1) / (bone2d * (c2d > 0.0)).mean(1))
                self.trans[:, :, 2] = init_z[:, None].expand(self.batch_size, self.seq_len)

    @staticmethod
    def estimate_linear_velocity(x, h):
        """forward / central / backward differences (motion_optimizer.py:766-780)."""
        return torch.cat([(x[:, 1:2] - x[:, :1]) / h, (x[:, 2:] - x[:, :-2]) / (2 * h), (x[:, -1:] - x[:, -2:-1]) / h], 1)

    def estimate_angular_velocity(self, rot_seq, h):
        """angular velocity from dR/dt R^T (motion_optimizer.py:782-800)."""
        W = torch.matmul(self.estimate_linear_velocity(rot_seq, h), rot_seq.transpose(-1, -2))
        return torch.stack([(W[..., 2, 1] - W[..., 1, 2]) / 2.0, (W[..., 0, 2] - W[..., 2, 0]) / 2.0,
                            (W[..., 1, 0] - W[..., 0, 1]) / 2.0], -1)

    def estimate_velocities(self, trans, root_orient, body_pose, betas, data_fps, joints=None):
        B, T, _ = trans.size()
        h = 1.0 / data_fps
        if joints is None:
            joints = self.joints_only(trans, root_orient, body_pose, betas)
        R = batch_rodrigues(root_orient.reshape(-1, 3)).reshape(B, T, 3, 3)
        return self.estimate_linear_velocity(trans, h), self.estimate_linear_velocity(joints, h), \
            self.estimate_angular_velocity(R, h)

    def infer_latent_motion(self, trans, root_orient, body_pose, betas, data_fps, full_forward_pass=False):
        """Posterior-mean latent sequence of the current SMPL sequence (motion_optimizer.py:802-874)."""
        B, T, _ = trans.size()
        if self.optim_floor:
            pd = self.apply_cam2prior({'trans': trans, 'root_orient': root_orient}, self.cam2prior_R, self.cam2prior_t,
                                      self.cam2prior_root_height, body_pose, betas, self.init_fidx)
            trans, root_orient = pd['trans'], pd['root_orient']
        joints = self.joints_only(trans, root_orient, body_pose, betas)
        tv, jv, rv = self.estimate_velocities(trans, root_orient, body_pose, betas, data_fps, joints=joints)
        seq = {'trans': trans, 'trans_vel': tv,
               'root_orient': batch_rodrigues(root_orient.reshape(-1, 3)).reshape(B, T, 9), 'root_orient_vel': rv,
               'pose_body': batch_rodrigues(body_pose.reshape(-1, 3)).reshape(B, T, J_BODY * 9),
               'joints': joints.reshape(B, T, -1), 'joints_vel': jv.reshape(B, T, -1)}
        _, post = self.motion_prior.infer_global_seq(seq)
        return post[0]

    def get_optim_result(self, body_pose=None):
        if body_pose is None:
            body_pose = self.latent2pose(self.latent_pose)
        res = {'trans': self.trans.clone().detach(), 'root_orient': self.root_orient.clone().detach(),
               'pose_body': body_pose.clone().detach(), 'betas': self.betas.clone().detach(),
               'latent_pose': self.latent_pose.clone().detach(), 'latent_motion': self.latent_motion.clone().detach()}
        if self.optim_floor:
            res['floor_plane'] = parse_floor_plane(self.floor_plane).clone().detach()
        return res

    # ------------------------------------------------------------------------------------------------
    # the three stages
    # ------------------------------------------------------------------------------------------------
    def _lbfgs(self, params, lr, max_iter):
        """motion_optimizer.py:228-231,281-284,461-478.  One process: the library optimiser the reference uses (or, with
        ``lbfgs_impl = 'native'`` / HB_LBFGS=native, humor_b200.lbfgs with one packed host read per evaluation).  Sharded
        over ranks: the JOINT L-BFGS of the reference (one step length and curvature history for all sub-sequences,
        SURVEY.md 8e) needs global inner products, which only humor_b200.lbfgs provides."""
        sharded = self.shard is not None and self.shard.world > 1
        if sharded or self.lbfgs_impl == 'native':
            from .lbfgs import LBFGS
            group = (self.shard.group if self.shard.group is not None else True) if sharded else None
            return LBFGS(params, max_iter=max_iter, lr=lr, line_search_fn=LINE_SEARCH, group=group)
        return torch.optim.LBFGS(params, max_iter=max_iter, lr=lr, line_search_fn=LINE_SEARCH)

    def _boundary_energy(self, loss, stats, verts3d, observed_data, mode='motion'):
        """Overlap-consistency energy across the rank boundary (parallel.py) — added on the receiving rank only.
        Terms per stage as in the reference: root_fit key vertices (fitting_loss.py:136-157), smpl_fit + betas (:211-215),
        motion_fit + floor (:296-300)."""
        if self.shard is None or self.shard.world <= 1 or 'seq_interval' not in observed_data \
                or self.fitting_loss.loss_weights['rgb_overlap_consist'] <= 0.0:
            return loss, stats
        from .parallel import boundary_overlap_energy
        e, bstats = boundary_overlap_energy(self.shard, verts3d, self.betas,
                                            self.floor_plane if (self.optim_floor and mode == 'motion') else None,
                                            observed_data['seq_interval'], self.seq_len, with_betas=mode != 'root')
        loss = loss + self.fitting_loss.loss_weights['rgb_overlap_consist'] * e
        for k, v in bstats.items():
            stats[k] = stats[k] + v if k in stats else v
        return loss, stats

    def stage12_forward(self, observed_data, stage):
        """Body of the Stage-I (stage 0: root_fit) / Stage-II (stage 1: smpl_fit) closures of motion_optimizer.py:237-250,
        289-304 up to (loss, stats, pred).  The caller has selected the stage's weights (fitting_loss.set_stage)."""
        full = stage == 1
        body_pose = self.latent2pose(self.latent_pose)
        pred, _ = self.smpl_results(self.trans, self.root_orient, body_pose, self.betas, dense=False,
                                    dense_grad=self.points3d_active(observed_data))
        pred['betas'] = self.betas
        if full:
            pred['latent_pose'] = self.latent_pose
            loss, st = self.fitting_loss.smpl_fit(observed_data, pred, self.seq_len)
            loss, st = self._boundary_energy(loss, st, pred['verts3d'], observed_data, 'smpl')
        else:
            loss, st = self.fitting_loss.root_fit(observed_data, pred)
            loss, st = self._boundary_energy(loss, st, pred['verts3d'], observed_data, 'root')
        return loss, st, pred

    def _stage12(self, observed_data, stage, num_iter, lr, lbfgs_max_iter):
        """Stage I (root only) / Stage II (pose + shape): motion_optimizer.py:224-306."""
        self.fitting_loss.set_stage(stage)
        full = stage == 1
        self.trans.requires_grad_(True)
        self.root_orient.requires_grad_(True)
        self.betas.requires_grad_(full)
        self.latent_pose.requires_grad_(full)
        params = [self.trans, self.root_orient] + ([self.betas, self.latent_pose] if full else [])
        optim = self._lbfgs(params, lr, lbfgs_max_iter)
        for i in range(num_iter):
            self.fitting_loss.cur_optim_step = i

            def closure():
                optim.zero_grad()
                loss, _, _ = self.stage12_forward(observed_data, stage)
                loss.backward()
                return loss
            optim.step(closure)
        with torch.no_grad():
            body_pose = self.latent2pose(self.latent_pose)
            out, _ = self.smpl_results(self.trans, self.root_orient, body_pose, self.betas)
        return out, body_pose

    def run(self, observed_data, data_fps=30, lr=1.0, num_iter=[30, 70, 70], lbfgs_max_iter=20, stages_res_out=None,
            fit_gender='neutral'):
        if len(num_iter) != 3:
            raise ValueError('Must have num iters for 3 stages!')
        per_stage_outputs = {}
        self._graphs.clear()                      # a new run installs new variables: graphs of an earlier run are stale
        self.initialize(observed_data)
        per_stage_outputs['stage1'], body_pose = self._stage12(observed_data, 0, num_iter[0], lr, lbfgs_max_iter)
        self._save_stage(stages_res_out, 'stage1_results.npz', body_pose)
        per_stage_outputs['stage2'], body_pose = self._stage12(observed_data, 1, num_iter[1], lr, lbfgs_max_iter)
        self._save_stage(stages_res_out, 'stage2_results.npz', body_pose)
        stage2 = {'trans': self.trans.clone().detach(), 'root_orient': self.root_orient.clone().detach(),
                  'pose_body': body_pose.clone().detach(), 'betas': self.betas.clone().detach()}

        # ---- Stage III set-up (motion_optimizer.py:332-404)
        self.fitting_loss.set_stage(2)
        B, T = self.batch_size, self.seq_len
        with torch.no_grad():
            cur_body_pose = self.latent2pose(self.latent_pose)
            if self.optim_floor:
                j = self.joints_only(self.trans, self.root_orient, cur_body_pose, self.betas)
                self.cam2prior_R, self.cam2prior_t, self.cam2prior_root_height = compute_cam2prior(
                    self.floor_plane, self.trans[:, 0], self.root_orient[:, 0], j[:, 0])
            self.latent_motion = self.infer_latent_motion(self.trans, self.root_orient, cur_body_pose, self.betas, data_fps)
            vel_trans, vel_root = self.trans, self.root_orient
            if self.optim_floor:
                pd = self.apply_cam2prior({'trans': self.trans, 'root_orient': self.root_orient}, self.cam2prior_R,
                                          self.cam2prior_t, self.cam2prior_root_height, cur_body_pose, self.betas,
                                          self.init_fidx)
                vel_trans, vel_root = pd['trans'], pd['root_orient']
            tv, jv, rv = self.estimate_velocities(vel_trans, vel_root, cur_body_pose, self.betas, data_fps)
        leaf = lambda t: t.detach().clone().requires_grad_(True)
        self.latent_motion = leaf(self.latent_motion)
        self.trans_vel, self.joints_vel, self.root_orient_vel = leaf(tv[:, :1]), leaf(jv[:, :1].reshape(B, 1, -1)), leaf(rv[:, :1])
        self.trans, self.root_orient, self.latent_pose = leaf(self.trans[:, :1]), leaf(self.root_orient[:, :1]), leaf(self.latent_pose[:, :1])
        self.betas = leaf(self.betas)
        if self.optim_floor:
            self.floor_plane = leaf(self.floor_plane)
        # ---- initialisation stats (motion_optimizer.py:406-455): one roll-out of the initial state, before any Stage-III iteration
        with torch.no_grad():
            roll0, cam0 = self.rollout_latent_motion(self.trans, self.root_orient, self.latent2pose(self.latent_pose), self.betas,
                                                     [self.trans_vel, self.joints_vel, self.root_orient_vel], self.latent_motion,
                                                     fit_gender=fit_gender)
            init_pred, _ = self.smpl_results(cam0['trans'], cam0['root_orient'], cam0['pose_body'], self.betas)
            init_pred['contacts'] = roll0['contacts']
            per_stage_outputs['stage3_init'] = init_pred
            self._save_dict(stages_res_out, 'stage3_init_results.npz', self.betas, cam0['trans'], cam0['root_orient'], cam0['pose_body'],
                            contacts=roll0['contacts'], floor=self.floor_plane if self.optim_floor else None)
            if self.optim_floor:
                self._save_dict(stages_res_out, 'stage3_init_results_prior.npz', self.betas, roll0['trans'], roll0['root_orient'],
                                cam0['pose_body'], contacts=roll0['contacts'])
        init_params = [self.trans, self.root_orient, self.latent_pose, self.trans_vel, self.joints_vel, self.root_orient_vel]
        all_params = self.stage3_params()
        frozen = [self.latent_motion, self.betas] + ([self.floor_plane] if self.optim_floor else [])
        optim_all = self._lbfgs(all_params, lr, lbfgs_max_iter)
        optim_frozen = self._lbfgs(frozen, lr, lbfgs_max_iter) if self.stage3_tune_init_state else None
        optim_refine = self._lbfgs(all_params, lr, lbfgs_max_iter) if self.stage3_tune_init_state else None
        saved_ch = self.fitting_loss.loss_weights['contact_height']
        saved_cv = self.fitting_loss.loss_weights['contact_vel']
        nfr = self.stage3_tune_init_num_frames
        motion_optim, scale = optim_all, 1.0
        for i in range(num_iter[2]):
            tune = self.stage3_tune_init_state
            if tune and self.stage3_tune_init_freeze_start <= i < self.stage3_tune_init_freeze_end:
                motion_optim = optim_frozen
                for p in init_params:
                    p.requires_grad_(False)
                if self.stage3_contact_refine_only:
                    self.fitting_loss.loss_weights['contact_height'] = 0.0
                    self.fitting_loss.loss_weights['contact_vel'] = 0.0
                scale = float(T) / nfr
            elif tune and i >= self.stage3_tune_init_freeze_end:
                motion_optim = optim_refine
                for p in all_params:
                    p.requires_grad_(True)
                if self.stage3_contact_refine_only:
                    self.fitting_loss.loss_weights['contact_height'] = saved_ch
                    self.fitting_loss.loss_weights['contact_vel'] = saved_cv
                scale = float(T) / nfr
            nsteps = nfr if (tune and i < self.stage3_tune_init_freeze_start) else None

            def closure():
                return self.stage3_step(observed_data, nsteps, scale, all_params)
            motion_optim.step(closure)

        # ---- final rollout (motion_optimizer.py:612-676)
        with torch.no_grad():
            body_pose0 = self.latent2pose(self.latent_pose)
            roll, cam = self.rollout_latent_motion(self.trans, self.root_orient, body_pose0, self.betas,
                                                   [self.trans_vel, self.joints_vel, self.root_orient_vel],
                                                   self.latent_motion, fit_gender=fit_gender)
            body_pose = roll['pose_body']
            self.latent_pose = self.pose2latent(body_pose)
            self.trans, self.root_orient = cam['trans'], cam['root_orient']
            stage3, _ = self.smpl_results(self.trans, self.root_orient, body_pose, self.betas)
            stage3['prior_joints3d_rollout' if self.optim_floor else 'joints3d_rollout'] = roll['joints']
            stage3['contacts'] = roll['contacts']
            if self.optim_floor:
                stage3['prior_trans'], stage3['prior_root_orient'] = roll['trans'], roll['root_orient']
            per_stage_outputs['stage3'] = stage3
            final = self.get_optim_result(body_pose)
            final['contacts'] = roll['contacts']
            if self.optim_floor and stages_res_out is not None:
                # Stage-II result in the prior frame of the FINAL floor (motion_optimizer.py:650-674); stage3_results.npz itself is
                # written by the caller's save_optim_result (io_formats.py), with the 4-parameter floor plane
                j2 = self.joints_only(stage2['trans'], stage2['root_orient'], stage2['pose_body'], stage2['betas'])
                R2, t2, h2 = compute_cam2prior(self.floor_plane, stage2['trans'][:, 0], stage2['root_orient'][:, 0], j2[:, 0])
                p2 = self.apply_cam2prior(stage2, R2, t2, h2, stage2['pose_body'], stage2['betas'], self.init_fidx)
                self._save_dict(stages_res_out, 'stage2_results_prior.npz', self.betas, p2['trans'], p2['root_orient'], stage2['pose_body'])
        return final, per_stage_outputs

    def _save_dict(self, stages_res_out, fname, betas, trans, root_orient, pose_body, contacts=None, floor=None):
        """One npz per sub-sequence with the reference's keys (motion_optimizer.py:424-455, 665-674)."""
        if stages_res_out is None:
            return
        import os
        cpu = lambda t: t.clone().detach().cpu().numpy()
        b, tr, ro, bp = cpu(betas), cpu(trans), cpu(root_orient), cpu(pose_body)
        for i, path in enumerate(stages_res_out):
            d = {'betas': b[i], 'trans': tr[i], 'root_orient': ro[i], 'pose_body': bp[i]}
            if contacts is not None:
                d['contacts'] = cpu(contacts[i])
            if floor is not None:
                d['floor_plane'] = cpu(floor[i])
            np.savez(os.path.join(path, fname), **d)

    def _save_stage(self, stages_res_out, fname, body_pose, contacts=None):
        """per-stage npz dumps with the reference's keys (motion_optimizer.py:260-270,312-322)."""
        if stages_res_out is None:
            return
        import os
        cpu = lambda t: t.clone().detach().cpu().numpy()
        b, tr, ro, bp = cpu(self.betas), cpu(self.trans), cpu(self.root_orient), cpu(body_pose)
        for i, path in enumerate(stages_res_out):
            d = {'betas': b[i], 'trans': tr[i], 'root_orient': ro[i], 'pose_body': bp[i]}
            if contacts is not None:
                d['contacts'] = cpu(contacts[i])
            np.savez(os.path.join(path, fname), **d)
