"""HumorModel — the HuMoR CVAE (humor/models/humor_model.py:100-1203) with its hot methods on
the sm_100a rollout kernels (csrc/rollout.cu).

Keeps the reference's constructor, attributes and state-dict keys (``encoder.net.N.*``,
``decoder.net.N.*``, ``prior_net.net.N.*``) so ``best_model.pth`` loads through the reference's
``load_state`` (utils/torch.py:44-82).  Native: ``roll_out`` (+reverse), ``prior``, ``decode``,
``sample_step``.  Scope: the configuration every fitting config ships (fitting/config.py:97-101):
``in_rot_rep='mat'``, ``out_rot_rep='aa'``, ``steps_in=1``, ``output_delta``, MLP archs,
``model_data_config='smpl+joints+contacts'`` / ``'smpl+joints'``.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _ext
from .transforms import compute_world2aligned_mat

STATE_D = 339
WORLD_D = 348
DATA_NAMES = ['trans', 'trans_vel', 'root_orient', 'root_orient_vel', 'pose_body', 'joints', 'joints_vel']
DATA_DIMS = [3, 3, 9, 3, 189, 66, 66]
WORLD_NAMES = DATA_NAMES + ['contacts']
WORLD_DIMS = DATA_DIMS + [9]


class MLP(nn.Module):
    """Parameter container with the reference's layout (humor_model.py:1206-1229):
    ``net = [Linear, (GroupNorm(16), ReLU, Linear)*]``; ``skip_input_idx`` features are re-concatenated
    before every later Linear.  ``forward`` (torch) exists for the cold posterior path only."""

    def __init__(self, layers, skip_input_idx=None):
        super().__init__()
        skip = 0 if skip_input_idx is None else layers[0] - skip_input_idx
        mods = [nn.Linear(layers[0], layers[1])]
        for i in range(1, len(layers) - 1):
            mods += [nn.GroupNorm(16, layers[i]), nn.ReLU(), nn.Linear(layers[i] + skip, layers[i + 1])]
        self.net = nn.ModuleList(mods)
        self.skip_input_idx = skip_input_idx

    def forward(self, x):
        skip = None if self.skip_input_idx is None else x[:, self.skip_input_idx:]
        for i, layer in enumerate(self.net):
            if skip is not None and i > 0 and isinstance(layer, nn.Linear):
                x = torch.cat([x, skip], 1)
            x = layer(x)
        return x

    def linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def norms(self):
        return [m for m in self.net if isinstance(m, nn.GroupNorm)]


def _pad_cols(w, k):
    out = torch.zeros(w.shape[0], k, dtype=torch.float32, device=w.device)
    out[:, :w.shape[1]] = w
    return out


class PackedWeights:
    """Zero-padded fp32 copies (+ transposes for the reverse pass) in the HbHumorWeights layout."""

    DEC_K = [416, 1088, 1088, 576]
    DEC_K16 = [448, 1088, 1088, 576]      # multiples of 64 halves
    PRI_K16 = [384, 1024, 1024, 1024, 1024]
    PRI_K = [352, 1024, 1024, 1024, 1024]

    def __init__(self, decoder, prior_net, device):
        self.keep = []

        def dev(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        def split(t):
            hi = (t.contiguous().view(torch.int32) & -8192).view(torch.float32)      # keep the top 11 mantissa bits
            return hi, t - hi

        def dev16(t):
            t = t.detach().to(device=device).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        def split16(t):
            """x = h + l * 2^-11 with h = fp16(x), l = fp16((x - h) * 2^11): operand planes of csrc/umma_gemm16.cuh"""
            h = t.to(torch.float16)
            return h, ((t - h.float()) * 2048.0).to(torch.float16)

        s = _ext.HbHumorWeights()
        dl, dn = decoder.linears(), decoder.norms()
        wts = []
        for i, lin in enumerate(dl):
            w = _pad_cols(lin.weight.detach().float(), self.DEC_K[i])
            s.dec_w[i] = dev(w)
            s.dec_b[i] = dev(lin.bias)
            wt = w.t().contiguous()                                  # [Kp][N]
            if i == 3:
                wt = _pad_cols(wt, 224)
            s.dec_wt[i] = dev(wt)
            wts.append(wt)
            (h, l), (ht, lt) = split(w), split(wt)
            s.dec_w_hi[i], s.dec_w_lo[i] = dev(h), dev(l)
            s.dec_wt_hi[i], s.dec_wt_lo[i] = dev(ht), dev(lt)
            h16, l16 = split16(_pad_cols(w, self.DEC_K16[i]))                    # precision 'tensor16' (forward chain only)
            s.dec_w16_h[i], s.dec_w16_l[i] = dev16(h16), dev16(l16)
        for i, gn in enumerate(dn):
            s.dec_g[i] = dev(gn.weight)
            s.dec_be[i] = dev(gn.bias)
        # persistent decoder chain: z-skip rows of the transposed weights side by side (d raw | d pre3 | d pre2 | d pre1 columns)
        wz = torch.cat([wts[3][512:560, :224], wts[2][1024:1072, :512], wts[1][1024:1072, :1024], wts[0][339:387, :1024]], 1).contiguous()
        wzh, wzl = split(wz)
        s.dec_wz_hi, s.dec_wz_lo = dev(wzh), dev(wzl)
        pl, pn = prior_net.linears(), prior_net.norms()
        for i, lin in enumerate(pl):
            w = _pad_cols(lin.weight.detach().float(), self.PRI_K[i])
            s.pri_w[i] = dev(w)
            s.pri_b[i] = dev(lin.bias)
            wt = w.t().contiguous()
            s.pri_wt[i] = dev(wt)
            (h, l), (ht, lt) = split(w), split(wt)
            s.pri_w_hi[i], s.pri_w_lo[i] = dev(h), dev(l)
            s.pri_wt_hi[i], s.pri_wt_lo[i] = dev(ht), dev(lt)
            h16, l16 = split16(_pad_cols(w, self.PRI_K16[i]))
            s.pri_w16_h[i], s.pri_w16_l[i] = dev16(h16), dev16(l16)
        for i, gn in enumerate(pn):
            s.pri_g[i] = dev(gn.weight)
            s.pri_be[i] = dev(gn.bias)
        import os
        s.use_umma = 0 if os.environ.get('HB_NO_UMMA') else 1
        s.reserved = 0
        self.struct = s
        self.device = device
        self._ws = {}
        self._gen = {}          # (B, S) -> number of forwards that wrote the cached tape (a reverse pass must see ITS forward's tape)

    def workspace(self, B, S):
        ws = self._ws.get((B, S))
        if ws is None:
            nbytes = _ext.lib().humor_rollout_workspace_bytes(B, S)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws[(B, S)] = ws
        return ws


# callables run right after the reverse roll-out's kernels were queued (MotionOptimizer queues its gradient-free dense LBS pass behind
# the reverse decoder chain there); each is called once and removed
AFTER_ROLLOUT_BWD = []


class _RolloutFn(torch.autograd.Function):
    """(init_state (B,339), z_seq (B,S,48)) -> (world (S,B,348), prior_out (S,B,96))."""

    @staticmethod
    def forward(ctx, pw, init_state, z_seq, want_prior):
        _ext.require_cuda(init_state, z_seq)
        x0, z = _ext.f32c(init_state), _ext.f32c(z_seq)
        B, S = z.shape[0], z.shape[1]
        ws = pw.workspace(B, S)
        world = torch.empty(S, B, WORLD_D, device=x0.device, dtype=torch.float32)
        prior = torch.empty(S, B, 96, device=x0.device, dtype=torch.float32) if want_prior else None
        nl = C.c_int64(0)
        _ext.check(_ext.lib().humor_rollout_fwd(C.byref(pw.struct), B, S, _ext.ptr(x0), _ext.ptr(z), _ext.ptr(ws),
                                                ws.numel() * 4, _ext.ptr(world), _ext.ptr(prior), C.byref(nl),
                                                _ext.stream_ptr()), 'humor_rollout_fwd')
        _ext.LaunchCounter.total += nl.value
        ctx.pw, ctx.B, ctx.S, ctx.want_prior = pw, B, S, want_prior
        pw._gen[(B, S)] = ctx.gen = pw._gen.get((B, S), 0) + 1
        ctx.set_materialize_grads(False)
        if prior is None:
            prior = torch.empty(0, device=x0.device)
        return world, prior

    @staticmethod
    def backward(ctx, d_world, d_prior):
        pw, B, S = ctx.pw, ctx.B, ctx.S
        dev = pw.device
        if pw._gen.get((B, S)) != ctx.gen:
            # the tape (step inputs, decoder outputs, GroupNorm statistics) lives in ONE workspace per (B, S): a later forward of
            # the same shape has overwritten what this reverse pass needs - wrong gradients, so refuse instead
            raise RuntimeError(f'humor_b200 rollout: the BPTT tape for (B={B}, S={S}) was overwritten by a later forward pass of the '
                               'same shape before this backward ran; call backward() before the next roll_out of that shape')
        ws = pw.workspace(B, S)
        if d_world is None:
            d_world = torch.zeros(S, B, WORLD_D, device=dev, dtype=torch.float32)
        d_world = _ext.f32c(d_world)
        dp = _ext.f32c(d_prior) if (ctx.want_prior and d_prior is not None and d_prior.numel() > 0) else None
        d_init = torch.empty(B, STATE_D, device=dev, dtype=torch.float32)
        d_z = torch.empty(B, S, 48, device=dev, dtype=torch.float32)
        nl = C.c_int64(0)
        _ext.check(_ext.lib().humor_rollout_bwd(C.byref(pw.struct), B, S, _ext.ptr(ws), ws.numel() * 4, _ext.ptr(d_world),
                                                _ext.ptr(dp), _ext.ptr(d_init), _ext.ptr(d_z), C.byref(nl),
                                                _ext.stream_ptr()), 'humor_rollout_bwd')
        _ext.LaunchCounter.total += nl.value
        while AFTER_ROLLOUT_BWD:
            AFTER_ROLLOUT_BWD.pop(0)()
        return None, d_init, d_z, None


class HumorModel(nn.Module):

    def __init__(self, in_rot_rep='aa', out_rot_rep='aa', latent_size=48, steps_in=1, conditional_prior=True,
                 output_delta=True, posterior_arch='mlp', decoder_arch='mlp', prior_arch='mlp',
                 model_data_config='smpl+joints+contacts', detach_sched_samp=True,
                 model_use_smpl_joint_inputs=False, model_smpl_batch_size=1):
        super().__init__()
        if (in_rot_rep, out_rot_rep) != ('mat', 'aa') or steps_in != 1 or not output_delta or latent_size != 48 \
                or not conditional_prior or model_use_smpl_joint_inputs \
                or model_data_config not in ('smpl+joints+contacts',) \
                or (posterior_arch, decoder_arch, prior_arch) != ('mlp', 'mlp', 'mlp'):
            raise NotImplementedError(
                'humor_b200.HumorModel implements the configuration the fitting path ships '
                "(in_rot_rep='mat', out_rot_rep='aa', latent 48, steps_in=1, output_delta, conditional prior, "
                "'smpl+joints+contacts'; humor/fitting/config.py:97-101)")
        self.ignore_keys = []
        self.in_rot_rep, self.out_rot_rep = in_rot_rep, out_rot_rep
        self.latent_size, self.steps_in, self.steps_out = latent_size, steps_in, 1
        self.use_conditional_prior = conditional_prior
        self.output_delta = output_delta
        self.model_data_config = model_data_config
        self.data_names = list(DATA_NAMES)
        self.aux_out_data_names = ['contacts']
        self.pred_contacts = True
        self.need_trans2joint = True
        self.input_dim_list = list(DATA_DIMS)
        self.input_data_dim = STATE_D
        self.output_data_dim = 216
        self.use_smpl_joint_inputs = False
        self.posterior_arch, self.decoder_arch, self.prior_arch = posterior_arch, decoder_arch, prior_arch
        self.detach_sched_samp = detach_sched_samp
        self.encoder = MLP([2 * STATE_D, 1024, 1024, 1024, 1024, 2 * latent_size])
        self.decoder = MLP([STATE_D + latent_size, 1024, 1024, 512, 216], skip_input_idx=STATE_D)
        self.prior_net = MLP([STATE_D, 1024, 1024, 1024, 1024, 2 * latent_size])
        self._packed = None
        import os
        self.precision = 'exact' if os.environ.get('HB_NO_UMMA') else 'tensor'

    # 'tensor16' (opt-in, not yet measured on hardware): 'tensor' with the FORWARD decoder chain on fp16 hi + scaled lo operand
    # planes (4 bytes per element instead of 8; csrc/umma_gemm16.cuh) - same 22-bit operand significand, same tape for the reverse
    _UMMA_MODE = {'exact': 0, 'tensor': 1, 'tensor16': 2}

    def set_precision(self, mode):
        """'tensor': every GEMM on tcgen05 (3xTF32 split, fp32 promotion) — forward states / log-prob within ~2e-6 of
        fp64, gradients through the 59-step reverse pass within ~3e-3 (measured; the BPTT amplifies the tensor core's
        ~1e-6-of-sum(|a||b|) product error).  'exact': fp32 FFMA kernels — gradients within ~2e-6, as the fp32 reference."""
        if mode not in ('tensor', 'exact', 'tensor16'):
            raise ValueError(mode)
        self.precision = mode
        if self._packed is not None:
            self._packed.struct.use_umma = self._UMMA_MODE[mode]

    # -- weights -----------------------------------------------------------------------------------
    def packed(self):
        dev = self.decoder.net[0].weight.device
        if dev.type != 'cuda':
            raise RuntimeError('HumorModel must live on a CUDA device (there is no CPU path)')
        if self._packed is None or self._packed.device != dev:
            self._packed = PackedWeights(self.decoder, self.prior_net, dev)
        self._packed.struct.use_umma = self._UMMA_MODE[self.precision]
        return self._packed

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # -- helpers -----------------------------------------------------------------------------------
    @staticmethod
    def _split(x, names, dims):
        out, s = {}, 0
        for n, d in zip(names, dims):
            out[n] = x[..., s:s + d]
            s += d
        return out

    def split_output(self, decoder_out, convert_rots=True):
        """humor_model.py:316-349 for output_delta (rotations already matrices)."""
        B = decoder_out.size(0)
        return self._split(decoder_out.reshape(B, 1, -1), WORLD_NAMES, WORLD_DIMS)

    # -- native single-step entry points ---------------------------------------------------------------
    def _one_step(self, past_in, z, want_prior):
        """One rollout step through the fused kernels: returns decode() output (B,348) and prior."""
        B = past_in.shape[0]
        world, prior = _RolloutFn.apply(self.packed(), past_in, z.reshape(B, 1, 48), want_prior)
        return world[0], (prior[0] if want_prior else None)

    def prior(self, past_in):
        """humor_model.py:407-418 -> (mean, var)."""
        z0 = torch.zeros(past_in.shape[0], 48, device=past_in.device)
        _, po = self._one_step(past_in, z0, True)
        return po[:, :48], torch.exp(po[:, 48:])

    def decode(self, z, past_in):
        """humor_model.py:445-498.  With Gr=I, Gt=0 and trans2joint cancelling, the world-frame output of
        the first rollout step IS decode()'s output."""
        out, _ = self._one_step(past_in, z, False)
        return out

    def posterior(self, past_in, t_in):
        """humor_model.py:420-435 (Stage-III initialisation only; plain torch)."""
        o = self.encoder(torch.cat([past_in, t_in], 1))
        return o[:, :self.latent_size], torch.exp(o[:, self.latent_size:])

    def rsample(self, mu, var):
        return mu + torch.randn_like(mu) * torch.sqrt(var)

    def sample_step(self, past_in, t_in=None, use_mean=False, z=None, return_prior=False, return_z=False):
        """humor_model.py:1019-1059."""
        B = past_in.size(0)
        pm = pv = None
        if t_in is not None:
            pm, pv = self.posterior(past_in, t_in)
        if z is None:
            if pm is None:
                pm, pv = self.prior(past_in)
            z = pm if use_mean else self.rsample(pm, pv)
        dec, po = self._one_step(past_in, z, return_prior and t_in is None)
        out = {'decoder_out': dec.reshape(B, 1, -1)}
        if return_prior:
            out['prior'] = (pm, pv) if pm is not None else (po[:, :48], torch.exp(po[:, 48:]))
        if return_z:
            out['z'] = z
        return out

    # -- roll_out ------------------------------------------------------------------------------------
    def roll_out_raw(self, init_state, z_seq, return_prior=True):
        """Fused rollout: init_state (B,339), z_seq (B,S,48) -> world (S,B,348), prior_out (S,B,96)|None."""
        world, prior = _RolloutFn.apply(self.packed(), init_state, z_seq, return_prior)
        return world, (prior if return_prior else None)

    def roll_out(self, x_past, init_input_dict, num_steps, use_mean=False, z_seq=None, return_prior=False,
                 gender=None, betas=None, return_z=False, canonicalize_input=False, uncanonicalize_output=False):
        """humor_model.py:785-1017.  Returns the world-frame dict of (B, num_steps, D) tensors
        (rotations as matrices) and, with return_prior, (pm, pv)."""
        d = {k: v for k, v in init_input_dict.items()}
        B = d['trans'].shape[0]
        w2a_R = w2a_t = None
        if canonicalize_input:
            R0 = d['root_orient'][:, -1].reshape(B, 3, 3)
            w2a_R = compute_world2aligned_mat(R0)
            zero = torch.zeros(B, 1, device=R0.device, dtype=R0.dtype)
            w2a_t = torch.cat([-d['trans'][:, -1, :2], zero], 1)
            t2j = -torch.cat([d['joints'][:, -1, :2] + w2a_t[:, :2], zero], 1)
            d = self.apply_world2local_trans(w2a_t, w2a_R, t2j, d, dict(), invert=False)
        if x_past is not None and not canonicalize_input:
            init_state = x_past[:, -1].reshape(B, STATE_D)
        else:
            init_state = torch.cat([d[k][:, -1] for k in DATA_NAMES], 1)
        if z_seq is None:
            return self._roll_out_sampling(init_state, num_steps, use_mean, return_prior, return_z,
                                           canonicalize_input and uncanonicalize_output, w2a_R, w2a_t)
        world, prior = self.roll_out_raw(init_state, z_seq[:, :num_steps], return_prior)
        out = self._split(world.permute(1, 0, 2), WORLD_NAMES, WORLD_DIMS)
        if canonicalize_input and uncanonicalize_output:
            out = self._uncanonicalize(out, init_state, w2a_R, w2a_t)
        if return_z:
            out['z'] = z_seq[:, :num_steps]
        if return_prior:
            pm = prior[..., :48].permute(1, 0, 2)
            pv = torch.exp(prior[..., 48:]).permute(1, 0, 2)
            return out, (pm, pv)
        return out

    def _uncanonicalize(self, out, init_state, R, t):
        """Outputs of a canonicalised rollout back in the caller's frame (humor_model.py:856-859:
        the running world2local transform starts from the canonicalisation instead of identity)."""
        B = init_state.shape[0]
        zero = torch.zeros(B, 1, device=init_state.device, dtype=init_state.dtype)
        t2j = -torch.cat([init_state[:, 207:209], zero], 1)
        res = self.apply_world2local_trans(t, R, t2j, {k: v for k, v in out.items()}, dict(), invert=True)
        return res

    def _roll_out_sampling(self, init_state, num_steps, use_mean, return_prior, return_z, uncanon, R, t):
        """z_seq=None: the latent of step t depends on the prior at step t, so steps are issued one by
        one (sampling is not on the Stage-III path)."""
        B = init_state.shape[0]
        pw = self.packed()
        zs, pms, pvs = [], [], []
        z_so_far = torch.zeros(B, 0, 48, device=init_state.device)
        world = None
        for s in range(num_steps):
            # re-run the fused rollout on the prefix (keeps exact roll_out semantics; O(S^2) but cold path)
            zcat = torch.cat([z_so_far, torch.zeros(B, 1, 48, device=init_state.device)], 1)
            _, prior = _RolloutFn.apply(pw, init_state, zcat, True)
            pm, pv = prior[s, :, :48], torch.exp(prior[s, :, 48:])
            z = pm if use_mean else self.rsample(pm, pv)
            z_so_far = torch.cat([z_so_far, z[:, None]], 1)
            pms.append(pm)
            pvs.append(pv)
        world, _ = _RolloutFn.apply(pw, init_state, z_so_far, False)
        out = self._split(world.permute(1, 0, 2), WORLD_NAMES, WORLD_DIMS)
        if uncanon:
            out = self._uncanonicalize(out, init_state, R, t)
        if return_z:
            out['z'] = z_so_far
        if return_prior:
            return out, (torch.stack(pms, 1), torch.stack(pvs, 1))
        return out

    # -- frame changes on dicts (cold paths) -------------------------------------------------------------
    def apply_world2local_trans(self, world2local_trans, world2local_rot, trans2joint, input_dict, output_dict,
                                invert=False):
        """humor_model.py:696-772 on (B,S,D) tensors."""
        B = world2local_trans.size(0)
        R = world2local_rot.reshape(B, 1, 3, 3)
        t = world2local_trans.reshape(B, 1, 3)
        t2j = trans2joint.reshape(B, 1, 1, 3)
        M = R.transpose(3, 2) if invert else R
        rot = lambda v: torch.matmul(M[:, :, None], v[..., None])[..., 0]
        for k, v in input_dict.items():
            S = v.size(1)
            if k == 'root_orient':
                output_dict[k] = torch.matmul(M, v.reshape(B, S, 3, 3)).reshape(B, S, 9)
            elif k == 'trans':
                output_dict[k] = (rot(v.reshape(B, S, 1, 3))[:, :, 0] - t) if invert else \
                    rot((v + t).reshape(B, S, 1, 3))[:, :, 0]
            elif k in ('joints', 'verts'):
                p = v.reshape(B, S, -1, 3)
                if invert:
                    o = rot(p + t2j) - t2j - t.reshape(B, 1, 1, 3)
                else:
                    o = rot(p + t.reshape(B, 1, 1, 3) + t2j) - t2j
                output_dict[k] = o.reshape(B, S, -1)
            elif k in ('joints_vel', 'verts_vel'):
                output_dict[k] = rot(v.reshape(B, S, -1, 3)).reshape(B, S, -1)
            elif k in ('trans_vel', 'root_orient_vel'):
                output_dict[k] = rot(v.reshape(B, S, 1, 3))[:, :, 0]
            else:
                output_dict[k] = v
        return output_dict

    def infer_global_seq(self, global_seq, full_forward_pass=False):
        """humor_model.py:1061-1162 — posterior/prior over a ground-truth sequence, each step
        canonicalised on its own (Stage-III initialisation, run once per batch; plain torch + our prior)."""
        if full_forward_pass:
            raise NotImplementedError('full_forward_pass is a training-time diagnostic, outside the fitting path')
        B, T = global_seq['trans'].shape[:2]
        dev = global_seq['trans'].device
        zero = torch.zeros(B * (T - 1), 1, device=dev)
        flat = lambda k, sl: global_seq[k][:, sl].reshape(B * (T - 1), 1, -1)
        past = {k: flat(k, slice(0, T - 1)) for k in DATA_NAMES}
        nxt = {k: flat(k, slice(1, T)) for k in DATA_NAMES}
        R0 = past['root_orient'][:, 0].reshape(-1, 3, 3)
        Ra = compute_world2aligned_mat(R0)
        ta = torch.cat([-past['trans'][:, 0, :2], zero], 1)
        # trans2joint comes from the FIRST frame only and is reused for every step (:1085-1088)
        t2j0 = -torch.cat([global_seq['joints'][:, 0, :2] - global_seq['trans'][:, 0, :2],
                           torch.zeros(B, 1, device=dev)], 1)
        t2j = t2j0[:, None].expand(B, T - 1, 3).reshape(B * (T - 1), 3)
        lp = self.apply_world2local_trans(ta, Ra, t2j, past, dict(), invert=False)
        ln = self.apply_world2local_trans(ta, Ra, t2j, nxt, dict(), invert=False)
        past_in = torch.cat([lp[k][:, 0] for k in DATA_NAMES], 1)
        t_in = torch.cat([ln[k][:, 0] for k in DATA_NAMES], 1)
        pm, pv = self.prior(past_in)
        qm, qv = self.posterior(past_in, t_in)
        r = lambda x: x.reshape(B, T - 1, -1)
        return (r(pm), r(pv)), (r(qm), r(qv))

    def forward(self, x_past, x_t):
        raise NotImplementedError('training forward is outside the Stage-III fitting path (SURVEY.md §8)')
