"""BodyModel — drop-in for humor/body_model/body_model.py:11-115 on the fused sm_100a LBS kernels.

Same constructor and ``forward`` signature, same output struct (``v f betas Jtr pose_body full_pose
pose_hand``), same attributes the reference's callers read (``model_type``, ``num_joints``,
``bm.faces_tensor``, ``use_vtx_selector``).  The arithmetic of smplx==0.1.28 (``lbs``,
``SMPLH.forward``, ``VertexJointSelector``) runs in csrc/lbs.cu; there is no CPU path.
Scope: SMPL+H (the only model HuMoR supports, run_fitting.py:356-358) with identity hands
(``pose_hand=None`` + ``flat_hand_mean=True``, body_model.py:56-57,82-83).
"""
import ctypes as C
import types

import numpy as np
import torch
import torch.nn as nn

from . import _ext

NUM_VERTS = 6890
NUM_JOINTS = 52
KF = 208
# smplx.vertex_ids.vertex_ids['smplh'] in VertexJointSelector order
EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
# humor/body_model/utils.py:17-19
KEYPT_VERTS = [4404, 920, 3076, 3169, 823, 4310, 1010, 1085, 4495, 4569, 6615, 3217, 3313, 6713,
               6785, 3383, 6607, 3207, 1241, 1508, 4797, 4122, 1618, 1569, 5135, 5040, 5691, 5636,
               5404, 2230, 2173, 2108, 134, 3645, 6543, 3123, 3024, 4194, 1306, 182, 3694, 4294, 744]


class Struct(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


def pack_smplh(asset, num_betas=16):
    """numpy packing of the model constants into the layouts of include/humor_b200.h (HbLbsModel)."""
    if num_betas > 16:
        raise NotImplementedError('humor_b200 LBS kernels are built for <= 16 betas (HuMoR uses 16)')
    vt = np.asarray(asset['v_template'], np.float64)
    V = vt.shape[0]
    if V != NUM_VERTS:
        raise ValueError(f'expected an SMPL+H mesh with {NUM_VERTS} vertices, got {V}')
    sd = np.asarray(asset['shapedirs'], np.float64)[:, :, :num_betas]          # (V,3,nb)
    pd = np.asarray(asset['posedirs'], np.float64)                            # (V,3,459)
    if pd.shape[-1] < 189:
        raise ValueError('posedirs must hold at least the 21 body joints (189 columns)')
    Jreg = np.asarray(asset['J_regressor'], np.float64)                       # (52,V)
    if Jreg.shape[0] != NUM_JOINTS:
        raise ValueError('expected a 52-joint SMPL+H regressor')
    W = np.asarray(asset['weights'], np.float64)                              # (V,52)
    v3_ld = ((3 * V + 63) // 64) * 64
    blend = np.zeros((KF, v3_ld), np.float32)
    blend[:num_betas, :3 * V] = sd.reshape(3 * V, num_betas).T
    blend[16:16 + 189, :3 * V] = pd[:, :, :189].reshape(3 * V, 189).T
    j_template = (Jreg @ vt).astype(np.float32)                                # (52,3)
    j_dirs = np.zeros((NUM_JOINTS * 3, 16), np.float32)
    j_dirs[:, :num_betas] = np.einsum('jv,vcl->jcl', Jreg, sd).reshape(NUM_JOINTS * 3, num_betas)
    nnz = (W != 0).sum(1)
    wk = int(max(1, nnz.max()))
    order = np.argsort(-np.abs(W), axis=1, kind='stable')[:, :wk]
    w_idx = order.astype(np.int32)
    w_val = np.take_along_axis(W, order, axis=1).astype(np.float32)
    par = np.asarray(asset['kintree_table'])[0].astype(np.int64).copy()
    par[0] = -1
    # kinematic tree tables (depth per joint, children in CSR order) for the level-parallel chain kernels
    depth = np.zeros(NUM_JOINTS, np.int32)
    for j in range(1, NUM_JOINTS):
        if not (0 <= par[j] < j):
            raise ValueError('kintree_table must list parents before children')
        depth[j] = depth[par[j]] + 1
    kids = [[c for c in range(1, NUM_JOINTS) if par[c] == j] for j in range(NUM_JOINTS)]
    child_start = np.cumsum([0] + [len(k) for k in kids]).astype(np.int32)
    child_list = np.asarray([c for k in kids for c in k], np.int32)
    # lane = frame group skinning (csrc/lbs_fuseg.cuh): groups of 8 consecutive vertices -> union of their joints + per-joint weight rows
    G = 8
    ng = (V + G - 1) // G
    Wf = W.astype(np.float32)
    g_start = np.zeros(ng + 1, np.int32)
    g_joint, g_w = [], []
    for g in range(ng):
        blk = Wf[g * G:(g + 1) * G]                                           # (<=8, 52)
        js = np.nonzero((blk != 0).any(0))[0]
        for j in js:
            row = np.zeros(G, np.float32)
            row[:blk.shape[0]] = blk[:, j]
            g_joint.append(j * 12)
            g_w.append(row)
        g_start[g + 1] = len(g_joint)
    g_joint = np.asarray(g_joint, np.int32)
    g_w = np.stack(g_w, 0) if g_w else np.zeros((1, G), np.float32)
    g_slot, ft_tab = fuseg_tables(g_start, g_joint, ng)
    ft_rec = fuseg_records(g_start, g_joint, g_slot, g_w, ng, ft_tab)
    return {
        'g_start': g_start, 'g_joint': g_joint, 'g_w': np.ascontiguousarray(g_w), 'num_groups': ng,
        'g_slot': g_slot, 'ft_tab': ft_tab, 'ft_nct': ft_tab.shape[0], 'ft_rec': ft_rec,
        'w_rows_sum_to_one': bool(np.abs(Wf.astype(np.float64).sum(1) - 1.0).max() < 1e-6),
        'depth': depth, 'child_start': child_start, 'child_list': child_list, 'max_depth': int(depth.max()),
        'num_verts': V, 'v3_ld': v3_ld, 'wk': wk,
        'v_template': vt.astype(np.float32).reshape(-1), 'blend': blend,
        'blend_t': np.ascontiguousarray(blend.T), 'j_template': j_template.reshape(-1), 'j_dirs': j_dirs,
        'w_idx': np.ascontiguousarray(w_idx), 'w_val': np.ascontiguousarray(w_val),
        'parents': par.astype(np.int32), 'extra_ids': np.asarray(EXTRA_VERTEX_IDS, np.int32),
    }


FG_NSLOT = 13          # csrc/lbs_fuseg.cuh: shared-memory slots for skinning transforms, [128 frames][12 floats] each
FG_GPT = 8             # vertex groups per 192-column tile (64 vertices)
FG_SLOT_BYTES = 128 * 48
FG_TAB = 4 + 2 * FG_NSLOT      # ints per column tile of ft_tab: n_fresh, n_inc, record bytes, 0, fresh loads, incremental loads
FG_REC_HEAD = 64               # bytes: 9 group offsets (entries, relative to the tile's first) + padding
FG_REC_ENTRY = 48              # bytes: slot byte offset | joint * 12 | 0 | 0 | 8 weights
FG_REC_MAX = FG_REC_HEAD + FG_REC_ENTRY * 96       # csrc/lbs_fuseg.cuh: one of the kernel's two record buffers (SMPL+H: <= 60 entries)


def fuseg_tables(g_start, g_joint, num_groups, nslot=FG_NSLOT, gpt=FG_GPT):
    """Static schedule of the transform slots of the fused blend + group-skinning kernel (csrc/lbs_fuseg.cuh).

    A CTA walks consecutive 64-vertex column tiles for the same 128 frames, so the 3x4 transforms of the joints a tile is
    skinned to are kept in ``nslot`` shared-memory slots across tiles.  Per tile c the table lists which (joint, slot) pairs to
    load when the CTA arrives from tile c-1 of the same frames ("inc") and when it starts at c ("fresh" = every slot the tile
    uses).  A joint keeps its slot only while consecutive tiles need it (so the state at c does not depend on where a CTA
    started), and a new joint never takes a slot tile c-1 used (its epilogue may still be reading).  Joints that find no slot
    are read from global memory by the epilogue (g_slot = -1): correct for any mesh, fast for SMPL-like locality.

    Returns g_slot [E] (byte offset of the slot of group entry e in ITS tile, or -1) and ft_tab [nct][4 + 2*nslot]:
    n_fresh, n_inc, bytes of the tile's skinning record (filled by fuseg_records), 0, fresh entries, inc entries;
    entry = joint*12 | slot << 16."""
    nct = (num_groups + gpt - 1) // gpt
    g_slot = np.full(len(g_joint), -1, np.int32)
    tab = np.zeros((nct, 4 + 2 * nslot), np.int32)
    prev = {}                                                   # joint*12 -> slot of tile c-1
    for c in range(nct):
        e0, e1 = int(g_start[min(c * gpt, num_groups)]), int(g_start[min((c + 1) * gpt, num_groups)])
        js, cnt = np.unique(g_joint[e0:e1], return_counts=True)
        order = [int(j) for j in js[np.argsort(-cnt, kind='stable')]]      # most used first
        cur = {j: prev[j] for j in order if j in prev}
        free = [s for s in range(nslot) if s not in prev.values()]
        inc = []
        for j in order:
            if j not in cur and free:
                cur[j] = free.pop(0)
                inc.append(j)
        fresh = sorted(cur, key=lambda j: cur[j])
        tab[c, 0], tab[c, 1] = len(fresh), len(inc)
        tab[c, 4:4 + len(fresh)] = [j | (cur[j] << 16) for j in fresh]
        tab[c, 4 + nslot:4 + nslot + len(inc)] = [j | (cur[j] << 16) for j in inc]
        for e in range(e0, e1):
            j = int(g_joint[e])
            if j in cur:
                g_slot[e] = cur[j] * FG_SLOT_BYTES
        prev = cur
    return g_slot, tab


def fuseg_records(g_start, g_joint, g_slot, g_w, num_groups, ft_tab, gpt=FG_GPT):
    """Per column tile ONE contiguous skinning record, which the kernel's producer copies into shared memory with a single bulk
    copy while the previous tile is skinned (csrc/lbs_fuseg.cuh): the epilogue warps then read joint lists and weights at
    shared-memory latency, warp-uniformly (the same entries straight from global memory missed the 28 KB of L1 the kernel
    leaves 3 times out of 4: 41 % of all stall cycles, profiles/r02g_fuseg35_set_full_details.txt).

    record = 16 ints (offsets of the tile's 8 groups + end, in entries relative to the tile's first entry; padding) followed by
    48-byte entries: slot byte offset (or -1) | joint * 12 | 0 | 0 | the 8 weights of the group's vertices.
    Returns a uint8 array [nct][stride] (stride = the longest record, a multiple of 16) and writes every tile's byte count into
    ft_tab[:, 2].  None when a tile has more entries than a record buffer holds (the dispatcher then takes skin form 1)."""
    nct = ft_tab.shape[0]
    gs = np.asarray(g_start, np.int64)
    ends = gs[np.minimum((np.arange(nct) + 1) * gpt, num_groups)]
    begs = gs[np.minimum(np.arange(nct) * gpt, num_groups)]
    emax = int((ends - begs).max())
    stride = FG_REC_HEAD + FG_REC_ENTRY * max(emax, 1)
    if stride > FG_REC_MAX:
        return None
    rec = np.zeros((nct, stride // 4), np.int32)
    gw = np.ascontiguousarray(g_w, np.float32).view(np.int32)
    for c in range(nct):
        e0, e1 = int(begs[c]), int(ends[c])
        for i in range(gpt + 1):
            rec[c, i] = int(gs[min(c * gpt + i, num_groups)]) - e0
        body = rec[c, FG_REC_HEAD // 4:FG_REC_HEAD // 4 + 12 * (e1 - e0)].reshape(e1 - e0, 12)
        body[:, 0], body[:, 1] = g_slot[e0:e1], g_joint[e0:e1]
        body[:, 4:12] = gw[e0:e1]
        ft_tab[c, 2] = FG_REC_HEAD + FG_REC_ENTRY * (e1 - e0)
    return rec.view(np.uint8).reshape(nct, stride)


def _tf32_rn(x):
    """x rounded to tf32 precision (to nearest, ties to even) — the hi plane of a 3xTF32 operand; x - hi is exact in fp32."""
    u = x.contiguous().view(torch.int32)
    return ((u + 0x0fff + ((u >> 13) & 1)) & -8192).view(torch.float32)


class LbsModel:
    """Device copy of the packed constants + the ctypes HbLbsModel handed to the C-ABI."""

    def __init__(self, packed, device):
        if torch.device(device).type != 'cuda':
            raise RuntimeError('humor_b200.BodyModel needs a CUDA device (there is no CPU path)')
        self._build(packed, device)

    def _build(self, packed, device):
        """Tables, operand planes and the C struct in `device` memory (tests/host/emul builds them in host memory for the
        kernel-executing CUDA runtime stand-in)."""
        self.device = torch.device(device)
        self.t = {k: torch.as_tensor(v).to(self.device).contiguous() for k, v in packed.items()
                  if isinstance(v, np.ndarray)}
        s = _ext.HbLbsModel()
        s.num_verts, s.v3_ld, s.wk, s.flags = packed['num_verts'], packed['v3_ld'], packed['wk'], 0
        for k in ('v_template', 'blend', 'blend_t', 'j_template', 'j_dirs', 'w_idx', 'w_val', 'parents', 'extra_ids'):
            setattr(s, k, self.t[k].data_ptr())
        # operand planes of the tensor-core blend GEMM: blend_t with K padded 208 -> 224, split x = hi + lo
        import os
        bt = torch.zeros(packed['v3_ld'], 224, device=self.device)
        bt[:, :KF] = self.t['blend_t']
        # column 205 (a zero padding column of blend_t; the pose kernels write feature 205 = 1 into the operand planes) carries
        # v_template: the tensor-core products return v_posed itself and no epilogue adds the template (HB_LBS_PLANES_TEMPLATE)
        bt[:self.t['v_template'].numel(), 205] = self.t['v_template']
        s.flags = 1 | (2 if packed.get('w_rows_sum_to_one') else 0)       # HB_LBS_PLANES_TEMPLATE | HB_LBS_WEIGHTS_SUM_1
        hi = _tf32_rn(bt)
        self.t['blend_t_hi'], self.t['blend_t_lo'] = hi.contiguous(), (bt - hi).contiguous()
        s.blend_t_hi, s.blend_t_lo = self.t['blend_t_hi'].data_ptr(), self.t['blend_t_lo'].data_ptr()
        s.use_umma = 0 if os.environ.get('HB_NO_UMMA') else 1
        s.g_start, s.g_joint, s.g_w = (self.t[k].data_ptr() for k in ('g_start', 'g_joint', 'g_w'))
        s.num_groups = packed['num_groups']
        s.g_slot, s.ft_tab, s.ft_nct = self.t['g_slot'].data_ptr(), self.t['ft_tab'].data_ptr(), packed['ft_nct']
        if packed.get('ft_rec') is not None:
            s.ft_rec, s.ft_rec_stride = self.t['ft_rec'].data_ptr(), packed['ft_rec'].shape[1]
        # blend form 5: blend_t * 2^10 (exact: keeps pose offsets down to 1e-7 m in fp16's normal range; the kernel's epilogue scales
        # back), every column, K padded to 256, as fp16 hi + UNSCALED fp16 lo plane (x = h + l)
        bs = torch.zeros(packed['v3_ld'], 256, device=self.device)
        bs[:, :224] = bt * 1024.0
        bh = bs.to(torch.float16)
        self.t['blend16a_h'], self.t['blend16a_l'] = bh.contiguous(), (bs - bh.float()).to(torch.float16).contiguous()
        # fp16's range is a property of the asset (SMPL+H in metres: |x * 2^10| < 1300): planes that would overflow are withheld and
        # humor_lbs_fwd then runs the fused kernel on the tf32 planes (blend form 1) instead
        fits16 = bool(bs.abs().max() < 6.0e4)
        if fits16:
            s.blend16a_h, s.blend16a_l = self.t['blend16a_h'].data_ptr(), self.t['blend16a_l'].data_ptr()
        # ... and the pose columns alone (features 16..204 -> 189 columns padded to 192) for calls with one shape per >= 32 frames:
        # template + shape blend are then added per SEQUENCE by the kernel's epilogue, K drops from 256 to 192
        bp = torch.zeros(packed['v3_ld'], 192, device=self.device)
        bp[:, :189] = self.t['blend_t'][:, 16:205] * 1024.0
        ph = bp.to(torch.float16)
        self.t['blend16p_h'], self.t['blend16p_l'] = ph.contiguous(), (bp - ph.float()).to(torch.float16).contiguous()
        if fits16:
            s.blend16p_h, s.blend16p_l = self.t['blend16p_h'].data_ptr(), self.t['blend16p_l'].data_ptr()
        self.ws_slot = 0
        s.max_depth = packed['max_depth']
        s.depth, s.child_start, s.child_list = (self.t[k].data_ptr() for k in ('depth', 'child_start', 'child_list'))
        self.struct = s
        self._ws = {}
        self._vlists = {}

    def workspace(self, N, slot=None):
        """Scratch for one C-ABI call.  ``slot`` separates calls that may be in flight on different streams."""
        key = (N, self.ws_slot if slot is None else slot)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _ext.lib().humor_lbs_workspace_bytes(N)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws[key] = ws
        return ws

    def vlist(self, ids):
        key = tuple(ids)
        t = self._vlists.get(key)
        if t is None:
            t = torch.tensor(list(ids), dtype=torch.int32, device=self.device)
            self._vlists[key] = t
            if self.struct.sel_nv == 0 and 0 < len(key) <= 43:
                self._select(t)
        return t

    def _select(self, ids):
        """The first short vertex list a caller asks for (the 43 key vertices of the fitting energies) becomes the model's selected
        set: its blend columns and those of the 21 vertex-picked joints are packed once as [208][192], and calls that pass this very
        tensor as vlist read them contiguously (csrc/lbs.cu, HbLbsModel.sel_blend)."""
        cols = torch.cat([ids.long(), self.t['extra_ids'].long()])
        tab = torch.zeros(self.t['blend'].shape[0], 192, device=self.device)
        src = (cols[:, None] * 3 + torch.arange(3, device=self.device)[None]).reshape(-1)
        tab[:, :src.numel()] = self.t['blend'][:, src]
        self.t['sel_ids'], self.t['sel_blend'] = ids, tab.contiguous()
        self.struct.sel_ids, self.struct.sel_blend, self.struct.sel_nv = ids.data_ptr(), self.t['sel_blend'].data_ptr(), ids.numel()


class _LbsFn(torch.autograd.Function):
    """(root_orient, pose_body, betas, trans) -> (v_dense | None, v_sel | None, joints).

    v_dense carries gradient only when ``dense_grad``; v_sel (the listed vertices) always does —
    the Stage-III energies touch 43 key vertices + 73 joints, so their reverse pass skins only those.
    """

    @staticmethod
    def forward(ctx, model, root_orient, pose_body, betas, trans, fpb, sel_ids, want_dense, dense_grad, njo):
        _ext.require_cuda(root_orient, pose_body, betas, trans)
        ro, pb, be, tr = (_ext.f32c(x) for x in (root_orient, pose_body, betas, trans))
        N = ro.shape[0]
        L = _ext.lib()
        ws = model.workspace(N)
        nl = C.c_int64(0)
        joints = torch.empty(N, njo, 3, device=ro.device, dtype=torch.float32)
        v_dense = v_sel = None
        sel = model.vlist(sel_ids) if sel_ids is not None else None
        if want_dense:
            v_dense = torch.empty(N, model.struct.num_verts, 3, device=ro.device, dtype=torch.float32)
            _ext.check(L.humor_lbs_fwd(C.byref(model.struct), N, fpb, _ext.ptr(ro), _ext.ptr(pb), _ext.ptr(be), _ext.ptr(tr),
                                       _ext.ptr(ws), ws.numel() * 4, None, 0, _ext.ptr(v_dense), _ext.ptr(joints), njo,
                                       C.byref(nl), _ext.stream_ptr()), 'humor_lbs_fwd')
            if sel is not None:
                v_sel = v_dense[:, sel.long()].contiguous()
        else:
            if sel is not None:
                v_sel = torch.empty(N, sel.numel(), 3, device=ro.device, dtype=torch.float32)
            _ext.check(L.humor_lbs_fwd(C.byref(model.struct), N, fpb, _ext.ptr(ro), _ext.ptr(pb), _ext.ptr(be), _ext.ptr(tr),
                                       _ext.ptr(ws), ws.numel() * 4, _ext.ptr(sel), 0 if sel is None else sel.numel(),
                                       _ext.ptr(v_sel), _ext.ptr(joints), njo, C.byref(nl), _ext.stream_ptr()), 'humor_lbs_fwd')
        _ext.LaunchCounter.total += nl.value
        ctx.model, ctx.fpb, ctx.sel, ctx.njo, ctx.dense_grad = model, fpb, sel, njo, dense_grad
        ctx.betas_shape = betas.shape
        ctx.save_for_backward(ro, pb, be, tr)
        ctx.set_materialize_grads(False)
        outs = []
        if v_dense is None:
            v_dense = torch.empty(0, device=ro.device)
        if v_sel is None:
            v_sel = torch.empty(0, device=ro.device)
        if not dense_grad:
            ctx.mark_non_differentiable(v_dense)
        return v_dense, v_sel, joints

    @staticmethod
    def backward(ctx, d_dense, d_sel, d_joints):
        ro, pb, be, tr = ctx.saved_tensors
        model, N = ctx.model, ro.shape[0]
        L = _ext.lib()
        ws = model.workspace(N)
        d_ro, d_pb, d_tr = torch.empty_like(ro), torch.empty_like(pb), torch.empty_like(tr)
        d_be = torch.empty(N, 16, device=ro.device, dtype=torch.float32)
        dj = _ext.f32c(d_joints) if d_joints is not None else None

        def run(vlist, nv, dv, djoints, out):
            nl = C.c_int64(0)
            _ext.check(L.humor_lbs_bwd(C.byref(model.struct), N, ctx.fpb, _ext.ptr(ro), _ext.ptr(pb), _ext.ptr(be), _ext.ptr(tr),
                                       _ext.ptr(ws), ws.numel() * 4, _ext.ptr(vlist), nv, _ext.ptr(dv), _ext.ptr(djoints),
                                       ctx.njo, _ext.ptr(out[0]), _ext.ptr(out[1]), _ext.ptr(out[2]), _ext.ptr(out[3]),
                                       C.byref(nl), _ext.stream_ptr()), 'humor_lbs_bwd')
            _ext.LaunchCounter.total += nl.value

        have_dense = ctx.dense_grad and d_dense is not None and d_dense.numel() > 0
        have_sel = d_sel is not None and d_sel.numel() > 0 and ctx.sel is not None
        outs = (d_ro, d_pb, d_be, d_tr)
        if have_dense:
            run(None, 0, _ext.f32c(d_dense), dj, outs)
            if have_sel:
                o2 = tuple(torch.empty_like(x) for x in outs)
                run(ctx.sel, ctx.sel.numel(), _ext.f32c(d_sel), None, o2)
                for a, b in zip(outs, o2):
                    a += b
        elif have_sel:
            run(ctx.sel, ctx.sel.numel(), _ext.f32c(d_sel), dj, outs)
        else:
            run(None, 0, None, dj, outs)
        nb = ctx.betas_shape[0]
        d_betas = d_be.view(nb, -1, 16).sum(1) if ctx.fpb > 1 else d_be
        d_betas = d_betas[:, :ctx.betas_shape[1]] if ctx.betas_shape[1] < 16 else d_betas
        return None, d_ro, d_pb, d_betas, d_tr, None, None, None, None, None


def lbs_dense_into(model, root_orient, pose_body, betas, trans, frames_per_beta, out):
    """Dense vertices of N frames, gradient-free, into the caller's (N, V, 3) tensor on the current stream (the pass a caller queues
    on a second stream, MotionOptimizer._launch_deferred_dense)."""
    if betas.shape[1] < 16:
        betas = torch.nn.functional.pad(betas, (0, 16 - betas.shape[1]))
    ro, pb, be, tr = (_ext.f32c(x) for x in (root_orient, pose_body, betas, trans))
    N = ro.shape[0]
    ws = model.workspace(N)
    joints = torch.empty(N, 52, 3, device=ro.device, dtype=torch.float32)
    nl = C.c_int64(0)
    _ext.check(_ext.lib().humor_lbs_fwd(C.byref(model.struct), N, frames_per_beta, _ext.ptr(ro), _ext.ptr(pb), _ext.ptr(be), _ext.ptr(tr),
                                        _ext.ptr(ws), ws.numel() * 4, None, 0, _ext.ptr(out), _ext.ptr(joints), 52, C.byref(nl),
                                        _ext.stream_ptr()), 'humor_lbs_fwd')
    _ext.LaunchCounter.total += nl.value
    return out


def lbs(model, root_orient, pose_body, betas, trans, frames_per_beta=1, sel_ids=None, want_dense=True,
        dense_grad=True, num_joints_out=52):
    if betas.shape[1] < 16:
        betas = torch.nn.functional.pad(betas, (0, 16 - betas.shape[1]))
    if root_orient.shape[0] != betas.shape[0] * frames_per_beta:
        raise ValueError('betas rows * frames_per_beta must equal the number of frames')
    return _LbsFn.apply(model, root_orient, pose_body, betas, trans, frames_per_beta, sel_ids, want_dense,
                        dense_grad, num_joints_out)


class BodyModel(nn.Module):
    """Wrapper with the reference's interface (body_model.py:16-115)."""

    def __init__(self, bm_path, num_betas=10, batch_size=1, num_expressions=10, use_vtx_selector=False,
                 model_type='smplh'):
        super().__init__()
        if model_type != 'smplh':
            raise NotImplementedError('Only SMPL+H is supported (as in HuMoR, run_fitting.py:356-358)')
        asset = bm_path if isinstance(bm_path, dict) else dict(np.load(bm_path, encoding='latin1'))
        self.model_type = model_type
        self.num_joints = 51                       # smplx SMPLH.NUM_JOINTS
        self.use_vtx_selector = use_vtx_selector
        self.num_betas = num_betas
        self.batch_size = batch_size
        self._packed = pack_smplh(asset, num_betas)
        self._model = None
        self.bm = nn.Module()
        self.bm.register_buffer('faces_tensor', torch.as_tensor(np.asarray(asset['f']).astype(np.int64)))

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        dev = self.bm.faces_tensor.device
        if dev.type == 'cuda' and (self._model is None or self._model.device != dev):
            self._model = LbsModel(self._packed, dev)
        return self

    def set_precision(self, mode):
        """'tensor': dense LBS forward blends on tcgen05 (3xTF32); 'exact': fp32 FFMA kernels.  Both keep vertices
        within ~2e-6 m of the fp64 oracle; the reverse pass is always exact fp32."""
        if mode not in ('tensor', 'exact', 'tensor16'):      # 'tensor16' concerns the motion prior's decoder chain only
            raise ValueError(mode)
        self.lbs_model.struct.use_umma = 0 if mode == 'exact' else 1

    @property
    def lbs_model(self):
        if self._model is None:
            raise RuntimeError('BodyModel must be moved to a CUDA device first (.to("cuda")); there is no CPU path')
        return self._model

    def forward(self, root_orient=None, pose_body=None, pose_hand=None, pose_jaw=None, pose_eye=None, betas=None,
                trans=None, dmpls=None, expression=None, return_dict=False, **kwargs):
        assert dmpls is None
        m = self.lbs_model
        dev = m.device
        given = [x for x in (root_orient, pose_body, betas, trans) if x is not None]
        N = given[0].shape[0] if given else self.batch_size
        z = lambda d: torch.zeros(N, d, device=dev, dtype=torch.float32)
        root_orient = z(3) if root_orient is None else root_orient
        pose_body = z(63) if pose_body is None else pose_body
        betas = z(self.num_betas) if betas is None else betas
        trans = z(3) if trans is None else trans
        if pose_hand is not None and bool((pose_hand != 0).any()):
            raise NotImplementedError('non-identity hand poses are outside the Stage-III path '
                                      '(the reference always passes pose_hand=None, motion_optimizer.py:1087-1092)')
        njo = 73 if self.use_vtx_selector else 52
        v, _, J = lbs(m, root_orient, pose_body, betas, trans, 1, None, True, True, njo)
        hands = z(90)
        out = {'v': v, 'f': self.bm.faces_tensor, 'betas': betas, 'Jtr': J, 'pose_body': pose_body,
               'full_pose': torch.cat([root_orient, pose_body, hands], 1), 'pose_hand': hands}
        return out if return_dict else Struct(**out)
