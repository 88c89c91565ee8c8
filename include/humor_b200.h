/* humor_b200 — C-ABI of the B200-native HuMoR Stage-III hot path.
 *
 * The reference (davrempe/humor) has no FFI for this path: its de-facto plugin surface is the
 * duck-typed Python objects handed to MotionOptimizer (humor/fitting/run_fitting.py:385-406).
 * This header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); each
 * entry point names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t (>0) or HB_ERR_* (>1000) otherwise;
 *     no exceptions, no printf, no hidden allocation, no hidden synchronisation.
 *   - all pointers are DEVICE pointers to fp32 (or int32) arrays owned by the caller, including
 *     workspaces (query the size with the *_workspace_bytes function).
 *   - all work is enqueued on the given stream; functions are re-entrant per stream.
 *   - `launches` (nullable) receives the number of kernels the call enqueued.
 */
#ifndef HUMOR_B200_H_
#define HUMOR_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* hb_stream_t; /* == cudaStream_t */

/* return codes: 0, a cudaError_t (1..999), or one of */
#define HB_OK 0
#define HB_ERR_ARG 1001       /* NULL / out-of-range argument; nothing was launched */
#define HB_ERR_WORKSPACE 1002 /* workspace_bytes smaller than the *_workspace_bytes() query; nothing was launched */

#define HB_NUM_VERTS 6890
#define HB_NUM_JOINTS 52
#define HB_NUM_JOINTS_X 73 /* + 21 vertex-picked joints (smplx VertexJointSelector) */
#define HB_LBS_KF 208      /* feature row: betas 0:16 | pose feature 16:205 | pad */

/* ---------------------------------------------------------------------------------------------
 * SMPL+H model constants, packed by humor_b200/body_model.py (pack_smplh) from the npz the
 * reference loads in humor/body_model/body_model.py:37-58.
 * ------------------------------------------------------------------------------------------- */
typedef struct HbLbsModel {
  int num_verts;           /* V = 6890 */
  int v3_ld;               /* leading dimension of blend (>= 3V, multiple of 64) */
  int wk;                  /* ELL width of the skinning weights = max non-zeros per vertex */
  int flags;               /* HB_LBS_PLANES_TEMPLATE | HB_LBS_WEIGHTS_SUM_1 (below) */
  const float* v_template; /* [3V] */
  const float* blend;      /* [208][v3_ld]  row k: shapedirs (k<16) / posedirs (16<=k<205) of coord 3v+c */
  const float* blend_t;    /* [v3_ld][208]  transpose of blend */
  const float* j_template; /* [52*3]   J_regressor @ v_template */
  const float* j_dirs;     /* [52*3][16] J_regressor @ shapedirs */
  const int* w_idx;        /* [V][wk] joint index of each non-zero weight (padding: 0) */
  const float* w_val;      /* [V][wk] weight (padding: 0.0) */
  const int* parents;      /* [52] kintree_table[0], parents[0] = -1 */
  const int* extra_ids;    /* [21] smplx vertex_ids['smplh'] in VertexJointSelector order */
  /* tensor-core blend: hi/lo planes (x = hi + lo) of blend_t with K padded to 224: [v3_ld][224]; NULL / use_umma = 0
   * keeps the exact-fp32 FFMA kernels.  With HB_LBS_PLANES_TEMPLATE in flags, column 205 of these planes (a padding column of
   * blend_t; the pose kernels write feature 205 = 1 into the operand planes) holds v_template - likewise column 205 of
   * blend16a_* (times 2^10) - so that the products are v_posed itself and no epilogue adds the template.  The fused kernel
   * (skin form 3) REQUIRES the flag. */
  const float* blend_t_hi;
  const float* blend_t_lo;
  int use_umma;
  int max_depth;           /* deepest level of the kinematic tree (root = 0) */
  /* kinematic tree tables for the warp-per-frame kernels: depth of every joint, children in CSR form */
  const int* depth;        /* [52] */
  const int* child_start;  /* [53] */
  const int* child_list;   /* [51] */
  /* lane = frame group skinning (the epilogue of csrc/lbs_fuseg.cuh): groups of 8 consecutive vertices; per group the union of
     the joints its vertices are skinned to and, per joint, the 8 weights (0 where a vertex is not influenced) */
  const int* g_start;      /* [num_groups + 1] offsets into g_joint / g_w */
  const int* g_joint;      /* [E] joint * 12 */
  const float* g_w;        /* [E][8], 16-byte aligned */
  int num_groups;          /* 0: group tables absent */
  /* fused blend + group skinning (csrc/lbs_fuseg.cuh, skin form 3): 192-column tiles = 8 groups; the transforms of a tile's
     joints live in 13 shared-memory slots that persist across the consecutive column tiles a CTA walks */
  int ft_nct;              /* column tiles = ceil(num_groups / 8); 0: tables absent */
  const int* g_slot;       /* [E] byte offset of entry e's slot in the tile of its group, or -1: read A from global memory */
  const int* ft_tab;       /* [ft_nct][30] n_fresh, n_inc, bytes of the tile's record in ft_rec, 0, 13 fresh + 13 incremental
                              loads (joint*12 | slot << 16) */
  /* blend form 5 (skin form 3 only): blend_t * 2^10, all 208 columns padded to 256, as fp16 hi plane and UNSCALED fp16 lo
     plane (x = h + l) [v3_ld][256] each; NULL: form unavailable */
  const void* blend16a_h;
  const void* blend16a_l;
  /* selected-vertex set of the fitting energies (the key vertices of fitting_utils.KEYPT_VERTS, then the 21 vertex-picked
     joints): when humor_lbs_fwd / _bwd are called with vlist == sel_ids (the same device pointer) and nv == sel_nv, the skinning
     passes read the blend columns of those vertices from sel_blend [208][192] (slot s at columns 3s..3s+2; slots sel_nv ..
     sel_nv+20 = extra_ids) instead of gathering them from blend.  NULL / 0: no such set (any vlist works, gathered). */
  const int* sel_ids;      /* [sel_nv], sel_nv + 21 <= 64 */
  const float* sel_blend;
  int sel_nv;
  int ft_rec_stride;       /* bytes per column tile of ft_rec (multiple of 16, <= 64 + 48 * 96); 0: records absent */
  /* skin form 3: per column tile one contiguous skinning record, bulk-copied into shared memory by the kernel's producer:
     16 ints (entry offsets of the tile's 8 groups + end, relative to the tile's first entry; padding), then 48-byte entries
     { slot byte offset or -1, joint*12, 0, 0, 8 weights } - the contents of g_slot / g_joint / g_w in tile order */
  const void* ft_rec;      /* [ft_nct][ft_rec_stride], 16-byte aligned */
  /* blend form 5 when one shape serves >= 32 frames (frames_per_beta): the 189 POSE columns of blend_t * 2^10 (features 16..204)
     padded to 192, fp16 hi plane and unscaled lo plane [v3_ld][192]; template and shape blend are then added per sequence.
     NULL: the kernel keeps all columns in the product (blend16a_*) */
  const void* blend16p_h;
  const void* blend16p_l;
} HbLbsModel;
#define HB_LBS_PLANES_TEMPLATE 1 /* column 205 of blend_t_hi/lo and blend16a_h/l carries v_template (see above) */
#define HB_LBS_WEIGHTS_SUM_1 2   /* the skinning weights of every vertex sum to 1 (|sum - 1| < 1e-6): the dense pass may add the
                                    root translation to the transforms' translation column instead of to every vertex */

/* Replaces BodyModel.forward -> smplx.SMPLH.forward -> smplx.lbs.lbs
 * (humor/body_model/body_model.py:72-115).  N frames; betas row of frame n is n / frames_per_beta.
 *   vlist == NULL : all vertices, verts is [N][V][3];  else verts is [N][nv][3] for the listed ids;
 *   verts == NULL : no vertex output.    joints is [N][num_joints_out][3], num_joints_out in {52,73}. */
size_t humor_lbs_workspace_bytes(int N);
int humor_lbs_fwd(const HbLbsModel* m, int N, int frames_per_beta, const float* root_orient,
                  const float* pose_body, const float* betas, const float* trans, float* workspace,
                  size_t workspace_bytes, const int* vlist, int nv, float* verts, float* joints,
                  int num_joints_out, int64_t* launches, hb_stream_t stream);
/* Kernel forms of the DENSE tensor-core forward (results agree to fp32 rounding; 0 leaves a setting unchanged):
 *   skin_form   3 (default) blend + lane = frame group skinning fused in one persistent tcgen05 kernel (lbs_fuseg.cuh)
 *               1 blend GEMM (umma_gemm3_kernel, 128x128 tiles) into v_posed slabs + lane = vertex lbs_skin_apply_kernel: the
 *                 round-1 default; what a model without group tables falls back to
 *   blend_form  5 (default, with skin_form 3) every column as fp16 hi + lo planes, three products: the accuracy of form 1
 *                 (1e-6 m) from 4 instead of 8 bytes per operand element
 *               1 three TF32 passes on fp32 hi/lo planes
 *   slab_frames (skin_form 1) frames per v_posed slab kept in L2 between the two kernels (128..512)
 * Forms 2 (lane = frame skin pass / persistent 128x256 blend), blend 3 and 4 (single-pass pose columns) and the round-1 fused
 * kernel measured slower or less accurate on the B200 (profiles/r01*, r02a*, r02f*) and were removed: their numbers are refused.
 * Process-wide; not to be changed while a call is in flight.  Environment defaults: HB_LBS_SKIN, HB_LBS_BLEND, HB_LBS_SLAB. */
int humor_lbs_configure(int skin_form, int blend_form, int slab_frames);
/* The forms the most recent dense tensor-core call actually ran (a requested form falls back to form 1 when the model's
 * layout does not allow it); 0 before the first such call. */
int humor_lbs_forms_used(int* skin_form, int* blend_form);
/* CTAs of the fused dense forward from now on (0 = one persistent CTA per SM, the default).  More CTAs than SMs = shorter chunks of
 * the tile list per CTA: a pass queued on a second stream then fills the SMs other kernels leave idle instead of holding the chip. */
int humor_lbs_set_fuseg_ctas(int n);
/* Measurement: with enable != 0 the library brackets the fused kernel (lbs_fuseg_kernel, the dominant kernel of the dense forward) of
 * every following EAGER dense call with a CUDA event pair on the stream it is launched on.  Each call of this function first
 * collects the pending pair and returns the accumulated milliseconds / number of launches since timing was switched on; switching
 * it off (or on again) resets them.  Not for use while a stream capture is in progress. */
int humor_lbs_fuseg_timing(int enable, float* ms_sum, int* launches);
/* Reverse mode of the above (what autograd does through smplx in the reference).  d_verts follows the
 * same vlist convention; d_betas is per frame [N][16] (the caller reduces over frames_per_beta). */
int humor_lbs_bwd(const HbLbsModel* m, int N, int frames_per_beta, const float* root_orient,
                  const float* pose_body, const float* betas, const float* trans, float* workspace,
                  size_t workspace_bytes, const int* vlist, int nv, const float* d_verts,
                  const float* d_joints, int num_joints_out, float* d_root_orient, float* d_pose_body,
                  float* d_betas, float* d_trans, int64_t* launches, hb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * HuMoR CVAE weights, packed by humor_b200/humor_model.py (pack_humor_weights) from the state
 * dict keys decoder.net.{0,1,3,4,6,7,9}.* / prior_net.net.{0,1,3,...,12}.*
 * (humor/models/humor_model.py:181-206,1206-1229).  K dims zero-padded to the listed sizes.
 * ------------------------------------------------------------------------------------------- */
typedef struct HbHumorWeights {
  const float* dec_w[4];  /* [1024][416] [1024][1088] [512][1088] [216][576] */
  const float* dec_b[4];  /* [1024] [1024] [512] [216] */
  const float* dec_g[3];  /* GroupNorm gamma [1024] [1024] [512] */
  const float* dec_be[3]; /* GroupNorm beta */
  const float* dec_wt[4]; /* transposes: [416][1024] [1088][1024] [1088][512] [576][224] */
  const float* pri_w[5];  /* [1024][352] [1024][1024]x3 [96][1024] */
  const float* pri_b[5];
  const float* pri_g[4];
  const float* pri_be[4];
  const float* pri_wt[5]; /* [352][1024] [1024][1024]x3 [1024][96] */
  /* tensor-core path (batched prior AND sequential decoder steps): hi/lo operand planes (x = hi + lo, hi = top 11 mantissa bits) of
   * pri_w / pri_wt / dec_w / dec_wt for the 3xTF32 tcgen05 GEMM; use_umma = 0 keeps the exact-fp32 FFMA kernels */
  const float* pri_w_hi[5];
  const float* pri_w_lo[5];
  const float* pri_wt_hi[5];
  const float* pri_wt_lo[5];
  const float* dec_w_hi[4];
  const float* dec_w_lo[4];
  const float* dec_wt_hi[4];
  const float* dec_wt_lo[4];
  int use_umma;           /* 0 exact fp32 FFMA | 1 tcgen05 3xTF32 | 2 = 1 with the FORWARD decoder chain on fp16 hi/lo planes */
  int reserved;
  /* use_umma == 2: fp16 hi + scaled lo planes (x = h + l * 2^-11, csrc/umma_gemm16.cuh) of dec_w with K padded to a multiple
   * of 64: [1024][448] [1024][1088] [512][1088] [216][576] halves; NULL: mode 2 falls back to mode 1 */
  const void* dec_w16_h[4];
  const void* dec_w16_l[4];
  /* the same for the batched prior: pri_w as [1024][384] [1024][1024]x3 [96][1024] halves; NULL: the prior stays on 3xTF32 */
  const void* pri_w16_h[5];
  const void* pri_w16_l[5];
  /* persistent decoder chain (use_umma == 1; csrc/chain_persist.cuh): the z-skip rows of the four transposed decoder weights side
   * by side, [48][2784] = dec_wt[3][512:560][0:224] | dec_wt[2][1024:1072][0:512] | dec_wt[1][1024:1072][0:1024] |
   * dec_wt[0][339:387][0:1024], as hi/lo planes: d z of all steps is ONE batched GEMM over the reverse pass's operand tape.
   * NULL: the launch-per-layer chain runs instead */
  const float* dec_wz_hi;
  const float* dec_wz_lo;
} HbHumorWeights;

/* Replaces HumorModel.roll_out(x_past=None, init_input_dict, S, z_seq, return_prior=True)
 * (humor/models/humor_model.py:785-1017) for in_rot_rep='mat', out_rot_rep='aa', steps_in=1,
 * output_delta, 'smpl+joints+contacts'.
 *   init_state [B][339] = trans3|trans_vel3|root_orient9|root_orient_vel3|pose_body189|joints66|joints_vel66
 *   z_seq      [B][S][48]
 *   world      [S][B][348] world-frame outputs per step:
 *              trans3|trans_vel3|root_orient9|root_orient_vel3|pose_body189|joints66|joints_vel66|contacts9
 *   prior_out  [S][B][96]  (mean | log-variance) of the conditional prior at every step (nullable)
 * The workspace keeps the activations the reverse pass needs (the "tape"). */
size_t humor_rollout_workspace_bytes(int B, int S);
int humor_rollout_fwd(const HbHumorWeights* w, int B, int S, const float* init_state, const float* z_seq,
                      float* workspace, size_t workspace_bytes, float* world, float* prior_out,
                      int64_t* launches, hb_stream_t stream);
/* Diagnostics of the persistent decoder chain (csrc/chain_persist.cuh): when `buf` (device, >= S*5*16 int64) is set, the next
 * rollout launches record clock64 stamps of CTA 0 per step and phase (tools/chain_timeline.py); NULL switches it off. */
int humor_chain_debug(void* buf, size_t bytes);
/* Makes `side` wait for the point of the most recent humor_rollout_bwd right before its reverse decoder chain was launched
 * (cudaStreamWaitEvent on an event the library records there; capturable).  HB_ERR_ARG before the first humor_rollout_bwd. */
int humor_rollout_bwd_started_wait(hb_stream_t side);
/* BPTT through the rollout: d_world [S][B][348], d_prior_out [S][B][96] (nullable) ->
 * d_init [B][339], d_z [B][S][48].  Must follow humor_rollout_fwd on the same workspace. */
int humor_rollout_bwd(const HbHumorWeights* w, int B, int S, float* workspace, size_t workspace_bytes,
                      const float* d_world, const float* d_prior_out, float* d_init, float* d_z,
                      int64_t* launches, hb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Batched rotation conversions (humor/utils/transforms.py:139-170 batch_rodrigues,
 * :243-389 rotation_matrix_to_angle_axis) and their reverse modes.  n rotations.
 * ------------------------------------------------------------------------------------------- */
int humor_rodrigues_fwd(int n, const float* aa, float* R, hb_stream_t stream);
int humor_rodrigues_bwd(int n, const float* aa, const float* dR, float* daa, hb_stream_t stream);
int humor_mat2aa_fwd(int n, const float* R, float* aa, hb_stream_t stream);
int humor_mat2aa_bwd(int n, const float* R, const float* daa, float* dR, hb_stream_t stream);

/* Camera -> prior frame of B sub-sequences: replaces fitting_utils.compute_cam2prior (humor/fitting/fitting_utils.py:149-190) on a
 * (B,3) floor (normal * offset, parsed as fitting_utils.py:61-103) - what motion_optimizer.py:519-524 evaluates at the top of every
 * Stage-III closure.  trans0 / orient0 / joint0: frame-0 root translation, root orientation (axis-angle) and root joint of every
 * sequence, rows ld_* floats apart.  Outputs R [B][3][3] (rows right, forward, up), t [B][3] = -trans0, root_height [B].
 * Reverse: gR / gt / gh nullable (zero); all four gradients are [B][3]. */
int humor_cam2prior_fwd(int B, const float* floor_plane, const float* trans0, int ld_t, const float* orient0, int ld_r,
                        const float* joint0, int ld_j, float* R, float* t, float* root_height, hb_stream_t stream);
int humor_cam2prior_bwd(int B, const float* floor_plane, const float* trans0, int ld_t, const float* orient0, int ld_r,
                        const float* joint0, int ld_j, const float* gR, const float* gt, const float* gh, float* d_floor,
                        float* d_trans0, float* d_orient0, float* d_joint0, hb_stream_t stream);

/* Outputs of the latent roll-out in the layout the energies read: replaces the tensor shuffling of
 * MotionOptimizer.rollout_latent_motion after roll_out (humor/fitting/motion_optimizer.py:964-1019: matrix -> axis-angle of the
 * 22 rotations, concatenation with the frame-0 state, contact confidences / labels) and, when R / t are given, the camera-frame
 * root orientation and translation of apply_cam2prior(inverse=True) (:678-741).  world [S][B][348] as humor_rollout_fwd wrote it;
 * frame-0 state of every sub-sequence in the prior frame; outputs [B][T = S+1][.] except logits [B][S][9].
 * Reverse: every g_* is nullable (zero); d_world is written completely (velocity columns carry no gradient). */
int humor_rollout_outputs_fwd(int B, int S, const float* world, const float* trans0, const float* orient0, const float* pose0,
                              const float* joints0, const float* R, const float* t, const int* contact_idx, float thresh,
                              float* trans, float* orient, float* pose, float* joints, float* logits, float* conf, float* labels,
                              float* cam_trans, float* cam_orient, hb_stream_t stream);
int humor_rollout_outputs_bwd(int B, int S, const float* world, const float* trans0, const float* orient0, const float* R,
                              const float* g_trans, const float* g_orient, const float* g_pose, const float* g_joints,
                              const float* g_logits, const float* g_cam_trans, const float* g_cam_orient, float* d_world,
                              float* d_trans0, float* d_orient0, float* d_pose0, float* d_joints0, float* d_R, float* d_t,
                              hb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Stage-III energies (humor/fitting/fitting_loss.py:94-309 motion_fit/smpl_fit/root_fit and
 * the per-term functions :317-484, :504-518) — loss terms and their gradients in one pass.
 * ------------------------------------------------------------------------------------------- */
#define HB_NUM_TERMS 24
enum HbTerm {
  HB_T_JOINTS2D = 0, HB_T_JOINTS3D, HB_T_VERTS3D, HB_T_OV_POS, HB_T_OV_VEL, HB_T_POSE_PRIOR,
  HB_T_SHAPE_PRIOR, HB_T_SMOOTH, HB_T_OV_BETAS, HB_T_MOTION_PRIOR, HB_T_INIT_PRIOR, HB_T_JOINT_CONSIST,
  HB_T_BONE_LEN, HB_T_J3D_ROLLOUT, HB_T_CONTACT_VEL, HB_T_CONTACT_H, HB_T_FLOOR_REG, HB_T_OV_FLOOR
};

typedef struct HbFitArgs {
  int B, T;               /* sequences, frames actually rolled out (nsteps) */
  int njx;                /* joints per frame in cam_joints: 73 (with OpenPose extras) or 52 */
  float coef[HB_NUM_TERMS]; /* weight * scale of every term; 0 disables it (exactly zero grad) */
  float sigma2d;
  /* predictions (device) */
  const float* cam_joints;   /* [B][T][njx][3] camera-frame SMPL joints */
  const float* cam_verts;    /* [B][T][43][3]  camera-frame key vertices */
  const float* prior_joints; /* [B][T][22][3]  prior-frame SMPL joints */
  const float* roll_joints;  /* [B][T][22][3]  joints regressed by the rollout */
  const float* contact_logits; /* [B][T-1][9] */
  const float* betas;        /* [B][16] */
  const float* floor;        /* [B][3] (nullable) */
  const float* z;            /* [B][T-1][48] */
  const float* prior_out;    /* [T-1][B][96] (nullable -> standard normal) */
  const float* latent_pose;  /* [B][T][32] (nullable) */
  /* observations */
  const float* obs_joints2d; /* [B][T_obs][25][3] (nullable) */
  const float* obs_joints3d; /* [B][T_obs][22][3] (nullable) */
  const float* obs_verts3d;  /* [B][T_obs][43][3] (nullable) */
  const float* obs_floor;    /* [B][4] (nullable) */
  const int* seq_interval;   /* [B][2] (nullable) */
  const float* cam_f;        /* [B][2] */
  const float* cam_c;        /* [B][2] */
  int T_obs;                 /* frame stride of the observation tensors */
  /* outputs */
  float* terms;              /* [HB_NUM_TERMS] unweighted term values */
  float* loss;               /* [1] sum coef*term (+ init prior, added by the caller via coef) */
  float* d_cam_joints;       /* same shapes as the predictions; overwritten */
  float* d_cam_verts;
  float* d_prior_joints;
  float* d_roll_joints;
  float* d_contact_logits;
  float* d_betas;
  float* d_floor;
  float* d_z;
  float* d_prior_out;
  float* d_latent_pose;
  float* partials;           /* [B*T + 64][HB_NUM_TERMS] scratch for the deterministic two-level reduction (rows, then 64 block sums) */
} HbFitArgs;
int humor_fit_losses(const HbFitArgs* a, int64_t* launches, hb_stream_t stream);

/* Init-state GMM prior (fitting_loss.py:416-429 + torch MixtureSameFamily):
 *   x [B][D], logw [K], mean [K][D], Linv [K][D][D] (inverse Cholesky factor, lower), logdet [K]
 *   -> nll [B] and d_x [B][D] = d(sum nll)/dx.  D <= 160, K <= 32. */
int humor_gmm_nll(int B, int D, int K, const float* x, const float* logw, const float* mean,
                  const float* Linv, const float* logdet, float* nll, float* d_x, hb_stream_t stream);
/* The same with caller-owned scratch (humor_gmm_workspace_bytes): three launches over (component, 32-row chunk) blocks that stage
 * every Linv_k once per chunk instead of once per pair of rows - what the Stage-III closure calls. */
size_t humor_gmm_workspace_bytes(int B, int D, int K);
int humor_gmm_nll_ws(int B, int D, int K, const float* x, const float* logw, const float* mean, const float* Linv,
                     const float* logdet, float* nll, float* d_x, float* workspace, size_t workspace_bytes, hb_stream_t stream);

/* C = A[M,K] * B[N,K]^T (+bias) at fp32-level accuracy on the 5th-gen tensor cores (tcgen05, 3xTF32 operand split,
 * TMA-staged tiles).  K % 32 == 0, leading dimensions % 4 == 0.  The building block of the batched prior MLP. */
size_t humor_umma_gemm_workspace_bytes(int M, int N, int lda, int ldb);
int humor_umma_gemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N,
                    int K, float* workspace, size_t workspace_bytes, hb_stream_t stream);

/* The same product from FOUR-byte operand elements: fp16 hi + scaled fp16 lo planes (x = h + l * 2^-11), tcgen05 kind::f16, cross
 * terms in a second TMEM accumulator.  Operand ingest per SM is what bounds the GEMMs of this path; this is the building block
 * for halving it where fp16's range suffices (forward activations, weights).  K % 64 == 0, lda / ldb % 8 == 0, ldc % 4 == 0. */
size_t humor_umma_gemm16_workspace_bytes(int M, int N, int lda, int ldb);
int humor_umma_gemm16(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N,
                      int K, float* workspace, size_t workspace_bytes, hb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer nearest-neighbour search — the reference's only native module.  Replaces
 * cd.forward_cuda / cd.backward_cuda (humor/utils/chamfer_distance/chamfer_distance.cpp:26-55,180-185;
 * kernels chamfer_distance.cu:7-209), called by ChamferDistanceFunction (chamfer_distance.py:13-55) for
 * FittingLoss.points3d_loss (humor/fitting/fitting_loss.py:378-396).
 *   b clouds; xyz1 [b][n][3], xyz2 [b][m][3] (fp32, contiguous)
 *   dist1 [b][n], idx1 [b][n] (int32): squared distance to / index of the nearest point of xyz2, first minimum wins;
 *   dist2 [b][m], idx2 [b][m]: the other direction.  Passing dist2 == idx2 == NULL (or dist1 == idx1 == NULL)
 *   skips that direction (points3d_loss consumes only dist1).
 * Arithmetic and tie-breaking are those of the reference's CPU path (nnsearch, chamfer_distance.cpp:58-87):
 * results are bit-identical to it.  Unlike the reference (printf-only errors, chamfer_distance.cu:155-157)
 * launch failures are returned. */
int humor_chamfer_fwd(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1,
                      float* dist2, int* idx2, int64_t* launches, hb_stream_t stream);
/* Reverse mode (chamfer_distance.cpp:114-177): grad_dist1 [b][n] / grad_dist2 [b][m] (NULL: that direction carries no
 * gradient) -> grad_xyz1 [b][n][3], grad_xyz2 [b][m][3] (either may be NULL; both are overwritten, not accumulated).
 * Deterministic: one owner thread per destination point applies the contributions in the reference's CPU loop
 * order (the reference's CUDA path uses atomicAdd, chamfer_distance.cu:166-185). */
int humor_chamfer_bwd(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                      const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2,
                      int64_t* launches, hb_stream_t stream);

const char* humor_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HUMOR_B200_H_ */
