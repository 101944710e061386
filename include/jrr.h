/*
 * jrr.h -- C ABI of libjrr.so, the B200-native (sm_100a) replacement for the
 * data-parallel hot path of ubc-vision/joint-regressor-refinement.
 *
 * Conventions
 *   - every entry point returns 0 on success, a JrrStatus otherwise; the message of the
 *     last failure on the calling thread is available from jrr_last_error()
 *   - plain pointers and sizes only (no torch / C++ types); all data pointers are DEVICE
 *     pointers unless the name ends in _host
 *   - no entry point allocates caller-visible memory, synchronises the device or calls
 *     back into the host: work is enqueued on the cudaStream_t handed in (passed as
 *     void*), so sequences of calls can be captured into a CUDA graph
 *   - "B" is the number of poses (frames) of a call, "BP" = B rounded up to 128; device
 *     workspaces are sized by jrr_workspace_bytes() and owned by the caller
 *
 * Each declaration cites the reference interface (file:line under /root/reference) it
 * replaces.  The reference is pure Python; INTEGRATION.md shows the ctypes stub a
 * maintainer adds to call these from scripts/smpl.py, scripts/utils.py and
 * scripts/optimize.py.
 */
#ifndef JRR_H_
#define JRR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JRR_ABI_VERSION 2

#define JRR_NUM_VERTS 6890
#define JRR_NUM_JOINTS 24
#define JRR_NUM_BETAS 10
#define JRR_NUM_POSE_FEATS 207
#define JRR_NUM_H36M 17
#define JRR_NUM_EXTRA 9
#define JRR_NUM_PICKS 21
#define JRR_NUM_OUT_JOINTS 49
#define JRR_CRITIC_PARAMS 1840153
#define JRR_SHAPE_CRITIC_PARAMS 171
#define JRR_LOSS_TERMS 5 /* length of every refine-step loss_out: {total, joint, pose, 2d, shape} */

typedef enum JrrStatus {
  JRR_OK = 0,
  JRR_ERR_INVALID = 1,   /* bad argument / shape */
  JRR_ERR_CUDA = 2,      /* a CUDA runtime/driver call failed */
  JRR_ERR_STATE = 3,     /* e.g. regressor or critic not set */
  JRR_ERR_WORKSPACE = 4  /* workspace too small */
} JrrStatus;

/* how the 24 joint rotations of a call are encoded */
typedef enum JrrPoseKind {
  JRR_POSE_ROTMAT = 0,     /* [B,24,3,3] row-major (pose2rot=False; utils.py:94-95) */
  JRR_POSE_AXIS_ANGLE = 1, /* [B,24,3]   (pose2rot=True, smplx default; smpl.py:72-74) */
  JRR_POSE_ROT6D = 2       /* [B,24,6]   interleaved view(-1,3,2) layout (utils.py:198-200) */
} JrrPoseKind;

typedef struct JrrModel JrrModel; /* opaque: packed, padded, tf32 hi/lo-split constants */

/* Host-side description of the body-model constants, i.e. the buffers smplx.SMPL registers
 * plus the two that scripts/smpl.py:67-70 adds.  All pointers are HOST pointers, dense,
 * row-major, fp32 unless noted. */
typedef struct JrrModelDesc {
  const float* v_template_host;        /* [6890,3] */
  const float* shapedirs_host;         /* [6890,3,10] */
  const float* posedirs_host;          /* [207,20670] */
  const float* J_regressor_host;       /* [24,6890] */
  const int64_t* parents_host;         /* [24], parents[0] = -1 */
  const float* lbs_weights_host;       /* [6890,24] (any sparsity; packed as ELL-4 runs when
                                          every vertex has <= 4 non-zeros, which SMPL has) */
  const float* J_regressor_extra_host; /* [9,6890]  (smpl.py:67-69) */
  const int64_t* joint_map_host;       /* [49] into the 54-joint stack (smpl.py:66,70) */
  const int64_t* vertex_picks_host;    /* [21] VertexJointSelector ids (smplx) */
  int32_t device;                      /* CUDA device ordinal the model lives on */
  int32_t gemm_impl;                   /* 0 = tcgen05 3xTF32 (product), 1 = SIMT fp32 (kernel validation only) */
} JrrModelDesc;

const char* jrr_last_error(void);
int jrr_abi_version(void);

/* replaces: SMPL.__init__ (scripts/smpl.py:64-70, smplx.SMPL.__init__) */
int jrr_model_create(const JrrModelDesc* desc, JrrModel** out);
int jrr_model_destroy(JrrModel* model);

/* replaces: the per-call "J*mask, ReLU, row-normalise" of find_joints (scripts/utils.py:87-92);
 * done once per regressor version.  J17_raw / mask are DEVICE [17,6890] row-major, mask may
 * be NULL (== all ones, which is what utils.py:182-187 returns). */
int jrr_set_regressor(JrrModel* model, const float* J17_raw, const float* mask, void* stream);

/* Loss-path formulation used by jrr_refine_step / jrr_refine_step_2d (the results are the same function
 * of the inputs; see DESIGN.md "folded loss path"):
 *   JRR_LOSS_PATH_VERTEX  per-vertex: blend GEMM [B,224]x[224,20736] with skinning and the 17x6890
 *                         regressor reduction fused into it, and the mirrored backward (default)
 *   JRR_LOSS_PATH_FOLDED  the constant linear maps between blend features and regressed joints
 *                         (regressor o skinning weights o blend matrix) are folded once per regressor
 *                         version into T[24*17*3, 224]; a step then runs two N = 1224 GEMMs and a per-frame
 *                         contraction with the joint transforms instead of any per-vertex work.
 * Re-folds inside jrr_set_regressor / jrr_regressor_apply while selected. */
#define JRR_LOSS_PATH_VERTEX 0
#define JRR_LOSS_PATH_FOLDED 1
int jrr_set_loss_path(JrrModel* model, int mode, void* stream);

/* replaces: Discriminator.__init__/load_state_dict (scripts/discriminator.py:7-30).
 * `params` is DEVICE fp32, the state_dict tensors flattened and concatenated in this order:
 * conv_operations.0.weight[32,6] .bias[32] conv_operations.2.weight[32,32] .bias[32]
 * linears.0..23 (weight[32], bias[1]) x24, linear_operations.0.weight[1024,768] .bias[1024]
 * linear_operations.2.weight[1024,1024] .bias[1024] linear_operations.4.weight[1024] .bias[1]
 * (JRR_CRITIC_PARAMS floats). */
int jrr_critic_load(JrrModel* model, const float* params, void* stream);

/* replaces: Shape_Discriminator.__init__/load_state_dict (scripts/discriminator.py:57-74) and the
 * weight of its loss term (optimize.py:244,249-250,253: 10).  `params` is DEVICE fp32, the
 * state_dict flattened in order: shape_operations.0.weight[10,10] .bias[10]
 * shape_operations.2.weight[5,10] .bias[5] shape_operations.4.weight[1,5] .bias[1]
 * (JRR_SHAPE_CRITIC_PARAMS floats).  From then on every jrr_refine_step / jrr_refine_step_2d adds
 * w_shape * mean((sigmoid(shape_critic(betas)) - 1)^2) to the loss and its gradient to the betas.
 * params == NULL switches the term off again. */
int jrr_shape_critic_load(JrrModel* model, const float* params, float w_shape, void* stream);

/* replaces: Shape_Discriminator.forward (scripts/discriminator.py:70-74): betas [B,10] ->
 * sigmoid scores [B]. */
int jrr_shape_critic_forward(JrrModel* model, int64_t B, const float* betas, float* scores_out, void* stream);

/* bytes of caller-owned device workspace needed by any entry point below for B poses */
size_t jrr_workspace_bytes(const JrrModel* model, int64_t B);

/* replaces: SMPL.forward (scripts/smpl.py:72-85 -> smplx.SMPL.forward -> smplx.lbs.lbs).
 * betas [B,10]; pose per `kind`; vertices_out [B,6890,3] or NULL; joints49_out [B,49,3] or
 * NULL. */
int jrr_smpl_forward(JrrModel* model, int64_t B, const float* betas, const float* pose,
                     int kind, float* vertices_out, float* joints49_out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* replaces: autograd backward of SMPL.forward (optimize.py:264 through smpl.py:72-85).
 * dvertices [B,6890,3] or NULL, djoints49 [B,49,3] or NULL -> dbetas_out [B,10],
 * dpose_out in the layout of `kind`.  Recomputes the forward intermediates. */
int jrr_smpl_backward(JrrModel* model, int64_t B, const float* betas, const float* pose,
                      int kind, const float* dvertices, const float* djoints49,
                      float* dbetas_out, float* dpose_out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* replaces: utils.find_joints (scripts/utils.py:85-103) on the loss path: regressed
 * 17 joints [B,17,3] WITHOUT materialising vertices (needs jrr_set_regressor). */
int jrr_find_joints(JrrModel* model, int64_t B, const float* betas, const float* pose, int kind,
                    float* joints17_out, void* workspace, size_t workspace_bytes, void* stream);

/* replaces: Discriminator.forward (scripts/discriminator.py:32-54): rot6d [B,24,6] ->
 * sigmoid scores [B,25] ordered [global, joint0..23]. */
int jrr_critic_forward(JrrModel* model, int64_t B, const float* rot6d, float* scores_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* replaces: one iteration of the refinement loop, optimize.py:220-229,238-253,263-265
 * (rot6d_to_rotmat, find_joints, move_pelvis + MSELoss, Discriminator + MSELoss, backward,
 * Adam.step) restricted to the in-scope loss w_joint*joint + w_pose*pose_critic.
 *   x6 [B,24,6], betas [B,10]      updated in place
 *   adam_m / adam_v [B,154]        first/second moments, per pose [x6(144) | betas(10)]
 *   step_count                     DEVICE int32 scalar, incremented by the call (bias correction)
 *   gt_mm [B,17,3]                 pelvis-centred target joints in millimetres
 *   B_logical                      batch size used by the two mean reductions (optimize.py:128),
 *                                  so a shard of a larger batch reproduces its gradients
 *   loss_out                       DEVICE float[JRR_LOSS_TERMS] {total, joint_mse, pose_mse, 2d_mse,
 *                                  shape_mse} sums over this shard already divided by the logical
 *                                  element counts (terms that are off read 0); or NULL */
int jrr_refine_step(JrrModel* model, int64_t B, int64_t B_logical, float* x6, float* betas,
                    const float* gt_mm, float* adam_m, float* adam_v, int32_t* step_count,
                    float lr, float w_joint, float w_pose, float* loss_out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* --- widening row "2-D reprojection loss + camera fit"  ---------------------
 * replaces: return_2d_joints (scripts/renderer.py:10-51: flip x/y, x2, pytorch3d 0.3.0
 * PerspectiveCameras(T=cam, focal 5000/224, principal point 0).transform_points_screen at
 * 224x224) + MSELoss against gt_j2d.
 *
 * jrr_camera_fit: the 1000-iteration camera-only Adam loop of optimize.py:187-199.  The 3-D
 * joints do not depend on the camera, so the body model runs ONCE (not 1000x) and every frame
 * then iterates privately in one kernel.  cam [B,3] in/out; gt_j2d [B,17,2] pixels; loss_out
 * DEVICE float[1] (final mean squared pixel error) or NULL.
 *
 * jrr_refine_step_2d: jrr_refine_step with the loss_j2d term added (optimize.py:231-233,252-253,
 * weight w_2d = 1/100 there) and the camera translation as a fourth Adam parameter group
 * (optimize.py:201-202).  cam_adam_m/v [B,3]; loss_out as for jrr_refine_step. */
int jrr_camera_fit(JrrModel* model, int64_t B, int64_t B_logical, const float* x6, const float* betas,
                   const float* gt_j2d, float* cam, int iters, float lr, float* loss_out, void* workspace,
                   size_t workspace_bytes, void* stream);
int jrr_refine_step_2d(JrrModel* model, int64_t B, int64_t B_logical, float* x6, float* betas,
                       const float* gt_mm, const float* gt_j2d, float* cam, float* adam_m, float* adam_v,
                       float* cam_adam_m, float* cam_adam_v, int32_t* step_count, float lr, float w_joint,
                       float w_pose, float w_2d, float* loss_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* --- widening row "critic training step"  -----------------------------------
 * replaces: optimize.py:276-293 -- after a batch has been refined both discriminators take one
 * Adam step on MSE(D(refined), 0) + MSE(D(initial), 1) (disc_optimizer / shape_disc_optimizer,
 * optimize.py:113-123).
 *
 * jrr_critic_grad_accumulate: G_accum[JRR_CRITIC_PARAMS] (DEVICE, jrr_critic_load's flat order)
 * += d/dparams of sum_b sum_25 (D(x6)_b - target)^2 / (25 * B_logical) under the CURRENT critic
 * weights; loss_accum[1] += that loss (or NULL).  Call it with target 0 on the refined poses and
 * target 1 on the initial ones; per chunk (B <= 16384) and per rank, all-reduce G_accum
 * (7.36 MB) when frames are sharded.  x6 [B,24,6].
 *
 * jrr_critic_apply: torch.optim.Adam step (defaults) on the caller's flat `params` (DEVICE,
 * in/out; adam_m / adam_v same length; step_count DEVICE int32, incremented), then the model's
 * packed copies are refreshed from `params` as by jrr_critic_load.
 *
 * The jrr_shape_critic_* pair is the same for Shape_Discriminator (betas [B,10], 171 params,
 * one score per frame). */
int jrr_critic_grad_accumulate(JrrModel* model, int64_t B, int64_t B_logical, const float* x6, float target,
                               float* G_accum, float* loss_accum, void* workspace, size_t workspace_bytes,
                               void* stream);
int jrr_critic_apply(JrrModel* model, float* params, const float* G, float* adam_m, float* adam_v,
                     int32_t* step_count, float lr, void* stream);
int jrr_shape_critic_grad_accumulate(JrrModel* model, int64_t B, int64_t B_logical, const float* betas, float target,
                                     float* G_accum, float* loss_accum, void* workspace, size_t workspace_bytes,
                                     void* stream);
int jrr_shape_critic_apply(JrrModel* model, float* params, const float* G, float* adam_m, float* adam_v,
                           int32_t* step_count, float lr, void* stream);

/* --- widening row "evaluation"  ----------------------------------------------
 * replaces: utils.evaluate (scripts/utils.py:117-145) + batch_compute_similarity_transform_torch
 * (scripts/eval_utils.py:7-58): mean MPJPE and Procrustes-aligned MPJPE in millimetres.
 * pred_j3d [B,17,3] metres, target_j3d_mm [B,17,3] millimetres (both are pelvis-centred inside);
 * out_mm DEVICE float[2]; per_frame_mm DEVICE [B,2] or NULL; scratch: 16 bytes per 128 frames. */
int jrr_evaluate(int64_t B, const float* pred_j3d, const float* target_j3d_mm, float* out_mm,
                 float* per_frame_mm, void* scratch, size_t scratch_bytes, void* stream);

/* replaces: the autograd backward of utils.find_joints (scripts/utils.py:85-103) w.r.t. the body-model inputs, as reached
 * when a caller differentiates a loss on the regressed joints itself (renderer.py:27-28 -> optimize.py:193-199,231-233):
 * djoints17 [B,17,3] -> dbetas_out [B,10], dpose_out (layout of `pose`).  The regressor is the one jrr_set_regressor holds
 * (its own gradient is the refit's: jrr_regressor_grad_accumulate).  Recomputes the forward. */
int jrr_find_joints_backward(JrrModel* model, int64_t B, const float* betas, const float* pose, int kind,
                             const float* djoints17, float* dbetas_out, float* dpose_out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Gradients of loss terms computed OUTSIDE jrr_refine_step / jrr_refine_step_2d -- the silhouette term of
 * optimize.py:234-236,252-253, i.e. jrr_silhouette_backward chained through jrr_smpl_backward(kind = ROT6D) -- are added to
 * the step's own parameter gradients before the Adam update (optimize.py:252-253,263-265: one optimiser step on the sum of
 * the terms).  dx6 [B,24,6], dbetas [B,10], dcam [B,3] (used by jrr_refine_step_2d only); any may be NULL; the pointers are
 * read by every later step until cleared with three NULLs.  8-byte aligned. */
int jrr_set_external_gradient(JrrModel* model, const float* dx6, const float* dbetas, const float* dcam);

/* --- widening row "silhouette term"  ------------------------------------------
 * replaces: Mesh_Renderer.forward (scripts/mesh_renderer.py:23-79: pytorch3d 0.3.0 MeshRasterizer(blur_radius=0,
 * faces_per_pixel=1) + SoftSilhouetteShader(sigma=1e-4) behind PerspectiveCameras(T=cam, focal_length=5000/image_size)),
 * render_mesh (scripts/optimize.py:77-85: x, y flipped and the vertices scaled by 2 inside) and the MSELoss against the
 * Mask R-CNN silhouette (optimize.py:234-236), with its backward to the vertices and the camera translation.
 * vertices [B,V,3]: with flip_scale != 0 the body model's output (render_mesh's x, y flip and scale by 2 happen inside),
 * with flip_scale == 0 a ready mesh (what Mesh_Renderer.forward receives); cam [B,3]; faces [F,3] int32;
 * alpha_out [B,S,S] (channel 3 of the reference's [B,4,S,S] image; row 0 = top), pix_to_face_out [B,S,S] (-1 = background).
 * target [B,S,S] + loss_out DEVICE float[1] (optional, together): mean((alpha - target)^2) over B_logical*S*S elements.
 * The backward takes either d loss / d alpha (dalpha) or, with dalpha == NULL, the MSE target (gradient of
 * loss_weight * that mean); it reads the projected vertices the forward left in the workspace, so it must follow the
 * forward of the same inputs on the same workspace (one backward per forward: it reuses that buffer).  vert_face_ptr [V+1] / vert_face_idx: CSR vertex -> (face*3 + corner).
 * dcam_out [B,3] may be NULL.  All pointers DEVICE. */
size_t jrr_silhouette_workspace_bytes(int64_t B, int64_t V, int64_t F, int image_size);
int jrr_silhouette_forward(int64_t B, const float* vertices, int64_t V, const float* cam, const int32_t* faces, int64_t F,
                           int image_size, float focal, float sigma, int flip_scale, const float* target, int64_t B_logical,
                           float* alpha_out, int32_t* pix_to_face_out, float* loss_out, void* workspace,
                           size_t workspace_bytes, void* stream);
int jrr_silhouette_backward(int64_t B, const float* vertices, int64_t V, const float* cam, const int32_t* faces, int64_t F,
                            const int32_t* vert_face_ptr, const int32_t* vert_face_idx, int image_size, float focal,
                            float sigma, int flip_scale, const float* alpha, const int32_t* pix_to_face, const float* dalpha,
                            const float* target, int64_t B_logical, float loss_weight, float* dvertices_out,
                            float* dcam_out, void* workspace, size_t workspace_bytes, void* stream);

/* replaces: the forward/backward half of the regressor refit, optimize.py:300-309
 * (find_joints on detached refined poses, move_pelvis + MSELoss, backward to J_regressor).
 * Accumulates G += dL/dJhat (17x6890, gradient w.r.t. the NORMALISED regressor) and
 * loss_accum += shard loss; both are what a multi-GPU caller all-reduces. */
int jrr_regressor_grad_accumulate(JrrModel* model, int64_t B, int64_t B_logical, const float* x6,
                                  const float* betas, const float* gt_mm, float* G_accum,
                                  float* loss_accum, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* replaces: backward through row-normalise/ReLU/mask (utils.py:87-92) followed by
 * J_Regressor_optimizer.step() (optimize.py:125-126,310-312): Adam on the raw regressor
 * with persistent state.  J17_raw, adam_m, adam_v [17,6890] are updated in place; the
 * model's normalised copy is refreshed.  step_count as in jrr_refine_step. */
int jrr_regressor_apply(JrrModel* model, float* J17_raw, const float* mask, const float* G_accum,
                        float* adam_m, float* adam_v, int32_t* step_count, float lr,
                        void* stream);

/* Diagnostic (SYNCHRONISES the stream, not graph-capturable): jrr_refine_step with a CUDA
 * event recorded on `stream` between its kernel groups; ms_out_host[JRR_STEP_KERNELS] (HOST)
 * receives each group's device time in milliseconds, names from jrr_step_kernel_name().
 * bench.py uses it for the per-kernel roofline. */
#define JRR_STEP_KERNELS 14
int jrr_refine_step_profiled(JrrModel* model, int64_t B, int64_t B_logical, float* x6, float* betas,
                             const float* gt_mm, float* adam_m, float* adam_v, int32_t* step_count,
                             float lr, float w_joint, float w_pose, float* loss_out, void* workspace,
                             size_t workspace_bytes, void* stream, float* ms_out_host);
const char* jrr_step_kernel_name(int i);

/* Diagnostic: the 3xTF32 GEMM on its own, C[M,N] (row-major) = A[M,K] . B[N,K]^T with fp32
 * inputs that are split into tf32 hi/lo pairs inside the call (scratch = 2*(M+N)*K floats,
 * device).  impl 0 = tcgen05 kernel, 1 = SIMT validation kernel, 2 = tcgen05 kernel that takes the
 * fp32 A as it is and stages its tf32 hi/lo pair through tensor memory (N%128 == 0).  M%128 == 0, K%32 == 0,
 * N%128 == 0 or N == 224.  Used by the kernel-level parity tests and the GEMM roofline
 * micro-benchmark; not on the reference's interface. */
int jrr_debug_gemm(JrrModel* model, int impl, int64_t M, int64_t N, int64_t K, const float* A,
                   const float* B, float* C, float* scratch, void* stream);

/* Diagnostic (benchmarks/tma_probe.py): raw TMA tile-load rate.  `grid` CTAs each stream `iters` stages of `boxes` 2-D
 * boxes ([box_rows] x 32 fp32, SWIZZLE_128B) of the DEVICE matrix src[rows][cols] through a ring of `stages` stages; a consumer
 * lane frees every stage as soon as it has landed (after `dwell_ns`).  shared_tiles != 0: all CTAs read the same tiles.
 * Not on the reference's interface. */
int jrr_debug_tma_probe(const float* src, int64_t rows, int64_t cols, int box_rows, int boxes, int stages, int shared_tiles,
                        int iters, int grid, int dwell_ns, void* stream);

/* Diagnostic (benchmarks/gemm_prof.py): role timers (cycles each warp role of the critic GEMM kernel spends waiting) of the
 * next launches are written to consecutive [148][16] int64 slots of the DEVICE buffer `base`; base == NULL switches it off. */
int jrr_debug_set_gemm_prof(long long* base, int slots);

/* Measurement aid (bench.py's FLOP accounting; no reference counterpart): tensor-core products issued per K step by the
 * backward GEMM of the critic's second wide layer in a refinement step over B frames -- 2 when it takes the ReLU mask as a
 * 0/1 operand against diag(w3) W2 (exact in tf32, see csrc/jrr_critic.cu), 3 for the generic 3xTF32 scheme. */
int jrr_critic_layer2_bwd_products(const JrrModel* model, int64_t B);

/* number of kernels the last call of the named entry point enqueued (bench.py's
 * gpu_launches claim is counted, not guessed) */
int64_t jrr_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JRR_H_ */
