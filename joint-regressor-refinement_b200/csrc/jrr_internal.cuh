// Internal declarations shared by the libjrr.so translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jrr.h"

namespace jrr {

// ---- canonical sizes -------------------------------------------------------------------
constexpr int V = JRR_NUM_VERTS;    // 6890
constexpr int VP = 6912;            // vertices padded to 54*128
constexpr int NP = 3 * VP;          // 20736 blend columns (3 per packed vertex)
constexpr int NJ = 24;
constexpr int NB = 10;
constexpr int NF = 207;             // pose-feature length
constexpr int KA = 224;             // augmented K: 207 pose + 10 beta + 1 template, padded to 7*32
constexpr int FEAT_BETA = 207;
constexpr int FEAT_ONE = 217;
constexpr int NH = 17;              // regressed joints
constexpr int NACC = 51;            // 17*3
constexpr int JH_STRIDE = 20;       // padded Jhat column inside a vertex record
constexpr int VS_F = 576;           // packed vertices per CTA range, forward skinning
constexpr int NSPLIT = VP / VS_F;   // 12 regressor partial sums per pose
constexpr int VS_B = 768;           // packed vertices per CTA range, backward skinning
constexpr int NSPLIT_B = VP / VS_B; // 9
constexpr int VS_S = 192;           // ... of the module backward on small batches (more CTAs per 256-pose block)
constexpr int NSPLIT_S = VP / VS_S; // 36
constexpr int KSPLIT_MAX = 36;      // split-K of the backward blend GEMM: 6, 9, 12 or 18 (chosen per batch to fill whole waves)
constexpr int NPARAM = 154;         // 144 rot6d + 10 betas per pose
constexpr int MAXCH = 4;            // children per joint supported by the chain kernels

constexpr int C_H = 768;            // critic: 24*32 conv features
constexpr int C_Z = 1024;
// offsets into JrrModel::critic_small (floats)
constexpr int CS_C1W = 0;       // [32][6]
constexpr int CS_C1B = 192;     // [32]
constexpr int CS_C2W = 224;     // [32][32]
constexpr int CS_C2B = 1248;    // [32]
constexpr int CS_HW = 1280;     // [24][32]
constexpr int CS_HB = 2048;     // [24]
constexpr int CS_B1 = 2072;     // [1024]
constexpr int CS_B2 = 3096;     // [1024]
constexpr int CS_W3 = 4120;     // [1024]
constexpr int CS_B3 = 5144;     // [1]
constexpr int CS_TOTAL = 5145;


inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- error handling ----------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);
void count_launch(int n = 1);
void reset_launch_count();

#define JRR_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::jrr::fail(JRR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define JRR_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::jrr::fail(JRR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    ::jrr::count_launch();                                                                     \
  } while (0)

// ---- small tables passed to the chain kernels by value ------------------------------------
struct ChainTab {
  int parent[NJ];
  int depth[NJ];
  int child[NJ][MAXCH];  // -1 when absent
  int max_depth;
  int maxch[NJ];         // by depth d: the largest number of depth-d children any joint has (gather rounds)
};

// per packed vertex: skinning run record (ELL-4 with register-cached joint slots)
//   meta bits [5k,5k+5)  joint id of slot k (k=0..3)
//   meta bits [20+k]     slot k must be (re)loaded before this vertex
//   meta bit  24         this vertex has a non-zero column in the normalised regressor
//   meta bit  25         first vertex of a range (slots are loaded without a flush)
//   xptr/xcnt           slice of (source, coef) pairs: the joints49 sources (21 vertex picks,
//                       9 extra-regressor rows) this vertex feeds (module backward)
//   jh[0..16]           this vertex's column of the normalised 17x6890 regressor
struct VtxRec {
  uint32_t meta;
  float w[4];
  int xptr;
  int xcnt;
  int pad;
  float jh[JH_STRIDE];
};
static_assert(sizeof(VtxRec) == 112, "VtxRec tiles are staged as float4");
constexpr int REC_WORDS = 28;
constexpr int FOLD_N = 24 * 17 * 3;      // rows of the folded operator T: (joint j, regressor row i, coordinate c)
constexpr int FOLD_NP = 1280;            // ... padded to the GEMM tile
constexpr int LOSS_PART_POSE = 16384; // offset of the critic partials inside Workspace::loss_part
constexpr int LOSS_PART_2D = 8192;    // offset of the 2-D reprojection partials (<= 8192 pose blocks of 32 per call)
constexpr int64_t MAX_POSES_PER_CALL = 262144;

// sparse rows (CSR over ORIGINAL vertex ids) for the joints49 path
struct Csr {
  int* ptr = nullptr;
  int* col = nullptr;
  float* val = nullptr;
  int rows = 0;
};

}  // namespace jrr

namespace jrr {
// Record tables of one skinning PASS.  A vertex with more than four skinning weights (SMPL has at most four; the C ABI
// takes any dense [6890,24]) is handled by running the same kernels again with the vertex's next four weights: every
// output of the skinning kernels is linear in the weights, so passes simply add up (regressor partial sums, stored
// vertices, blend-gradient partials, joint-transform gradients, the folded operator).  Pass 0 lives in JrrModel's own fields.
struct PassTab {
  VtxRec *vrec = nullptr, *vrec_b = nullptr, *vrec_l = nullptr;
  int *flush_ptr = nullptr, *flush_idx = nullptr, *range_flush_base = nullptr;
  int n_flush = 0;
  int *flush_ptr_l = nullptr, *flush_idx_l = nullptr, *range_flush_base_l = nullptr;
  int n_flush_l = 0;
  // module backward on small batches: every vertex in 192-vertex ranges (36 K ranges per 256-pose block instead of 9)
  VtxRec* vrec_s = nullptr;
  int *flush_ptr_s = nullptr, *flush_idx_s = nullptr, *range_flush_base_s = nullptr;
  int n_flush_s = 0;
};
constexpr int MAX_PASS = 6;      // 24 weights per vertex
}  // namespace jrr

struct JrrModel {
  int device = 0;
  int gemm_impl = 0;
  int num_sms = 148;
  jrr::ChainTab chain;
  // augmented blend matrix, tf32 hi/lo split, both majors
  float *Pt_hi = nullptr, *Pt_lo = nullptr;  // [NP][KA]  (K contiguous)  forward B operand
  float *P_hi = nullptr, *P_lo = nullptr;    // [KA][NP]  (N contiguous)  backward B operand
  float *P3_hi = nullptr, *P3_lo = nullptr;  // [3][KA][VP] (vertex contiguous, per coordinate): B operand of the fold GEMM
  float* Pn = nullptr;                       // [KA][3*6890] natural-order fp32 master the packings are gathered from
  float* J0 = nullptr;                       // [24][3]      J_regressor . v_template
  float* JS = nullptr;                       // [24][3][10]  J_regressor . shapedirs
  jrr::VtxRec* vrec = nullptr;               // [VP]  forward ranges (VS_F)
  jrr::VtxRec* vrec_b = nullptr;             // [VP]  backward ranges (VS_B): reload/first flags differ
  int* perm = nullptr;                       // [VP] packed index -> original vertex id (-1 = padding)
  int* inv_perm = nullptr;                   // [V]  original vertex id -> packed index
  int* vx_src = nullptr;                     // joints49 sources per vertex (see VtxRec)
  float* vx_coef = nullptr;
  int n_flush = 0;                           // dA flush events per pose (all ranges)
  int n_flush_act = 0;                       // ... of the active ranges only (a prefix of the ids)
  int nv_act = 0;                            // packed vertices the loss path walks (VP when the regressor is dense)
  int vs_l = jrr::VS_B;                      // vertices per K range of the fused backward: 768, 384 or 192
  int nsplit_act = 0;                        // nv_act / vs_l (<= 36)
  jrr::VtxRec* vrec_l = nullptr;             // [VP] records with vs_l range boundaries (fused backward)
  int n_flush_l = 0;                         // flush events of the fused backward (active ranges only)
  int* flush_ptr_l = nullptr;                // [25]
  int* flush_idx_l = nullptr;
  int* range_flush_base_l = nullptr;         // [37]
  jrr::VtxRec* vrec_s = nullptr;             // [VP] records with 192-vertex range boundaries (module backward, small batches)
  int *flush_ptr_s = nullptr, *flush_idx_s = nullptr, *range_flush_base_s = nullptr;
  int n_flush_s = 0;
  bool compact_active = true;                // pack vertices with a non-zero regressor column first
  uint8_t* active_dev = nullptr;             // [V] scratch of jrr_set_regressor
  std::vector<uint8_t> packed_active;        // support the current packing was built for
  std::vector<uint32_t> h_key;               // host copies used by build_packing
  std::vector<std::vector<std::pair<int, float>>> h_lbs, h_vx;
  int* flush_ptr = nullptr;                  // [25] CSR joint -> flush ids
  int* flush_idx = nullptr;                  // [n_flush]
  int* range_flush_base = nullptr;           // [NSPLIT_B]
  // joints49 path
  jrr::Csr extra;                            // [9] rows over original vertex ids
  int* picks = nullptr;                      // [21]
  int* joint_map = nullptr;                  // [49]
  // gradients of loss terms computed OUTSIDE the refinement step (the silhouette term: rasteriser + module backward), added
  // to the step's own parameter gradients before Adam (jrr_set_external_gradient; NULL = none)
  const float *ext_dx6 = nullptr, *ext_dbetas = nullptr, *ext_dcam = nullptr;
  unsigned* small_counter = nullptr;         // block counter of the single-launch small-batch forward (self-resetting)
  // regressor (normalised), refreshed by jrr_set_regressor / jrr_regressor_apply
  float* Jhat = nullptr;                     // [17][V]   original vertex order
  float* rowsum = nullptr;                   // [17]
  float* regdot = nullptr;                   // [17] scratch of jrr_regressor_apply
  bool has_regressor = false;
  // critic
  float* critic_small = nullptr;             // conv + heads + w3/b3 + biases (see jrr_critic.cu)
  float *W1_hi = nullptr, *W1_lo = nullptr;    // [1024][768]
  float *W2_hi = nullptr, *W2_lo = nullptr;    // [1024][1024]
  float *W1t_hi = nullptr, *W1t_lo = nullptr;  // [768][1024]
  float *W2t_hi = nullptr, *W2t_lo = nullptr;  // [1024][1024]
  float *W2tw_hi = nullptr, *W2tw_lo = nullptr;  // [1024][1024]  (diag(w3) W2)^T: the backward of layer 2 with a 0/1 A operand
  bool has_critic = false;
  float* shape_critic = nullptr;             // 171 floats: shape_operations.{0,2,4} weight/bias (discriminator.py:57-74)
  bool has_shape_critic = false;
  float w_shape = 0.f;                       // weight of the shape-critic term in the refinement loss (optimize.py:253: 10)
  // the critic chain of a refinement step runs on a forked side stream (joins before Adam)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_seed = nullptr, ev_join2 = nullptr;
  bool overlap_critic = true;
  bool split_adam = true;                    // chain backward beside the critic branch, element-wise Adam after the join
  bool fold_ts = true;                       // folded loss path of the refine step: plain fp32 features / dQ, CTA-pair GEMMs
  bool critic_ts = true;                     // critic GEMMs of the refine step: plain fp32 activations staged through tensor
                                             // memory (JRR_CRITIC_TS=0: pre-split activations through shared memory)
  bool critic_headless = true;               // no head kernel: dL/dlogit inside the backward GEMM, joint heads in critic_post
  bool critic_head_fused = true;             // global critic head inside the layer-2 GEMM epilogue (refine step)
  // folded loss path (jrr_set_loss_path): T = Jhat o skinning weights o blend matrix, per regressor version
  bool folded = false;
  float *T_hi = nullptr, *T_lo = nullptr;    // [FOLD_NP][KA]   forward B operand
  float *Tt_hi = nullptr, *Tt_lo = nullptr;  // [KA][FOLD_NP]   backward B operand
  float* Tc = nullptr;                       // [24][17]        sum_v Jhat_iv w_vj
  double* fold_part = nullptr;               // scratch of fold_kernel
  float *fg_wj = nullptr, *fg_wj_hi = nullptr, *fg_wj_lo = nullptr;   // [512][VP] w * Jhat per (joint, regressor row): A operand of the fold GEMM
  float* fg_part = nullptr;                  // [3][36][512][224] K-split partials of the fold GEMM
  double* fold_wj = nullptr;                 // [17][VP][4] w * Jhat in double (fold_prep_kernel)
  double* fold_acc = nullptr;                // [1224*224 + 24*17] the folded operator in double (fold_gather_kernel)
  double* fold_ev = nullptr;                 // [fold_ev_cap][17][4][224] run sums per flush event (fold_runs_kernel)
  int fold_ev_cap = 0;
  bool fused_fwd = true;   // loss path: skinning + regressor in the blend GEMM's epilogue
  bool fused_bwd = true;   // loss path: skinning backward generates the A operand of the blend-gradient GEMM
  // skinning passes (see PassTab): n_pass = ceil(max non-zeros per lbs_weights row / 4)
  int n_pass = 1, cur_pass = 0;
  jrr::PassTab passes[jrr::MAX_PASS];        // [0] mirrors the fields above
  int flush_off[jrr::MAX_PASS + 1] = {0};    // offset (in flush events) of each pass inside Workspace::dAflush
  void select_pass(int p) {                  // point the table fields at pass p (launches capture them by value)
    const jrr::PassTab& t = passes[p];
    vrec = t.vrec; vrec_b = t.vrec_b; vrec_l = t.vrec_l;
    flush_ptr = t.flush_ptr; flush_idx = t.flush_idx; range_flush_base = t.range_flush_base; n_flush = t.n_flush;
    flush_ptr_l = t.flush_ptr_l; flush_idx_l = t.flush_idx_l; range_flush_base_l = t.range_flush_base_l; n_flush_l = t.n_flush_l;
    vrec_s = t.vrec_s; flush_ptr_s = t.flush_ptr_s; flush_idx_s = t.flush_idx_s; range_flush_base_s = t.range_flush_base_s; n_flush_s = t.n_flush_s;
    n_flush_act = t.n_flush_l;
    cur_pass = p;
  }
  std::vector<void*> allocs;
};

namespace jrr {

// ---- workspace ---------------------------------------------------------------------------
struct Workspace {
  int64_t B, BP;
  float* AT;        // [288][BP]   relative transforms, pose contiguous
  float* feat_hi;   // [BP][224]
  float* feat_lo;
  float* vpT;       // [NP][BP]    posed-blend vertices, pose contiguous
  float* part;      // [n_pass][slots][51][BP] regressor partial sums (one region per skinning pass)
  int64_t part_stride;   // floats per pass region
  float* gT;        // [51][BP]    loss seed (pelvis adjusted)
  float* pred;      // [BP][51]
  float* dvp_hi;    // [BP][NP]
  float* dvp_lo;
  float* dAflush;   // [n_flush][12][BP]
  float* dAT;       // [288][BP]
  float* dfeat;     // [ksplit][BP][224]
  int ksplit;       // split-K factor of the backward blend GEMM for this batch
  int n_joint_part; // number of joint-loss partials the last seed kernel wrote
  float* dJp;       // [BP][72]    grad wrt posed joints (module path)
  float* Jp;        // [BP][72]    posed joints
  float* d30T;      // [90][BP]    joints49 gradient gathered onto picks / extra rows
  float* loss_part; // [2][4096]   per-CTA loss partials: joint term, critic term
  int n_pose_part;  // CTAs that wrote critic partials in this step
  // critic
  float *h_hi, *h_lo;      // [BP][768]
  float *z1_hi, *z1_lo;    // [BP][1024]
  float *z2_hi, *z2_lo;    // [BP][1024]
  float *dz2_hi, *dz2_lo;  // [BP][1024]
  float *dz1_hi, *dz1_lo;  // [BP][1024]
  float* dh;               // [BP][768]
  float* dzj;              // [BP][24]   grads wrt the 24 joint-head logits
  float* dx6c;             // [BP][144]  critic grad wrt rot6d
  float* dbeta_s;          // [BP][10]   shape-critic grad wrt betas
  float* shape_part;       // [BP/128]   shape-critic loss partials
  float* zj;               // [BP][24]   joint-head logits (head-fused critic path)
  float* zg_part;          // [8][BP]    global-head logit partials per 128-column tile of layer 2
  float* dzg;              // [BP]       dL/d(global logit)
  uint2* cmask;            // [BP][24]   ReLU masks of the critic's two 1x1 convs (pre -> post)
  uint32_t* zmask;         // [BP][32]   ReLU mask bits of the critic's first wide layer
  uint32_t* zmask2;        // [BP][32]   ... of the second (the A operand of the layer-2 backward GEMM, bit-packed)
  float* gx6;              // [BP][144]  parameter gradients of the chain backward (split-Adam schedule)
  float* gbetas;           // [BP][10]
  float* adam_coef;        // [2]        this step's Adam bias corrections (adam_coef_kernel)
  float* scores;           // [BP][25]
  size_t bytes;
};
Workspace carve(const JrrModel* m, int64_t B, void* base);

// ---- kernels (host launch wrappers; all asynchronous on `st`) --------------------------------
int launch_pose_fwd(const JrrModel* m, int64_t B, int64_t BP, const float* betas, const float* pose,
                    int kind, float* AT, float* feat_hi, float* feat_lo, float* Jp, cudaStream_t st);

// C[m][n] = sum_k A[m][k] * B[n][k], operands given as tf32 hi/lo pairs
enum GemmEpi { EPI_STORE_T = 0, EPI_BIAS_RELU_SPLIT = 1, EPI_MASK_SPLIT = 2, EPI_STORE_SPLITK = 3, EPI_BIAS_RELU_HEAD = 4 };
struct GemmDesc {
  const float *A_hi, *A_lo; int64_t lda;
  const float *B_hi, *B_lo; int64_t ldb;
  int64_t M, N, K;          // K here is the per-split K extent
  int ksplit;               // number of K splits (EPI_STORE_SPLITK), else 1
  int epi;
  float* out0; float* out1; int64_t ldo;   // out (or out_hi/out_lo)
  const float* bias;                       // [N] (EPI_BIAS_RELU_SPLIT)
  const float* mask; int64_t ldmask;       // (EPI_MASK_SPLIT) multiply by (mask>0)
  const float* rowscale;                   // (EPI_MASK_SPLIT) and by rowscale[m] when given
  const float* vec; float* out2;           // (EPI_BIAS_RELU_HEAD) w3[N] in, logit partials [N/128][M] out
  const uint32_t* mask_bits;               // (EPI_MASK_SPLIT) ReLU mask as bits [M][N/32] instead of `mask`
  uint32_t* mask_bits_out;                 // (EPI_BIAS_RELU_SPLIT) also emit the ReLU mask as bits [M][N/32]
  const float* logit_part; int n_logit_part; const float* logit_bias; float logit_gscale; int64_t rows_valid;
                                           // (EPI_MASK_SPLIT) row scale = gscale * (s - 1) s (1 - s), s = sigmoid(bias +
                                           // sum of the row's logit partials), zero for rows >= rows_valid
  const uint32_t* a_bits;                  // (CTA-pair kernel) A is a 0/1 matrix given as bits [M][K/32] (bit k%32 of word k/32): exact
                                           // in tf32, so the A_lo . B_hi product is dropped -- two MMAs per K step instead of three,
                                           // no A tile through TMA
  int64_t k_valid;                         // > 0: only the first k_valid columns of A exist (TMA zero-fills the rest of K)
  bool probe_env;                          // jrr_debug_gemm: honour the JRR_GEMM_PROBE* diagnostic environment knobs
  bool a_via_tmem;                         // A is plain fp32 (A_hi): loaded, tf32-split and staged in tensor memory by the
                                           // kernel (B stays pre-split); *_SPLIT epilogues then write ONE fp32 array (out0)
};
int launch_gemm(const JrrModel* m, const GemmDesc& g, cudaStream_t st);
int launch_gemm_simt(const GemmDesc& g, cudaStream_t st);
int launch_gemm_tc(const JrrModel* m, const GemmDesc& g, cudaStream_t st);
bool gemm_pair_bits_available(const JrrModel* m, int64_t M);   // the 0/1-A variant of the CTA-pair kernel can run

int launch_skin_fwd(const JrrModel* m, const Workspace& w, float* vertices_out, float* vT_out,
                    bool want_part, cudaStream_t st);
// optional 2-D reprojection term of the refinement loss (optimize.py:231-233 via renderer.py:10-51);
// the camera translation is a per-pose parameter of the same Adam optimiser (optimize.py:201-202)
struct Proj2D {
  const float* gt2d = nullptr;   // [B,17,2] screen pixels; nullptr disables the term
  float* cam = nullptr;          // [B,3] camera translation, updated in place
  float* cam_m = nullptr;        // [B,3] Adam moments
  float* cam_v = nullptr;
  const int32_t* step_count = nullptr;
  float lr = 0.f;
  float scale = 0.f;             // w_2d * 2 / (34 * B_logical)
  const float* dcam_ext = nullptr;   // [B,3] camera gradient of an external loss term, added before the camera's Adam step
};
int launch_folded_seed(const JrrModel* m, Workspace& w, const float* gt_mm, int64_t B_logical, float w_joint,
                       float* joints17_out, const Proj2D& p2d, cudaStream_t st, float* dc_part = nullptr,
                       bool plain_dq = false);
// regressor refit through the folded operator (jrr_model.cu): G += unfold(dT, dc)
int regressor_accumulate_folded(const JrrModel* m, Workspace& w, const float* gt_mm, int64_t B_logical, float* G_accum,
                                cudaStream_t st);
int launch_transpose(const float* src, int64_t rows, int cols, float* dst, cudaStream_t st);
int launch_fold(JrrModel* m, cudaStream_t st);
int launch_adam_coef(const Workspace& w, const int32_t* step_count, float lr, cudaStream_t st);
int launch_adam_params(const Workspace& w, bool use_critic, bool use_shape, float* x6, float* betas, float* adam_m,
                       float* adam_v, int32_t* step_count, float lr, cudaStream_t st, const float* ext_dx6 = nullptr,
                       const float* ext_dbetas = nullptr);
int launch_loss_seed(const JrrModel* m, Workspace& w, bool fused_partials, const float* gt_mm,
                     int64_t B_logical, float w_joint, float* joints17_out, const Proj2D& p2d, cudaStream_t st);
int launch_seed_from_dpred(const Workspace& w, const float* dpred, cudaStream_t st);
int launch_camera_fit(const Workspace& w, const float* joints17, const float* gt2d, float* cam, int iters, float lr,
                      int64_t B_logical, float* loss_out, cudaStream_t st);
// whole module forward of a small batch (<= 32 poses, 4 weights per vertex) in one launch (jrr_pose.cu)
bool smpl_small_fwd_available(const JrrModel* m, int64_t B);
int launch_smpl_small_fwd(const JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                          float* vertices, float* joints49, cudaStream_t st);
// ... and its backward: joint-transform / blend-feature / posed-joint gradients into the workspace for launch_pose_bwd
bool smpl_small_bwd_available(const JrrModel* m, int64_t B);
int launch_smpl_small_bwd(const JrrModel* m, Workspace& w, const float* betas, const float* pose, int kind,
                          const float* dverts, const float* dj49, cudaStream_t st);
// fused blend GEMM + skinning + regressor partial sums (jrr_fused_fwd.cu)
struct FusedSched { int n_tiles, T, G, mdiv; };   // see jrr_fused_fwd.cu
FusedSched fused_fwd_sched(const JrrModel* m, int64_t BP, int nv);
int fused_fwd_slots(const JrrModel* m, int64_t BP);
int launch_fused_fwd(const JrrModel* m, const Workspace& w, int store /*0 none, 1 vp, 2 skinned v*/,
                     float* vT_out, cudaStream_t st, bool all_vertices = false);
// packed pose-contiguous vertices vT [3*VP][BP] -> natural order [B][6890][3] (module path)
int launch_unpack_vertices(const JrrModel* m, const Workspace& w, const float* vT, float* vertices_out, cudaStream_t st);
int launch_joints49_fwd_packed(const JrrModel* m, const Workspace& w, const float* vT, float* joints49_out, cudaStream_t st);
// fused skinning backward + transpose-side blend GEMM (jrr_fused_bwd.cu); writes dfeat[NSPLIT_B] and dAflush
int launch_fused_bwd(const JrrModel* m, const Workspace& w, cudaStream_t st, const float* dvT = nullptr);
// natural-order d loss / d vertices [B][6890][3] (may be NULL) + the joints49 gradient on its vertex-borne sources (d30T,
// may be unused) -> packed pose-contiguous dvT [3*VP][BP] for the fused backward
int launch_pack_dvertices(const JrrModel* m, const Workspace& w, const float* dvertices, bool use_x, float* dvT, cudaStream_t st);
int launch_skin_bwd(const JrrModel* m, const Workspace& w, const float* dvertices, bool use_g,
                    bool use_x, cudaStream_t st);
// module backward: 192-vertex K ranges when the batch has too few 256-pose blocks to fill the GPU with 768-vertex ones
inline bool module_small_ranges(const JrrModel* m, int64_t BP) { return m->n_pass == 1 && BP <= 1024; }
int launch_dA_reduce(const JrrModel* m, const Workspace& w, int lists /* 0 module (768), 1 loss path, 2 module (192) */, cudaStream_t st);
int launch_joints49_fwd(const JrrModel* m, const Workspace& w, const float* vertices,
                        float* joints49_out, cudaStream_t st);
int launch_joints49_bwd(const JrrModel* m, const Workspace& w, const float* djoints49,
                        cudaStream_t st);
int launch_pose_bwd(const JrrModel* m, const Workspace& w, const float* betas, const float* pose,
                    int kind, bool use_dJp, bool use_critic, bool use_shape, float* dbetas_out, float* dpose_out,
                    // Adam (refine step) -- all NULL on the module path
                    float* x6, float* betas_rw, float* adam_m, float* adam_v, int32_t* step_count,
                    float lr, cudaStream_t st);

int launch_critic_pre(const JrrModel* m, const Workspace& w, const float* x6, cudaStream_t st, bool want_zj = false);
int launch_critic_head(const JrrModel* m, Workspace& w, int64_t B_logical, float w_pose,
                       float* scores_out, bool want_grad, cudaStream_t st, float target = 1.f, float* dzg = nullptr);
int critic_forward_gemms(const JrrModel* m, const Workspace& w, cudaStream_t st, bool head_fused = false);
int critic_backward_gemms(const JrrModel* m, const Workspace& w, cudaStream_t st, const float* rowscale = nullptr,
                          float headless_gscale = 0.f);
// the loss read-out {total, joint, pose, 2d, shape} from the per-CTA partials (loss_finish_kernel)
struct LossFinishArgs {
  const float* lp_joint; int n_joint; float sj;
  const float* lp_pose; int n_pose; float sp;
  const float* lp_2d; float s2;
  const float* lp_shape; int n_shape; float ss;
  float wj, wp, w2, wsh;
  float* loss_out; float* loss_accum;
};
// one warp; lane-strided partial sums + xor-shuffle tree: a fixed summation order
__device__ __forceinline__ void loss_finish_warp(const LossFinishArgs& a, const int lane) {
  float sa = 0.f, p = 0.f, q = 0.f, r = 0.f;
  for (int i = lane; i < a.n_joint; i += 32) sa += __ldcg(a.lp_joint + i);
  if (a.wsh != 0.f)
    for (int i = lane; i < a.n_shape; i += 32) r += __ldcg(a.lp_shape + i);
  for (int i = lane; i < a.n_pose; i += 32) p += __ldcg(a.lp_pose + i);
  if (a.w2 != 0.f)
    for (int i = lane; i < a.n_joint; i += 32) q += __ldcg(a.lp_2d + i);
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    p += __shfl_xor_sync(0xffffffffu, p, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane != 0) return;
  sa *= a.sj;
  p *= a.sp;
  q *= a.s2;
  r *= a.ss;
  if (a.loss_out != nullptr) {
    a.loss_out[0] = a.wj * sa + a.wp * p + a.w2 * q + a.wsh * r;
    a.loss_out[1] = sa;
    a.loss_out[2] = p;
    a.loss_out[3] = q;
    a.loss_out[4] = r;
  }
  if (a.loss_accum != nullptr) a.loss_accum[0] += sa;
}
LossFinishArgs make_loss_finish_args(const Workspace& w, int64_t B_logical, float w_joint, float w_pose, bool have_pose,
                                     float w_2d, float w_shape, float* loss_out, float* loss_accum);
int launch_critic_head_light(const JrrModel* m, Workspace& w, int64_t B_logical, float w_pose, cudaStream_t st);
int launch_critic_post(const JrrModel* m, Workspace& w, const float* x6, cudaStream_t st, bool with_head = false,
                       int64_t B_logical = 1, float w_pose = 0.f);
int launch_shape_critic(const JrrModel* m, const Workspace& w, const float* betas, int64_t B_logical, cudaStream_t st);
int critic_load_impl(JrrModel* m, const float* params, cudaStream_t st);
int critic_grad_accumulate(JrrModel* m, Workspace& w, int64_t B_logical, const float* x6, float target, float* G_accum,
                           float* loss_accum, cudaStream_t st);
int shape_critic_grad_accumulate(JrrModel* m, Workspace& w, int64_t B_logical, const float* betas, float target,
                                 float* G_accum, float* loss_accum, cudaStream_t st);
int launch_adam_flat(float* p, const float* g, float* am, float* av, int32_t* step_count, float lr, int64_t n,
                     cudaStream_t st);
int launch_shape_critic_scores(const JrrModel* m, int64_t B, const float* betas, float* scores_out, cudaStream_t st);

int launch_loss_finish(const Workspace& w, int64_t B_logical, float w_joint, float w_pose,
                       bool have_pose, float w_2d, float w_shape, float* loss_out, float* loss_accum, cudaStream_t st);

int launch_regressor_normalise(JrrModel* m, const float* Jraw, const float* mask, cudaStream_t st);
int launch_regressor_accumulate(const JrrModel* m, const Workspace& w, const float* vT,
                                float* G_accum, cudaStream_t st);
int launch_regressor_apply(JrrModel* m, float* Jraw, const float* mask, const float* G,
                           float* adam_m, float* adam_v, int32_t* step_count, float lr,
                           cudaStream_t st);

// tf32 helpers (host): round-to-nearest-even into the 19-bit tf32 container
float tf32_round_host(float x);

}  // namespace jrr
