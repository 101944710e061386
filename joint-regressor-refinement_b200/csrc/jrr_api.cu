// C-ABI entry points: argument checks, workspace carving and the kernel sequences.
#include <algorithm>

#include <cstdlib>

#include "jrr_internal.cuh"

namespace jrr {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

Workspace carve(const JrrModel* m, int64_t B, void* base) {
  Workspace w{};
  w.B = B;
  w.BP = round_up(B, 256);   // 128-pose GEMM tiles; the fused backward works on 256-pose blocks
  const size_t BP = (size_t)w.BP;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t nfloats) -> float* {
    float* r = base ? (float*)(p + off) : nullptr;
    off += align256(nfloats * sizeof(float));
    return r;
  };
  w.AT = take(288 * BP);
  w.feat_hi = take(BP * KA);
  w.feat_lo = take(BP * KA);
  w.vpT = take((size_t)NP * BP);
  w.part_stride = (int64_t)std::max(NSPLIT, fused_fwd_slots(m, w.BP)) * NACC * (int64_t)BP;
  w.part = take((size_t)w.part_stride * m->n_pass);
  w.gT = take(NACC * BP);
  w.pred = take(BP * NACC);
  w.dvp_hi = take(BP * (size_t)NP);
  w.dvp_lo = take(BP * (size_t)NP);
  w.dAflush = take((size_t)m->flush_off[m->n_pass] * 12 * BP);
  w.dAT = take(288 * BP);
  w.dfeat = take((size_t)KSPLIT_MAX * BP * KA);
  {
    // tiles = (BP/128) * ksplit CTAs of the backward blend GEMM: pick the split that wastes the
    // least of the last wave (ties -> fewer splits, less partial traffic)
    const int cand[4] = {6, 9, 12, 18};
    const int64_t mt = w.BP / 128;
    double best = -1.0;
    w.ksplit = 6;
    for (int c : cand) {
      const int64_t tiles = mt * c;
      const int64_t waves = (tiles + m->num_sms - 1) / m->num_sms;
      const double eff = (double)tiles / (double)(waves * m->num_sms);
      if (eff > best + 1e-9) { best = eff; w.ksplit = c; }
    }
  }
  w.dJp = take(BP * 72);
  w.Jp = take(BP * 72);
  w.d30T = take(90 * BP);
  w.loss_part = take(LOSS_PART_POSE + BP / 8 + 64);
  w.h_hi = take(BP * C_H);
  w.h_lo = take(BP * C_H);
  w.z1_hi = take(BP * C_Z);
  w.z1_lo = take(BP * C_Z);
  w.z2_hi = take(BP * C_Z);
  w.z2_lo = take(BP * C_Z);
  w.dz2_hi = take(BP * C_Z);
  w.dz2_lo = take(BP * C_Z);
  w.dz1_hi = take(BP * C_Z);
  w.dz1_lo = take(BP * C_Z);
  w.dh = take(BP * C_H);
  w.dzj = take(BP * NJ);
  w.dx6c = take(BP * 144);
  w.dbeta_s = take(BP * NB);
  w.shape_part = take(BP / 128 + 64);
  w.zj = take(BP * NJ);
  w.zg_part = take(BP * (C_Z / 128));
  w.dzg = take(BP);
  w.cmask = reinterpret_cast<uint2*>(take(BP * NJ * 2));
  w.zmask = reinterpret_cast<uint32_t*>(take(BP * (C_Z / 32)));
  w.zmask2 = reinterpret_cast<uint32_t*>(take(BP * (C_Z / 32)));
  w.gx6 = take(BP * 144);
  w.gbetas = take(BP * NB);
  w.adam_coef = take(64);
  w.scores = take(BP * 25);
  w.bytes = off;
  return w;
}

static int check_common(const JrrModel* m, int64_t B, const void* ws, size_t ws_bytes, Workspace* w) {
  if (!m) return fail(JRR_ERR_INVALID, "null model");
  if (B <= 0 || B > MAX_POSES_PER_CALL) return fail(JRR_ERR_INVALID, "B out of range (1..262144 per call)");
  if (!ws) return fail(JRR_ERR_WORKSPACE, "null workspace");
  if (((uintptr_t)ws & 255) != 0) return fail(JRR_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  *w = carve(m, B, const_cast<void*>(ws));
  if (w->bytes > ws_bytes) return fail(JRR_ERR_WORKSPACE, "workspace too small; see jrr_workspace_bytes");
  reset_launch_count();
  return JRR_OK;
}

// pose decode + chain, then the augmented blend GEMM: feat[BP,224] x Pt[20736,224]^T -> vpT
static int blend_forward_gemm(const JrrModel* m, const Workspace& w, cudaStream_t st) {
  GemmDesc g{};
  g.A_hi = w.feat_hi; g.A_lo = w.feat_lo; g.lda = KA;
  g.B_hi = m->Pt_hi; g.B_lo = m->Pt_lo; g.ldb = KA;
  g.M = w.BP; g.N = NP; g.K = KA; g.ksplit = 1; g.epi = EPI_STORE_T;
  g.out0 = w.vpT; g.ldo = w.BP;
  return launch_gemm(m, g, st);
}

static int forward_common(const JrrModel* m, const Workspace& w, const float* betas, const float* pose,
                          int kind, bool want_Jp, cudaStream_t st) {
  int rc = launch_pose_fwd(m, w.B, w.BP, betas, pose, kind, w.AT, w.feat_hi, w.feat_lo,
                           want_Jp ? w.Jp : nullptr, st);
  if (rc) return rc;
  return blend_forward_gemm(m, w, st);
}

// The fused forward once per skinning pass (models with more than four weights per vertex; SMPL: one pass).  Pass p > 0
// re-runs the blend GEMM with the vertex's next four weights: its regressor partial sums go to their own region of
// `part` (loss_seed adds the regions), skinned vertices are ADDED to the stored ones, blended vertices are not re-stored.
static int fused_fwd_all_passes(const JrrModel* cm, const Workspace& w, int store, float* out, cudaStream_t st, bool all_vertices) {
  JrrModel* m = const_cast<JrrModel*>(cm);
  int rc = JRR_OK;
  for (int p = 0; p < m->n_pass && rc == JRR_OK; p++) {
    m->select_pass(p);
    const int sp = p == 0 ? store : (store == 2 ? 3 : 0);
    rc = launch_fused_fwd(m, w, sp, out, st, all_vertices);
  }
  m->select_pass(0);
  return rc;
}

// The fused backward once per skinning pass: split-K partials of pass p go to slots [p*nsplit, (p+1)*nsplit) of dfeat, its dA
// flush events to their own region, and the dA reduction of pass p > 0 accumulates.  Sets w.ksplit for the chain backward.
static int fused_bwd_all_passes(const JrrModel* cm, Workspace& w, cudaStream_t st, const float* dvT, cudaEvent_t* ev_mid) {
  JrrModel* m = const_cast<JrrModel*>(cm);
  const bool small = dvT != nullptr && module_small_ranges(m, w.BP);
  const int nsplit = dvT != nullptr ? (small ? NSPLIT_S : NSPLIT_B) : m->nsplit_act;
  if (nsplit * m->n_pass > KSPLIT_MAX) return fail(JRR_ERR_INVALID, "too many skinning passes for the split-K workspace");
  int rc = JRR_OK;
  for (int p = 0; p < m->n_pass && rc == JRR_OK; p++) {
    m->select_pass(p);
    rc = launch_fused_bwd(m, w, st, dvT);
  }
  if (ev_mid && rc == JRR_OK) { cudaError_t e = cudaEventRecord(*ev_mid, st); if (e != cudaSuccess) rc = fail(JRR_ERR_CUDA, cudaGetErrorString(e)); }
  for (int p = 0; p < m->n_pass && rc == JRR_OK; p++) {
    m->select_pass(p);
    rc = launch_dA_reduce(m, w, dvT == nullptr ? 1 : (small ? 2 : 0), st);
  }
  m->select_pass(0);
  w.ksplit = nsplit * m->n_pass;
  return rc;
}

// Forward of the loss path up to the regressor partial sums.  store: 0 nothing, 1 blended
// vertices vp -> out (pose-contiguous, the backward needs them), 2 skinned vertices -> out.
static int loss_forward(const JrrModel* m, const Workspace& w, const float* betas, const float* pose, int kind,
                        int store, float* out, cudaStream_t st, cudaEvent_t* ev_after_pose,
                        cudaEvent_t* ev_after_gemm) {
  if (int rc = launch_pose_fwd(m, w.B, w.BP, betas, pose, kind, w.AT, w.feat_hi, w.feat_lo, nullptr, st)) return rc;
  if (ev_after_pose) JRR_CUDA(cudaEventRecord(*ev_after_pose, st));
  if (m->fused_fwd) {
    if (int rc = fused_fwd_all_passes(m, w, store, out, st, false)) return rc;
    if (ev_after_gemm) JRR_CUDA(cudaEventRecord(*ev_after_gemm, st));
    return JRR_OK;
  }
  if (int rc = blend_forward_gemm(m, w, st)) return rc;      // writes w.vpT
  if (ev_after_gemm) JRR_CUDA(cudaEventRecord(*ev_after_gemm, st));
  return launch_skin_fwd(m, w, nullptr, store == 2 ? out : nullptr, true, st);
}

// dfeat[s][BP,224] = dvp[BP, 20736 (split s)] x P[224, 20736]^T
static int blend_backward_gemm(const JrrModel* m, const Workspace& w, cudaStream_t st) {
  GemmDesc g{};
  g.A_hi = w.dvp_hi; g.A_lo = w.dvp_lo; g.lda = NP;
  g.B_hi = m->P_hi; g.B_lo = m->P_lo; g.ldb = NP;
  g.M = w.BP; g.N = KA; g.K = NP / w.ksplit; g.ksplit = w.ksplit; g.epi = EPI_STORE_SPLITK;
  g.out0 = w.dfeat; g.ldo = KA;
  return launch_gemm(m, g, st);
}

int launch_gemm(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  if (m->gemm_impl == 1) return launch_gemm_simt(g, st);
  return launch_gemm_tc(m, g, st);
}

}  // namespace jrr

using namespace jrr;

extern "C" size_t jrr_workspace_bytes(const JrrModel* m, int64_t B) {
  if (!m || B <= 0) return 0;
  return carve(m, B, nullptr).bytes;
}

extern "C" int jrr_smpl_forward(JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                                float* vertices_out, float* joints49_out, void* ws, size_t ws_bytes,
                                void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!betas || !pose) return fail(JRR_ERR_INVALID, "null input");
  if (!vertices_out && !joints49_out) return fail(JRR_ERR_INVALID, "no output requested");
  cudaStream_t st = (cudaStream_t)stream;
  static const bool fused_module = [] { const char* e = getenv("JRR_FUSED_MODULE"); return !(e && e[0] == '0'); }();
  if (smpl_small_fwd_available(m, B)) {
    // a handful of poses: chain, blend, skinning, vertex store and the 49 joints in ONE launch (warp per vertex)
    // (joints49 reads the stored vertices: the caller's buffer, else scratch -- dvp_hi is [BP][NP] >= [B][6890][3])
    return launch_smpl_small_fwd(m, B, betas, pose, kind, vertices_out ? vertices_out : w.dvp_hi, joints49_out, st);
  }
  if (m->gemm_impl == 0 && m->fused_fwd && fused_module) {
    // chain | blend GEMM with the skinning epilogue over EVERY packed vertex (pose-contiguous store) | un-packing to the
    // model's vertex order with coalesced reads and writes | 49 joints gathered from the packed vertices
    if (int rc = launch_pose_fwd(m, w.B, w.BP, betas, pose, kind, w.AT, w.feat_hi, w.feat_lo,
                                 joints49_out ? w.Jp : nullptr, st)) return rc;
    if (int rc = fused_fwd_all_passes(m, w, 2, w.vpT, st, true)) return rc;
    if (vertices_out)
      if (int rc = launch_unpack_vertices(m, w, w.vpT, vertices_out, st)) return rc;
    if (joints49_out)
      if (int rc = launch_joints49_fwd_packed(m, w, w.vpT, joints49_out, st)) return rc;
    return JRR_OK;
  }
  if (m->n_pass > 1) return fail(JRR_ERR_STATE, "models with more than 4 skinning weights per vertex need the fused kernels");
  if (int rc = forward_common(m, w, betas, pose, kind, joints49_out != nullptr, st)) return rc;
  // joints49 reads vertices: use caller's buffer, else scratch (dvp_hi is [BP][NP] >= [B][6890][3])
  float* verts = vertices_out ? vertices_out : w.dvp_hi;
  if (int rc = launch_skin_fwd(m, w, verts, nullptr, false, st)) return rc;
  if (joints49_out)
    if (int rc = launch_joints49_fwd(m, w, verts, joints49_out, st)) return rc;
  return JRR_OK;
}

extern "C" int jrr_smpl_backward(JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                                 const float* dvertices, const float* djoints49, float* dbetas_out,
                                 float* dpose_out, void* ws, size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!betas || !pose || !dbetas_out || !dpose_out) return fail(JRR_ERR_INVALID, "null argument");
  if (!dvertices && !djoints49) return fail(JRR_ERR_INVALID, "no upstream gradient given");
  cudaStream_t st = (cudaStream_t)stream;
  const bool use_x = djoints49 != nullptr;
  static const bool fused_module = [] { const char* e = getenv("JRR_FUSED_MODULE"); return !(e && e[0] == '0'); }();
  if (smpl_small_bwd_available(m, B)) {
    // a handful of poses: warp-per-vertex backward (chain, blended vertex, dA and dfeat partials), a fixed-order reduction,
    // then the unchanged chain backward
    if (int rc = launch_smpl_small_bwd(m, w, betas, pose, kind, dvertices, djoints49, st)) return rc;
    return launch_pose_bwd(m, w, betas, pose, kind, use_x, false, false, dbetas_out, dpose_out, nullptr, nullptr,
                           nullptr, nullptr, nullptr, 0.f, st);
  }
  if (m->gemm_impl == 0 && m->fused_fwd && m->fused_bwd && fused_module) {
    // recompute: chain | blend GEMM + skinning (stores the blended vertices, pose-contiguous) ; then
    // joints49 gradient -> its sources | re-pack d vertices | skinning backward generating the A operand of the
    // blend-gradient GEMM (d blended vertices never reach memory) | dA reduction | chain backward
    if (int rc = launch_pose_fwd(m, w.B, w.BP, betas, pose, kind, w.AT, w.feat_hi, w.feat_lo, nullptr, st)) return rc;
    if (int rc = launch_fused_fwd(m, w, 1, w.vpT, st, true)) return rc;      // (the blended vertices do not depend on the pass)
    if (use_x)
      if (int rc = launch_joints49_bwd(m, w, djoints49, st)) return rc;
    float* dvT = w.dvp_hi;
    if (int rc = launch_pack_dvertices(m, w, dvertices, use_x, dvT, st)) return rc;
    if (int rc = fused_bwd_all_passes(m, w, st, dvT, nullptr)) return rc;
    return launch_pose_bwd(m, w, betas, pose, kind, use_x, false, false, dbetas_out, dpose_out, nullptr, nullptr,
                           nullptr, nullptr, nullptr, 0.f, st);
  }
  if (m->n_pass > 1) return fail(JRR_ERR_STATE, "models with more than 4 skinning weights per vertex need the fused kernels");
  if (int rc = forward_common(m, w, betas, pose, kind, false, st)) return rc;
  if (use_x)
    if (int rc = launch_joints49_bwd(m, w, djoints49, st)) return rc;
  if (int rc = launch_skin_bwd(m, w, dvertices, false, use_x, st)) return rc;
  if (int rc = launch_dA_reduce(m, w, 0, st)) return rc;
  if (int rc = blend_backward_gemm(m, w, st)) return rc;
  return launch_pose_bwd(m, w, betas, pose, kind, use_x, false, false, dbetas_out, dpose_out, nullptr, nullptr,
                         nullptr, nullptr, nullptr, 0.f, st);
}

extern "C" int jrr_find_joints(JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                               float* joints17_out, void* ws, size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  if (!betas || !pose || !joints17_out) return fail(JRR_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = loss_forward(m, w, betas, pose, kind, 0, nullptr, st, nullptr, nullptr)) return rc;
  return launch_loss_seed(m, w, m->fused_fwd, nullptr, 1, 0.f, joints17_out, Proj2D{}, st);
}

extern "C" int jrr_find_joints_backward(JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                                        const float* djoints17, float* dbetas_out, float* dpose_out, void* ws,
                                        size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  if (!betas || !pose || !djoints17 || !dbetas_out || !dpose_out) return fail(JRR_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  // recompute the forward (blended vertices kept for the backward), seed the skinning backward with the caller's gradient
  if (int rc = loss_forward(m, w, betas, pose, kind, 1, w.vpT, st, nullptr, nullptr)) return rc;
  if (int rc = launch_seed_from_dpred(w, djoints17, st)) return rc;
  if (m->fused_bwd) {
    if (int rc = fused_bwd_all_passes(m, w, st, nullptr, nullptr)) return rc;
  } else {
    if (m->n_pass > 1) return fail(JRR_ERR_STATE, "models with more than 4 skinning weights per vertex need the fused kernels");
    if (int rc = launch_skin_bwd(m, w, nullptr, true, false, st)) return rc;
    if (int rc = launch_dA_reduce(m, w, 0, st)) return rc;
    if (int rc = blend_backward_gemm(m, w, st)) return rc;
  }
  return launch_pose_bwd(m, w, betas, pose, kind, false, false, false, dbetas_out, dpose_out, nullptr, nullptr, nullptr,
                         nullptr, nullptr, 0.f, st);
}

extern "C" int jrr_critic_forward(JrrModel* m, int64_t B, const float* rot6d, float* scores_out, void* ws,
                                  size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_critic) return fail(JRR_ERR_STATE, "jrr_critic_load has not been called");
  if (!rot6d || !scores_out) return fail(JRR_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_critic_pre(m, w, rot6d, st)) return rc;
  if (int rc = critic_forward_gemms(m, w, st)) return rc;
  return launch_critic_head(m, w, B, 0.f, scores_out, false, st);
}

extern "C" int jrr_shape_critic_forward(JrrModel* m, int64_t B, const float* betas, float* scores_out, void* stream) {
  if (!m || !betas || !scores_out) return fail(JRR_ERR_INVALID, "null argument");
  if (B <= 0 || B > MAX_POSES_PER_CALL) return fail(JRR_ERR_INVALID, "B out of range");
  if (!m->has_shape_critic) return fail(JRR_ERR_STATE, "jrr_shape_critic_load has not been called");
  reset_launch_count();
  return launch_shape_critic_scores(m, B, betas, scores_out, (cudaStream_t)stream);
}

// Kernel groups of one refinement step (order of execution); jrr_refine_step_profiled
// reports one duration per group.
static const char* const kStepKernelNames[JRR_STEP_KERNELS] = {
    "pose_fwd", "blend_gemm_fwd", "skin_fwd", "loss_seed", "skin_bwd", "dA_reduce", "blend_gemm_bwd",
    "critic_pre", "critic_gemm_fwd", "critic_head", "critic_gemm_bwd", "critic_post", "loss_finish",
    "pose_bwd_adam"};

static int refine_step_impl(JrrModel* m, int64_t B, int64_t B_logical, float* x6, float* betas,
                            const float* gt_mm, float* adam_m, float* adam_v, int32_t* step_count, float lr,
                            float w_joint, float w_pose, float* loss_out, void* ws, size_t ws_bytes,
                            cudaStream_t st, cudaEvent_t* ev, Proj2D p2d = Proj2D{}, float w_2d = 0.f) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  if (m->folded && !m->T_hi) return fail(JRR_ERR_STATE, "folded loss path selected before a regressor was set");
  if (!x6 || !betas || !gt_mm || !adam_m || !adam_v || !step_count) return fail(JRR_ERR_INVALID, "null argument");
  if (((uintptr_t)x6 | (uintptr_t)betas | (uintptr_t)adam_m | (uintptr_t)adam_v) & 7)
    return fail(JRR_ERR_INVALID, "x6 / betas / adam_m / adam_v must be 8-byte aligned");
  if (B_logical < B) return fail(JRR_ERR_INVALID, "B_logical must be >= B");
  const bool critic = w_pose != 0.f;
  if (critic && !m->has_critic) return fail(JRR_ERR_STATE, "w_pose != 0 but jrr_critic_load has not been called");
  const bool shape = m->has_shape_critic && m->w_shape != 0.f;   // Shape_Discriminator term (optimize.py:244,249-250)
  int mark = 0;
#define JRR_MARK()                                                      \
  do {                                                                  \
    if (ev) JRR_CUDA(cudaEventRecord(ev[mark], st));                    \
    mark++;                                                             \
  } while (0)
  // The critic chain only reads x6 and its own workspace slices, so outside profiling it is
  // forked onto the model's side stream and joins again before the Adam kernel (inside a
  // CUDA-graph capture this becomes a parallel branch of the graph).
  const bool hf = m->critic_head_fused;      // global head inside the layer-2 GEMM epilogue
  const bool fork = critic && ev == nullptr && m->overlap_critic;
  cudaStream_t cs = fork ? m->side : st;
  if (fork) {
    JRR_CUDA(cudaEventRecord(m->ev_fork, st));
    JRR_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
  }
  static const int dbg_skip_c = [] { const char* e = getenv("JRR_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
  if (fork && (dbg_skip_c & 2)) {
    JRR_CUDA(cudaEventRecord(m->ev_join, m->side));
  } else if (fork) {
    if (int rc = launch_critic_pre(m, w, x6, cs, hf)) return rc;
    if (int rc = critic_forward_gemms(m, w, cs, hf)) return rc;
    const bool hl = hf && m->critic_headless;      // no head kernel on the chain
    if (hl) {
      if (int rc = critic_backward_gemms(m, w, cs, nullptr, w_pose * 2.f / (25.f * (float)B_logical))) return rc;
      if (int rc = launch_critic_post(m, w, x6, cs, true, B_logical, w_pose)) return rc;
    } else {
      if (hf) { if (int rc = launch_critic_head_light(m, w, B_logical, w_pose, cs)) return rc; }
      else if (int rc = launch_critic_head(m, w, B_logical, w_pose, nullptr, true, cs)) return rc;
      if (int rc = critic_backward_gemms(m, w, cs, hf ? w.dzg : nullptr)) return rc;
      if (int rc = launch_critic_post(m, w, x6, cs)) return rc;
    }
    if (shape) if (int rc = launch_shape_critic(m, w, betas, B_logical, cs)) return rc;
    JRR_CUDA(cudaEventRecord(m->ev_join, m->side));
  }
  JRR_MARK();
  if (fork && m->split_adam)
    if (int rc = launch_adam_coef(w, step_count, lr, st)) return rc;
  // diagnostic only (benchmarks/step_breakdown.py): JRR_DEBUG_SKIP bit 0 = leave out the loss-path kernels of the main
  // branch (results meaningless), bit 1 = leave out the critic chain -- what each branch costs the step when it runs alone
  static const int dbg_skip = [] { const char* e = getenv("JRR_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
  if (dbg_skip & 1) {
    if (fork) JRR_CUDA(cudaEventRecord(m->ev_seed, st));
    for (int i = 0; i < 7; i++) JRR_MARK();
  } else if (m->folded) {
    // folded loss path: chain | Q = feat . T^T | per-frame joints + loss seed + dA + dQ | dfeat = dQ . T (split-K)
    // (events: pose_fwd | blend_gemm_fwd [the N = 1224 GEMM] | skin_fwd [empty] | loss_seed [folded seed] |
    //  skin_bwd [empty] | dA_reduce [empty] | blend_gemm_bwd [the K = 1280 GEMM])
    // fold_ts: features and dQ stay plain fp32 (half the bytes) and both GEMMs run on CTA pairs, A through tensor memory
    const bool fts = m->fold_ts && m->gemm_impl == 0 && w.BP >= 256;
    if (int rc = launch_pose_fwd(m, w.B, w.BP, betas, x6, JRR_POSE_ROT6D, w.AT, w.feat_hi, fts ? nullptr : w.feat_lo, nullptr, st)) return rc;
    JRR_MARK();
    {
      GemmDesc g{};
      g.a_via_tmem = fts;
      g.A_hi = w.feat_hi; g.A_lo = w.feat_lo; g.lda = KA;
      g.B_hi = m->T_hi; g.B_lo = m->T_lo; g.ldb = KA;
      g.M = w.BP; g.N = FOLD_NP; g.K = KA; g.ksplit = 1; g.epi = EPI_STORE_T;
      g.out0 = w.vpT; g.ldo = w.BP;
      if (int rc = launch_gemm(m, g, st)) return rc;
    }
    JRR_MARK();
    JRR_MARK();
    if (int rc = launch_folded_seed(m, w, gt_mm, B_logical, w_joint, nullptr, p2d, st, nullptr, fts)) return rc;
    if (fork) JRR_CUDA(cudaEventRecord(m->ev_seed, st));
    JRR_MARK();
    JRR_MARK();
    JRR_MARK();
    {
      // (CTA pairs: 256-row tiles, one wave of BP / 256 * 4 tiles up to 4736 frames)
      w.ksplit = (fts || (w.BP / 128) * 4 >= m->num_sms) ? 4 : 8;
      GemmDesc g{};
      g.a_via_tmem = fts;
      g.A_hi = w.dvp_hi; g.A_lo = w.dvp_lo; g.lda = FOLD_NP;
      g.B_hi = m->Tt_hi; g.B_lo = m->Tt_lo; g.ldb = FOLD_NP;
      g.M = w.BP; g.N = KA; g.K = FOLD_NP / w.ksplit; g.ksplit = w.ksplit; g.epi = EPI_STORE_SPLITK;
      g.k_valid = FOLD_N;                      // the seed kernel writes 1224 columns per row; TMA zero-fills the K padding
      g.out0 = w.dfeat; g.ldo = KA;
      if (int rc = launch_gemm(m, g, st)) return rc;
    }
    JRR_MARK();
  } else {
  // forward: chain | blend GEMM with the skinning + 17x6890 regressor epilogue | loss seed
    // (events: pose_fwd | blend_gemm_fwd [fused: the whole forward] | skin_fwd [fused: empty])
    if (int rc = loss_forward(m, w, betas, x6, JRR_POSE_ROT6D, 1, w.vpT, st, ev ? &ev[1] : nullptr, ev ? &ev[2] : nullptr)) return rc;
    mark = 3;
    JRR_MARK();
    if (int rc = launch_loss_seed(m, w, m->fused_fwd, gt_mm, B_logical, w_joint, nullptr, p2d, st)) return rc;
    if (fork) JRR_CUDA(cudaEventRecord(m->ev_seed, st));
    JRR_MARK();
    // backward: regressor-transpose seed + skinning | dA reduction | blend GEMM
    // (events: skin_bwd [fused: skinning backward + blend-gradient GEMM] | dA_reduce | blend_gemm_bwd [fused: empty])
    if (m->fused_bwd) {
      if (int rc = fused_bwd_all_passes(m, w, st, nullptr, ev ? &ev[mark] : nullptr)) return rc;   // (event: end of the fused kernels)
      mark += 1;
      JRR_MARK();
      JRR_MARK();
    } else {
      if (int rc = launch_skin_bwd(m, w, nullptr, true, false, st)) return rc;
      JRR_MARK();
      if (int rc = launch_dA_reduce(m, w, 0, st)) return rc;
      JRR_MARK();
      if (int rc = blend_backward_gemm(m, w, st)) return rc;
      JRR_MARK();
    }
  }
  // critic forward + input gradient (inline when profiling or when the fork is disabled)
  const bool inl = critic && !fork;
  if (inl) if (int rc = launch_critic_pre(m, w, x6, st, hf)) return rc;
  JRR_MARK();
  if (inl) if (int rc = critic_forward_gemms(m, w, st, hf)) return rc;
  JRR_MARK();
  const bool hl_inl = hf && m->critic_headless;
  if (inl && !hl_inl) {
    if (hf) { if (int rc = launch_critic_head_light(m, w, B_logical, w_pose, st)) return rc; }
    else if (int rc = launch_critic_head(m, w, B_logical, w_pose, nullptr, true, st)) return rc;
  }
  JRR_MARK();
  if (inl) if (int rc = critic_backward_gemms(m, w, st, hl_inl ? nullptr : (hf ? w.dzg : nullptr),
                                              hl_inl ? w_pose * 2.f / (25.f * (float)B_logical) : 0.f)) return rc;
  JRR_MARK();
  if (inl) if (int rc = launch_critic_post(m, w, x6, st, hl_inl, B_logical, w_pose)) return rc;
  if (shape && !fork) if (int rc = launch_shape_critic(m, w, betas, B_logical, st)) return rc;
  JRR_MARK();
  // With the critic on its own branch, the chain backward does not have to wait for it: it leaves the
  // parameter gradients in the workspace and an element-wise Adam kernel runs after the join.
  // (external gradients -- jrr_set_external_gradient -- are added by the element-wise Adam kernel, whatever the branches)
  const bool ext = m->ext_dx6 != nullptr || m->ext_dbetas != nullptr;
  const bool split = (fork && m->split_adam) || ext;
  if (split && !(dbg_skip & 1))
    if (int rc = launch_pose_bwd(m, w, betas, x6, JRR_POSE_ROT6D, false, false, false, w.gbetas, w.gx6, nullptr, nullptr,
                                 nullptr, nullptr, nullptr, 0.f, st)) return rc;
  // The loss read-out does not feed the update: with the fork it runs on the critic's stream (after the
  // seed kernel's partials, ev_seed) beside the Adam kernel, and the step ends when both have.
  // (round 2, K = 100, two runs each: 0.2866 / 0.2873 ms against 0.2902 / 0.2900 ms inline; JRR_FINISH_ASIDE=0 switches it off)
  static const bool aside_on = [] { const char* e = getenv("JRR_FINISH_ASIDE"); return !(e && e[0] == '0'); }();
  const bool finish_aside = aside_on && fork && m->split_adam && loss_out != nullptr;
  if (finish_aside) {
    JRR_CUDA(cudaStreamWaitEvent(m->side, m->ev_seed, 0));
    if (int rc = launch_loss_finish(w, B_logical, w_joint, w_pose, critic, w_2d, shape ? m->w_shape : 0.f, loss_out, nullptr, m->side)) return rc;
    JRR_CUDA(cudaEventRecord(m->ev_join2, m->side));
  }
  if (fork) JRR_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
  if (loss_out && !finish_aside)
    if (int rc = launch_loss_finish(w, B_logical, w_joint, w_pose, critic, w_2d, shape ? m->w_shape : 0.f, loss_out, nullptr, st)) return rc;
  JRR_MARK();
  // chain backward + Adam
  if (split) {
    if (int rc = launch_adam_params(w, critic, shape, x6, betas, adam_m, adam_v, step_count, lr, st, m->ext_dx6, m->ext_dbetas)) return rc;
    if (finish_aside) JRR_CUDA(cudaStreamWaitEvent(st, m->ev_join2, 0));
  } else if (int rc = launch_pose_bwd(m, w, betas, x6, JRR_POSE_ROT6D, false, critic, shape, nullptr, nullptr, x6, betas,
                                      adam_m, adam_v, step_count, lr, st)) return rc;
  JRR_MARK();
#undef JRR_MARK
  return JRR_OK;
}

extern "C" int jrr_set_external_gradient(JrrModel* m, const float* dx6, const float* dbetas, const float* dcam) {
  if (!m) return fail(JRR_ERR_INVALID, "null model");
  if (((uintptr_t)dx6 | (uintptr_t)dbetas) & 7) return fail(JRR_ERR_INVALID, "external gradients must be 8-byte aligned");
  m->ext_dx6 = dx6;
  m->ext_dbetas = dbetas;
  m->ext_dcam = dcam;
  return JRR_OK;
}

extern "C" int jrr_refine_step(JrrModel* m, int64_t B, int64_t B_logical, float* x6, float* betas,
                               const float* gt_mm, float* adam_m, float* adam_v, int32_t* step_count,
                               float lr, float w_joint, float w_pose, float* loss_out, void* ws,
                               size_t ws_bytes, void* stream) {
  return refine_step_impl(m, B, B_logical, x6, betas, gt_mm, adam_m, adam_v, step_count, lr, w_joint, w_pose,
                          loss_out, ws, ws_bytes, (cudaStream_t)stream, nullptr);
}

extern "C" int jrr_refine_step_2d(JrrModel* m, int64_t B, int64_t B_logical, float* x6, float* betas,
                                  const float* gt_mm, const float* gt_j2d, float* cam, float* adam_m, float* adam_v,
                                  float* cam_adam_m, float* cam_adam_v, int32_t* step_count, float lr, float w_joint,
                                  float w_pose, float w_2d, float* loss_out, void* ws, size_t ws_bytes, void* stream) {
  if (!gt_j2d || !cam || !cam_adam_m || !cam_adam_v) return fail(JRR_ERR_INVALID, "null 2-D argument");
  Proj2D p2d;
  p2d.gt2d = gt_j2d; p2d.cam = cam; p2d.cam_m = cam_adam_m; p2d.cam_v = cam_adam_v;
  p2d.step_count = step_count; p2d.lr = lr;
  p2d.scale = w_2d * 2.f / (34.f * (float)B_logical);
  p2d.dcam_ext = m ? m->ext_dcam : nullptr;
  return refine_step_impl(m, B, B_logical, x6, betas, gt_mm, adam_m, adam_v, step_count, lr, w_joint, w_pose,
                          loss_out, ws, ws_bytes, (cudaStream_t)stream, nullptr, p2d, w_2d);
}

extern "C" int jrr_camera_fit(JrrModel* m, int64_t B, int64_t B_logical, const float* x6, const float* betas,
                              const float* gt_j2d, float* cam, int iters, float lr, float* loss_out, void* ws,
                              size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  if (!x6 || !betas || !gt_j2d || !cam || iters < 0) return fail(JRR_ERR_INVALID, "bad argument");
  if (B_logical < B) return fail(JRR_ERR_INVALID, "B_logical must be >= B");
  cudaStream_t st = (cudaStream_t)stream;
  // the 3-D joints do not depend on the camera: one forward, then every frame iterates privately
  if (int rc = loss_forward(m, w, betas, x6, JRR_POSE_ROT6D, 0, nullptr, st, nullptr, nullptr)) return rc;
  if (int rc = launch_loss_seed(m, w, m->fused_fwd, nullptr, 1, 0.f, w.pred, Proj2D{}, st)) return rc;
  return launch_camera_fit(w, w.pred, gt_j2d, cam, iters, lr, B_logical, loss_out, st);
}

extern "C" int jrr_critic_layer2_bwd_products(const JrrModel* m, int64_t B) {
  if (!m || B <= 0) return 3;
  const bool ts = m->critic_head_fused && m->critic_ts;
  return (ts && gemm_pair_bits_available(m, round_up(B, 256))) ? 2 : 3;
}

extern "C" const char* jrr_step_kernel_name(int i) {
  return (i >= 0 && i < JRR_STEP_KERNELS) ? kStepKernelNames[i] : nullptr;
}

extern "C" int jrr_refine_step_profiled(JrrModel* m, int64_t B, int64_t B_logical, float* x6, float* betas,
                                        const float* gt_mm, float* adam_m, float* adam_v,
                                        int32_t* step_count, float lr, float w_joint, float w_pose,
                                        float* loss_out, void* ws, size_t ws_bytes, void* stream,
                                        float* ms_out_host) {
  if (!ms_out_host) return fail(JRR_ERR_INVALID, "null ms_out_host");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t ev[JRR_STEP_KERNELS + 1];
  for (int i = 0; i <= JRR_STEP_KERNELS; i++) JRR_CUDA(cudaEventCreate(&ev[i]));
  int rc = refine_step_impl(m, B, B_logical, x6, betas, gt_mm, adam_m, adam_v, step_count, lr, w_joint, w_pose,
                            loss_out, ws, ws_bytes, st, ev);
  if (rc == JRR_OK) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(JRR_ERR_CUDA, cudaGetErrorString(e));
  }
  if (rc == JRR_OK)
    for (int i = 0; i < JRR_STEP_KERNELS; i++) cudaEventElapsedTime(&ms_out_host[i], ev[i], ev[i + 1]);
  for (int i = 0; i <= JRR_STEP_KERNELS; i++) cudaEventDestroy(ev[i]);
  return rc;
}

extern "C" int jrr_regressor_grad_accumulate(JrrModel* m, int64_t B, int64_t B_logical, const float* x6,
                                             const float* betas, const float* gt_mm, float* G_accum,
                                             float* loss_accum, void* ws, size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  if (!x6 || !betas || !gt_mm || !G_accum) return fail(JRR_ERR_INVALID, "null argument");
  if (B_logical < B) return fail(JRR_ERR_INVALID, "B_logical must be >= B");
  cudaStream_t st = (cudaStream_t)stream;
  static const bool folded_refit = [] { const char* e = getenv("JRR_FOLDED_REFIT"); return !(e && e[0] == '0'); }();
  if (m->folded && m->T_hi && folded_refit && m->gemm_impl == 0) {
    // folded form: chain | Q = feat . T^T | per-frame seed | G += unfold(dQ^T . feat, dc)   (jrr_model.cu)
    if (int rc = launch_pose_fwd(m, w.B, w.BP, betas, x6, JRR_POSE_ROT6D, w.AT, w.feat_hi, w.feat_lo, nullptr, st)) return rc;
    GemmDesc g{};
    g.A_hi = w.feat_hi; g.A_lo = w.feat_lo; g.lda = KA;
    g.B_hi = m->T_hi; g.B_lo = m->T_lo; g.ldb = KA;
    g.M = w.BP; g.N = FOLD_NP; g.K = KA; g.ksplit = 1; g.epi = EPI_STORE_T;
    g.out0 = w.vpT; g.ldo = w.BP;
    if (int rc = launch_gemm(m, g, st)) return rc;
    if (int rc = regressor_accumulate_folded(m, w, gt_mm, B_logical, G_accum, st)) return rc;
    if (loss_accum)
      if (int rc = launch_loss_finish(w, B_logical, 1.f, 0.f, false, 0.f, 0.f, nullptr, loss_accum, st)) return rc;
    return JRR_OK;
  }
  // skinned vertices kept pose-contiguous in the (otherwise idle) dvp_hi buffer
  float* vT = w.dvp_hi;
  if (int rc = loss_forward(m, w, betas, x6, JRR_POSE_ROT6D, 2, vT, st, nullptr, nullptr)) return rc;
  if (int rc = launch_loss_seed(m, w, m->fused_fwd, gt_mm, B_logical, 1.f, nullptr, Proj2D{}, st)) return rc;
  if (int rc = launch_regressor_accumulate(m, w, vT, G_accum, st)) return rc;
  if (loss_accum)
    if (int rc = launch_loss_finish(w, B_logical, 1.f, 0.f, false, 0.f, 0.f, nullptr, loss_accum, st)) return rc;
  return JRR_OK;
}

extern "C" int jrr_regressor_apply(JrrModel* m, float* J17_raw, const float* mask, const float* G_accum,
                                   float* adam_m, float* adam_v, int32_t* step_count, float lr,
                                   void* stream) {
  if (!m || !J17_raw || !G_accum || !adam_m || !adam_v || !step_count) return fail(JRR_ERR_INVALID, "null argument");
  if (!m->has_regressor) return fail(JRR_ERR_STATE, "jrr_set_regressor has not been called");
  reset_launch_count();
  return launch_regressor_apply(m, J17_raw, mask, G_accum, adam_m, adam_v, step_count, lr, (cudaStream_t)stream);
}

// ---- widening row 8f-1: critic training step (optimize.py:276-293) ------------------------------
extern "C" int jrr_critic_grad_accumulate(JrrModel* m, int64_t B, int64_t B_logical, const float* x6, float target,
                                          float* G_accum, float* loss_accum, void* ws, size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_critic) return fail(JRR_ERR_STATE, "jrr_critic_load has not been called");
  if (!x6 || !G_accum) return fail(JRR_ERR_INVALID, "null argument");
  if (B_logical < B) return fail(JRR_ERR_INVALID, "B_logical must be >= B");
  if (B > 16384) return fail(JRR_ERR_INVALID, "at most 16384 poses per critic-gradient call (chunk and accumulate)");
  return critic_grad_accumulate(m, w, B_logical, x6, target, G_accum, loss_accum, (cudaStream_t)stream);
}

extern "C" int jrr_critic_apply(JrrModel* m, float* params, const float* G, float* adam_m, float* adam_v,
                                int32_t* step_count, float lr, void* stream) {
  if (!m || !params || !G || !adam_m || !adam_v || !step_count) return fail(JRR_ERR_INVALID, "null argument");
  reset_launch_count();
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_adam_flat(params, G, adam_m, adam_v, step_count, lr, JRR_CRITIC_PARAMS, st)) return rc;
  return critic_load_impl(m, params, st);
}

extern "C" int jrr_shape_critic_grad_accumulate(JrrModel* m, int64_t B, int64_t B_logical, const float* betas,
                                                float target, float* G_accum, float* loss_accum, void* ws,
                                                size_t ws_bytes, void* stream) {
  Workspace w;
  if (int rc = check_common(m, B, ws, ws_bytes, &w)) return rc;
  if (!m->has_shape_critic) return fail(JRR_ERR_STATE, "jrr_shape_critic_load has not been called");
  if (!betas || !G_accum) return fail(JRR_ERR_INVALID, "null argument");
  if (B_logical < B) return fail(JRR_ERR_INVALID, "B_logical must be >= B");
  return shape_critic_grad_accumulate(m, w, B_logical, betas, target, G_accum, loss_accum, (cudaStream_t)stream);
}

extern "C" int jrr_shape_critic_apply(JrrModel* m, float* params, const float* G, float* adam_m, float* adam_v,
                                      int32_t* step_count, float lr, void* stream) {
  if (!m || !params || !G || !adam_m || !adam_v || !step_count) return fail(JRR_ERR_INVALID, "null argument");
  if (!m->has_shape_critic) return fail(JRR_ERR_STATE, "jrr_shape_critic_load has not been called");
  reset_launch_count();
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_adam_flat(params, G, adam_m, adam_v, step_count, lr, JRR_SHAPE_CRITIC_PARAMS, st)) return rc;
  JRR_CUDA(cudaMemcpyAsync(m->shape_critic, params, JRR_SHAPE_CRITIC_PARAMS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return JRR_OK;
}

namespace jrr {
__global__ void debug_split_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ hi,
                                   float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r;
  const float x = src[i];
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  const float h = __uint_as_float(r);
  hi[i] = h;
  const float d = x - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  lo[i] = __uint_as_float(r);
}
}  // namespace jrr

extern "C" int jrr_debug_gemm(JrrModel* m, int impl, int64_t M, int64_t N, int64_t K, const float* A,
                              const float* B, float* C, float* scratch, void* stream) {
  if (!m || !A || !B || !C || !scratch) return fail(JRR_ERR_INVALID, "null argument");
  reset_launch_count();
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 2) {      // tcgen05 kernel with plain fp32 A staged through tensor memory (B pre-split here)
    float* Bh2 = scratch;
    float* Bl2 = Bh2 + N * K;
    debug_split_kernel<<<(unsigned)((N * K + 255) / 256), 256, 0, st>>>(B, N * K, Bh2, Bl2);
    JRR_LAUNCH_CHECK();
    GemmDesc g{};
    g.a_via_tmem = true;
    g.probe_env = true;
    g.A_hi = A; g.lda = K; g.B_hi = Bh2; g.B_lo = Bl2; g.ldb = K;
    g.M = M; g.N = N; g.K = K; g.ksplit = 1; g.epi = EPI_STORE_SPLITK;
    g.out0 = C; g.ldo = N;
    return launch_gemm_tc(m, g, st);
  }
  float* Ah = scratch;
  float* Al = Ah + M * K;
  float* Bh = Al + M * K;
  float* Bl = Bh + N * K;
  debug_split_kernel<<<(unsigned)((M * K + 255) / 256), 256, 0, st>>>(A, M * K, Ah, Al);
  JRR_LAUNCH_CHECK();
  debug_split_kernel<<<(unsigned)((N * K + 255) / 256), 256, 0, st>>>(B, N * K, Bh, Bl);
  JRR_LAUNCH_CHECK();
  GemmDesc g{};
  g.A_hi = Ah; g.A_lo = Al; g.lda = K;
  g.B_hi = Bh; g.B_lo = Bl; g.ldb = K;
  g.M = M; g.N = N; g.K = K; g.ksplit = 1; g.epi = EPI_STORE_SPLITK;
  g.out0 = C; g.ldo = N;
  return impl == 1 ? launch_gemm_simt(g, st) : launch_gemm_tc(m, g, st);
}
