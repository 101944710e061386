// Model packing (host) and the regressor / critic state kernels.
//
// jrr_model_create turns the smplx buffers into the layouts the kernels want:
//   * augmented blend matrix [posedirs ; shapedirs^T ; v_template] (K = 218 -> 224,
//     N = 3*6890 -> 20736), tf32 hi/lo split, in both majors
//   * J0 = J24.v_template and JS = J24.shapedirs so rest joints are J0 + JS.beta (72x10)
//   * skinning "run" records: <= 4 (joint, weight) slots per vertex, slot-stable so that the
//     kernels re-fetch a joint transform only when a slot's joint changes; the matching list
//     of dA flush events per joint
//   * CSR rows of J_regressor_extra, the 21 vertex picks and the 49-joint map
// Replaces: smplx.SMPL.__init__ buffers + scripts/smpl.py:64-70.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "jrr_internal.cuh"

namespace jrr {

static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int status, const std::string& msg) { g_err = msg; return status; }
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

float tf32_round_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  // round to nearest, ties away from zero in magnitude (matches cvt.rna)
  u += 0x00001000u;
  u &= 0xffffe000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

template <typename T>
static int upload(JrrModel* m, T** dst, const std::vector<T>& src) {
  JRR_CUDA(cudaMalloc((void**)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
  m->allocs.push_back(*dst);
  if (!src.empty()) JRR_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return JRR_OK;
}
template <typename T>
static int dalloc(JrrModel* m, T** dst, size_t n, bool zero = true) {
  JRR_CUDA(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
  m->allocs.push_back(*dst);
  if (zero) JRR_CUDA(cudaMemset(*dst, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return JRR_OK;
}

static void split_hi_lo(const std::vector<float>& x, std::vector<float>& hi, std::vector<float>& lo) {
  hi.resize(x.size());
  lo.resize(x.size());
  for (size_t i = 0; i < x.size(); i++) {
    hi[i] = tf32_round_host(x[i]);
    lo[i] = tf32_round_host(x[i] - hi[i]);
  }
}

// --------------------------------------------------------------------- regressor kernels
__global__ void reg_rowsum_kernel(const float* __restrict__ J, const float* __restrict__ mask,
                                  float* __restrict__ rowsum) {
  __shared__ float red[256];
  const int j = blockIdx.x;
  float a = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    float x = J[j * V + v] * (mask ? mask[j * V + v] : 1.f);
    a += fmaxf(x, 0.f);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowsum[j] = red[0];
}

__global__ void reg_normalise_kernel(const float* __restrict__ J, const float* __restrict__ mask,
                                     const float* __restrict__ rowsum, float* __restrict__ Jhat) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NH * V) return;
  const int j = idx / V;
  float x = J[idx] * (mask ? mask[idx] : 1.f);
  Jhat[idx] = fmaxf(x, 0.f) / rowsum[j];
}

// scatter the normalised columns into the packed vertex records (both range layouts) and set
// the "column is non-zero" flag
__global__ void reg_records_kernel(const float* __restrict__ Jhat, const int* __restrict__ perm,
                                   VtxRec* __restrict__ vrec, VtxRec* __restrict__ vrec_b,
                                   VtxRec* __restrict__ vrec_l) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= VP) return;
  const int v = perm[i];
  bool any = false;
  for (int j = 0; j < JH_STRIDE; j++) {
    const float r = (v >= 0 && j < NH) ? Jhat[j * V + v] : 0.f;
    any |= (r != 0.f);
    vrec[i].jh[j] = r;
    vrec_b[i].jh[j] = r;
    vrec_l[i].jh[j] = r;
  }
  const uint32_t f = any ? (1u << 24) : 0u;
  vrec[i].meta = (vrec[i].meta & ~(1u << 24)) | f;
  vrec_b[i].meta = (vrec_b[i].meta & ~(1u << 24)) | f;
  vrec_l[i].meta = (vrec_l[i].meta & ~(1u << 24)) | f;
}

// G[j][i] += sum_b sum_c gT[3j+c][b] * vT[3i+c][b]; one warp per vertex, g tile in smem
constexpr int RA_WARPS = 8;
constexpr int RA_BCH = 128;
__global__ void __launch_bounds__(RA_WARPS * 32)
reg_accumulate_kernel(const float* __restrict__ gT, const float* __restrict__ vT, int64_t B,
                      int64_t BP, const int* __restrict__ perm, float* __restrict__ G) {
  __shared__ float sg[NACC][RA_BCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * RA_WARPS + warp;
  const int vorig = i < VP ? perm[i] : -1;
  float acc[NH];
#pragma unroll
  for (int j = 0; j < NH; j++) acc[j] = 0.f;
  for (int64_t bc = 0; bc < BP; bc += RA_BCH) {
    __syncthreads();
    for (int e = threadIdx.x; e < NACC * RA_BCH; e += blockDim.x) {
      const int a = e / RA_BCH, bb = e % RA_BCH;
      sg[a][bb] = (bc + bb < B) ? gT[(int64_t)a * BP + bc + bb] : 0.f;
    }
    __syncthreads();
    if (vorig >= 0) {
#pragma unroll
      for (int q = 0; q < RA_BCH / 32; q++) {
        const int bb = lane + 32 * q;
        const float x = vT[(int64_t)(3 * i + 0) * BP + bc + bb];
        const float y = vT[(int64_t)(3 * i + 1) * BP + bc + bb];
        const float z = vT[(int64_t)(3 * i + 2) * BP + bc + bb];
#pragma unroll
        for (int j = 0; j < NH; j++)
          acc[j] += sg[j * 3 + 0][bb] * x + sg[j * 3 + 1][bb] * y + sg[j * 3 + 2][bb] * z;
      }
    }
  }
  if (vorig >= 0) {
#pragma unroll
    for (int j = 0; j < NH; j++) {
      float a = acc[j];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) G[j * V + vorig] += a;
    }
  }
}

__global__ void reg_dot_kernel(const float* __restrict__ G, const float* __restrict__ Jhat,
                               float* __restrict__ dot) {
  __shared__ float red[256];
  const int j = blockIdx.x;
  float a = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) a += G[j * V + v] * Jhat[j * V + v];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dot[j] = red[0];
}

__global__ void reg_adam_kernel(float* __restrict__ J, const float* __restrict__ mask,
                                const float* __restrict__ G, const float* __restrict__ dot,
                                const float* __restrict__ rowsum, float* __restrict__ am,
                                float* __restrict__ av, const int32_t* __restrict__ step_count, float lr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NH * V) return;
  const int j = idx / V;
  const float mk = mask ? mask[idx] : 1.f;
  const float x = J[idx] * mk;
  // d/dJ of relu(J*mask)/rowsum contracted with G (utils.py:87-92 backward)
  const float g = x > 0.f ? (G[idx] - dot[j]) / rowsum[j] * mk : 0.f;
  const int t = *step_count + 1;
  const float bc2s = (float)sqrt(1.0 - pow(0.999, (double)t));
  const float step = (float)((double)lr / (1.0 - pow(0.9, (double)t)));
  const float m = 0.9f * am[idx] + 0.1f * g;
  const float v = 0.999f * av[idx] + 0.001f * g * g;
  am[idx] = m;
  av[idx] = v;
  J[idx] = J[idx] - step * (m / (sqrtf(v) / bc2s + 1e-8f));
}

__global__ void bump_kernel(int32_t* c) { *c += 1; }

// packed blend matrices from the natural-order master: P[k][3i+c] = Pn[k][3 perm[i] + c]
__global__ void repack_blend_kernel(const float* __restrict__ Pn, const int* __restrict__ perm,
                                    float* __restrict__ P_hi, float* __restrict__ P_lo,
                                    float* __restrict__ Pt_hi, float* __restrict__ Pt_lo,
                                    float* __restrict__ P3_hi, float* __restrict__ P3_lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)KA * NP) return;
  const int k = (int)(idx / NP), n = (int)(idx % NP);
  const int v = perm[n / 3];
  const float x = v >= 0 ? Pn[(int64_t)k * (3 * V) + 3 * v + n % 3] : 0.f;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  const float hi = __uint_as_float(r);
  const float d = x - hi;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  const float lo = __uint_as_float(r);
  P_hi[idx] = hi;
  P_lo[idx] = lo;
  Pt_hi[(int64_t)n * KA + k] = hi;
  Pt_lo[(int64_t)n * KA + k] = lo;
  // per coordinate, feature-major, vertex-contiguous: the B operand of the fold GEMM (K = vertices)
  P3_hi[((int64_t)(n % 3) * KA + k) * VP + n / 3] = hi;
  P3_lo[((int64_t)(n % 3) * KA + k) * VP + n / 3] = lo;
}

// active[v] = 1 when column v of the normalised regressor has a non-zero entry
__global__ void reg_active_kernel(const float* __restrict__ Jhat, uint8_t* __restrict__ active) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  bool any = false;
  for (int j = 0; j < NH; j++) any |= (Jhat[j * V + v] != 0.f);
  active[v] = any ? 1 : 0;
}

int build_packing(JrrModel* m, const std::vector<uint8_t>& active);

int launch_regressor_normalise(JrrModel* m, const float* Jraw, const float* mask, cudaStream_t st) {
  reg_rowsum_kernel<<<NH, 256, 0, st>>>(Jraw, mask, m->rowsum);
  JRR_LAUNCH_CHECK();
  reg_normalise_kernel<<<(NH * V + 255) / 256, 256, 0, st>>>(Jraw, mask, m->rowsum, m->Jhat);
  JRR_LAUNCH_CHECK();
  for (int pass = 0; pass < m->n_pass; pass++) {      // the regressor column rides in every pass's records
    const PassTab& t = m->passes[pass];
    reg_records_kernel<<<(VP + 255) / 256, 256, 0, st>>>(m->Jhat, m->perm, t.vrec, t.vrec_b, t.vrec_l);
    JRR_LAUNCH_CHECK();
  }
  m->has_regressor = true;
  return JRR_OK;
}

int launch_regressor_accumulate(const JrrModel* m, const Workspace& w, const float* vT,
                                float* G_accum, cudaStream_t st) {
  // inactive columns cannot receive gradient (relu'(<=0) = 0): only the active prefix is accumulated
  reg_accumulate_kernel<<<(m->nv_act + RA_WARPS - 1) / RA_WARPS, RA_WARPS * 32, 0, st>>>(w.gT, vT, w.B, w.BP, m->perm, G_accum);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_regressor_apply(JrrModel* m, float* Jraw, const float* mask, const float* G, float* adam_m,
                           float* adam_v, int32_t* step_count, float lr, cudaStream_t st) {
  reg_dot_kernel<<<NH, 256, 0, st>>>(G, m->Jhat, m->regdot);
  JRR_LAUNCH_CHECK();
  reg_adam_kernel<<<(NH * V + 255) / 256, 256, 0, st>>>(Jraw, mask, G, m->regdot, m->rowsum, adam_m,
                                                       adam_v, step_count, lr);
  JRR_LAUNCH_CHECK();
  bump_kernel<<<1, 1, 0, st>>>(step_count);
  JRR_LAUNCH_CHECK();
  if (int rc = launch_regressor_normalise(m, Jraw, mask, st)) return rc;
  return m->folded ? launch_fold(m, st) : JRR_OK;
}

// --------------------------------------------------------------------- regressor refit through the folded operator
// The refit gradient G_iv = d loss / d Jhat_iv = sum_b g~_b,i . V_b,v needs the skinned vertices of every frame in its
// literal form.  But V is linear in the blend features with the joint transforms as the only pose-dependent factors -- the
// same structure fold_kernel exploits -- so G is the ADJOINT of the fold applied to small per-batch sums:
//   dT_ji[c][f] = sum_b (A_j^R(b)^T g~_b,i)[c] feat_b[f]      (= dQ^T . feat: one [1224 x B] x [B x 224] GEMM, K = batch)
//   dc_ji       = sum_b g~_b,i . A_j^t(b)                      (24 x 17 sums, per-CTA partials of the seed kernel)
//   G_iv        = sum_j w_vj ( <dT_ji, P_v> + dc_ji )          (U = dT . P^T: one [408 x 672] x [672 x 6912] GEMM + a gather)
// with dQ, g~ from folded_seed_kernel (weight 1).  No per-vertex work per frame, no vertices in memory: 4096 frames cost two
// small tensor-core GEMMs instead of a blend GEMM, a 339 MB vertex store and its re-read.  (optimize.py:300-309)
constexpr int UNF_M = 512;                 // (j,i) rows of the unfold GEMM, 408 padded to the 128-row tile
constexpr int UNF_K = 3 * KA;              // 672: (coordinate, feature)
constexpr int DT_KSPLIT = 8;

__global__ void dT_finish_kernel(const float* __restrict__ part, float* __restrict__ hi, float* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= UNF_M * UNF_K) return;
  float a = 0.f;
  if (idx < FOLD_N * KA)
    for (int s = 0; s < DT_KSPLIT; s++) a += part[(int64_t)s * FOLD_NP * KA + idx];      // fixed order
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  const float h = __uint_as_float(r);
  const float d = a - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  hi[idx] = h;
  lo[idx] = __uint_as_float(r);
}

__global__ void dc_reduce_kernel(const float* __restrict__ part, int nblk, float* __restrict__ dc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NJ * NH) return;
  float a = 0.f;
  for (int b = 0; b < nblk; b++) a += part[(int64_t)b * (NJ * NH) + e];
  dc[e] = a;
}

// G[i][perm[p]] += sum_slots w_p,slot ( U[(j_slot, i)][p] + dc[j_slot][i] ), packed vertex p (thread), regressor row i (blockIdx.y)
__global__ void __launch_bounds__(256)
unfold_gather_kernel(const VtxRec* __restrict__ vrec, const int* __restrict__ perm, const float* __restrict__ U,
                     const float* __restrict__ dc, int nv, float* __restrict__ G) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (p >= nv) return;
  const int v = perm[p];
  if (v < 0) return;
  const uint32_t meta = vrec[p].meta;
  float a = 0.f;
#pragma unroll
  for (int s4 = 0; s4 < 4; s4++) {
    const float wgt = vrec[p].w[s4];
    if (wgt != 0.f) {
      const int j = (meta >> (5 * s4)) & 31u;
      a = fmaf(wgt, U[(int64_t)(j * NH + i) * VP + p] + dc[j * NH + i], a);
    }
  }
  G[i * V + v] += a;
}

int regressor_accumulate_folded(const JrrModel* m, Workspace& w, const float* gt_mm, int64_t B_logical, float* G_accum,
                                cudaStream_t st) {
  const int64_t BP = w.BP;
  // scratch in the (idle) blend-gradient buffers: [dQ | dQ^T | feat^T | ...] hi in dvp_hi, lo in dvp_lo
  float* dQT_hi = w.dvp_hi + BP * FOLD_NP;
  float* dQT_lo = w.dvp_lo + BP * FOLD_NP;
  float* fT_hi = w.dvp_hi + 2 * BP * FOLD_NP;
  float* fT_lo = w.dvp_lo + 2 * BP * FOLD_NP;
  float* dT_part = w.dvp_hi + 2 * BP * FOLD_NP + BP * KA;                  // [DT_KSPLIT][1280][224]
  float* u = w.dvp_lo + 2 * BP * FOLD_NP + BP * KA;
  float* dTs_hi = u; u += UNF_M * UNF_K;
  float* dTs_lo = u; u += UNF_M * UNF_K;
  float* U = u;      u += (int64_t)UNF_M * VP;
  float* dc_part = u; u += (BP / 32) * (NJ * NH);
  float* dc = u;      u += 512;
  if ((u - w.dvp_lo) > (int64_t)NP * BP || (dT_part + (int64_t)DT_KSPLIT * FOLD_NP * KA - w.dvp_hi) > (int64_t)NP * BP)
    return fail(JRR_ERR_WORKSPACE, "folded refit scratch does not fit the workspace");
  // per-frame joints, loss, seed g~ (weight 1), dQ = A^R^T g~ and the dc partials
  if (int rc = launch_folded_seed(m, w, gt_mm, B_logical, 1.f, nullptr, Proj2D{}, st, dc_part)) return rc;
  if (int rc = launch_transpose(w.dvp_hi, BP, FOLD_NP, dQT_hi, st)) return rc;
  if (int rc = launch_transpose(w.dvp_lo, BP, FOLD_NP, dQT_lo, st)) return rc;
  if (int rc = launch_transpose(w.feat_hi, BP, KA, fT_hi, st)) return rc;
  if (int rc = launch_transpose(w.feat_lo, BP, KA, fT_lo, st)) return rc;
  {
    GemmDesc g{};                       // dT[(j,i,c)][f] = sum_b dQ^T[(j,i,c)][b] feat^T[f][b]
    g.A_hi = dQT_hi; g.A_lo = dQT_lo; g.lda = BP;
    g.B_hi = fT_hi; g.B_lo = fT_lo; g.ldb = BP;
    g.M = FOLD_NP; g.N = KA; g.K = BP / DT_KSPLIT; g.ksplit = DT_KSPLIT; g.epi = EPI_STORE_SPLITK;
    g.out0 = dT_part; g.ldo = KA;
    if (int rc = launch_gemm(m, g, st)) return rc;
  }
  dT_finish_kernel<<<(UNF_M * UNF_K + 255) / 256, 256, 0, st>>>(dT_part, dTs_hi, dTs_lo);
  JRR_LAUNCH_CHECK();
  dc_reduce_kernel<<<(NJ * NH + 127) / 128, 128, 0, st>>>(dc_part, (int)(BP / 32), dc);
  JRR_LAUNCH_CHECK();
  const int nv = (int)round_up(m->nv_act, 128);
  {
    GemmDesc g{};                       // U[(j,i)][p] = < dT_ji, P_p >  over (c,f) = 672
    g.A_hi = dTs_hi; g.A_lo = dTs_lo; g.lda = UNF_K;
    g.B_hi = m->Pt_hi; g.B_lo = m->Pt_lo; g.ldb = UNF_K;
    g.M = UNF_M; g.N = nv; g.K = UNF_K; g.ksplit = 1; g.epi = EPI_STORE_SPLITK;
    g.out0 = U; g.ldo = VP;
    if (int rc = launch_gemm(m, g, st)) return rc;
  }
  for (int pass = 0; pass < m->n_pass; pass++) {
    unfold_gather_kernel<<<dim3((unsigned)((m->nv_act + 255) / 256), NH), 256, 0, st>>>(m->passes[pass].vrec, m->perm, U, dc, m->nv_act, G_accum);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

// --------------------------------------------------------------------- folded loss-path operator
// T[(j,i,c)][k] = sum_v Jhat_iv w_vj P[3v+c][k],  c_ji = sum_v Jhat_iv w_vj   (see folded_seed_kernel).
// grid (17 regressor rows in groups of FOLD_R, 4 = three coordinates + the homogeneous one, FOLD_CH vertex chunks), thread = k;
// double accumulators in shared memory, fixed summation order (chunk partials reduced in order).
// 12 chunks = the forward record ranges (a range start reloads all four slots).  Measured alternatives: 9 chunks of 768
// vertices (612 CTAs = ONE wave at five CTAs per SM instead of a wave and a 76-CTA tail) 301 vs 270 us -- the time follows the
// vertices a CTA walks, i.e. the kernel pulls 842 MB of blend-matrix rows through L2 at ~3 TB/s with 35 warps per SM
constexpr int FOLD_CH = VP / VS_F;

// wj[i][p][slot] = w_p,slot * Jhat_i,p in double (exact: 24 + 24 significand bits), once per regressor version -- the fold
// kernel's 224 threads per CTA would otherwise each redo these conversions and products for every vertex
// (wrec: the record table whose SLOT ORDER the fold kernel walks -- slots are re-assigned at every range start, so the tables of
// different range sizes order a vertex's weights differently; jrec: any table that carries the regressor columns)
__global__ void fold_prep_kernel(const VtxRec* __restrict__ wrec, const VtxRec* __restrict__ jrec, double* __restrict__ wj) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NH * VP) return;
  const int i = idx / VP, p = idx % VP;
  const double jh = (double)jrec[p].jh[i];
#pragma unroll
  for (int s4 = 0; s4 < 4; s4++) wj[(int64_t)idx * 4 + s4] = (double)wrec[p].w[s4] * jh;
}

// regressor rows per CTA.  Measured: 3 rows per CTA (the blend-matrix rows re-read 6 times instead of 17, but 129 KB of
// accumulators = one 7-warp CTA per SM) is SLOWER than 1 (332 vs 270 us): seven warps per SM cannot keep enough loads in
// flight; five co-resident CTAs can.
constexpr int FOLD_R = 1;
constexpr int FOLD_SMEM = FOLD_R * NJ * KA * (int)sizeof(double);   // 43 KB

__global__ void __launch_bounds__(KA)
fold_kernel(const VtxRec* __restrict__ vrec, const double* __restrict__ wj, const float* __restrict__ Pt_hi,
            const float* __restrict__ Pt_lo, double* __restrict__ part, int accumulate) {
  // Packed vertices are sorted by joint set, so a record slot keeps its joint over long runs: each thread keeps the four
  // slots' running sums in REGISTERS (independent double chains) and adds them to the per-joint shared-memory
  // accumulators only when a slot's joint changes (the records' reload bits: ~190 times over the whole model) -- the
  // first version updated the shared accumulators for every (vertex, slot), one dependent shared-memory round trip each
  // (0.46 ms).  Same fp64 arithmetic; groups of four vertices without a slot change take a branch-free path.
  // A CTA folds FOLD_R regressor rows at once (see FOLD_R).
  extern __shared__ double fold_acc[];          // [FOLD_R][NJ * KA]
  const int i0 = blockIdx.x * FOLD_R, c = blockIdx.y, ch = blockIdx.z, k = threadIdx.x;
  for (int e = k; e < FOLD_R * NJ * KA; e += KA) fold_acc[e] = 0.0;
  __syncthreads();                 // (each thread only ever touches column k, the barrier is for the zeroing loop)
  double run[FOLD_R][4];
#pragma unroll
  for (int r = 0; r < FOLD_R; r++)
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) run[r][s4] = 0.0;
  int jcur[4] = {0, 0, 0, 0};
  const int p0 = ch * VS_F;
  const double2* wj2[FOLD_R];
#pragma unroll
  for (int r = 0; r < FOLD_R; r++)
    wj2[r] = reinterpret_cast<const double2*>(wj + ((int64_t)min(i0 + r, NH - 1) * VP + p0) * 4);
#pragma unroll 1
  for (int q0 = 0; q0 < VS_F; q0 += 4) {
    uint32_t meta[4], many = 0;
    double p[4];
    double2 wa[FOLD_R][4], wb[FOLD_R][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {       // four vertices' loads in flight together
      const int pi = p0 + q0 + u;
      meta[u] = vrec[pi].meta;
      many |= meta[u];
#pragma unroll
      for (int r = 0; r < FOLD_R; r++) {
        wa[r][u] = __ldg(wj2[r] + (q0 + u) * 2);
        wb[r][u] = __ldg(wj2[r] + (q0 + u) * 2 + 1);
      }
      if (c < 3) p[u] = (double)__ldg(Pt_hi + (int64_t)(3 * pi + c) * KA + k) + (double)__ldg(Pt_lo + (int64_t)(3 * pi + c) * KA + k);
      else p[u] = k == 0 ? 1.0 : 0.0;
    }
    if (!((many >> 20) & 0xFu)) {       // no slot changes inside the group (uniform)
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int r = 0; r < FOLD_R; r++) {
          run[r][0] = fma(wa[r][u].x, p[u], run[r][0]);
          run[r][1] = fma(wa[r][u].y, p[u], run[r][1]);
          run[r][2] = fma(wb[r][u].x, p[u], run[r][2]);
          run[r][3] = fma(wb[r][u].y, p[u], run[r][3]);
        }
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
          if ((meta[u] >> (20 + s4)) & 1u) {          // the slot changes its joint before this vertex (uniform)
#pragma unroll
            for (int r = 0; r < FOLD_R; r++) {
              fold_acc[(r * NJ + jcur[s4]) * KA + k] += run[r][s4];
              run[r][s4] = 0.0;
            }
            jcur[s4] = (meta[u] >> (5 * s4)) & 31u;
          }
        }
#pragma unroll
        for (int r = 0; r < FOLD_R; r++) {
          run[r][0] = fma(wa[r][u].x, p[u], run[r][0]);
          run[r][1] = fma(wa[r][u].y, p[u], run[r][1]);
          run[r][2] = fma(wb[r][u].x, p[u], run[r][2]);
          run[r][3] = fma(wb[r][u].y, p[u], run[r][3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < FOLD_R; r++)
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) fold_acc[(r * NJ + jcur[s4]) * KA + k] += run[r][s4];
#pragma unroll
  for (int r = 0; r < FOLD_R; r++) {
    if (i0 + r >= NH) break;
    double* out = part + ((int64_t)(ch * NH + i0 + r) * 4 + c) * (NJ * KA);
    for (int j = 0; j < NJ; j++)
      out[j * KA + k] = accumulate ? out[j * KA + k] + fold_acc[(r * NJ + j) * KA + k] : fold_acc[(r * NJ + j) * KA + k];
  }
}

__global__ void fold_finish_kernel(const double* __restrict__ part, float* __restrict__ T_hi, float* __restrict__ T_lo,
                                   float* __restrict__ Tt_hi, float* __restrict__ Tt_lo, float* __restrict__ Tc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)FOLD_N * KA) {
    const int n = (int)(idx / KA), k = (int)(idx % KA);
    const int c = n % 3, i = (n / 3) % NH, j = n / (3 * NH);
    double a = 0.0;
    for (int ch = 0; ch < FOLD_CH; ch++) a += part[((int64_t)(ch * NH + i) * 4 + c) * (NJ * KA) + j * KA + k];
    const float x = (float)a;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    const float hi = __uint_as_float(r);
    const float d = x - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
    const float lo = __uint_as_float(r);
    T_hi[idx] = hi;
    T_lo[idx] = lo;
    Tt_hi[(int64_t)k * FOLD_NP + n] = hi;
    Tt_lo[(int64_t)k * FOLD_NP + n] = lo;
  } else if (idx < (int64_t)FOLD_N * KA + NJ * NH) {
    const int e = (int)(idx - (int64_t)FOLD_N * KA), j = e / NH, i = e % NH;
    double a = 0.0;
    for (int ch = 0; ch < FOLD_CH; ch++) a += part[((int64_t)(ch * NH + i) * 4 + 3) * (NJ * KA) + j * KA];
    Tc[e] = (float)a;
  }
}

// ---- the fold on register run-sums with NO shared accumulators (JRR_FOLD_RUNS=1; measured SLOWER, so not the default) --------
// fold_kernel keeps one regressor row per CTA because its per-joint accumulators take 43 KB of shared memory per row, so the
// blend matrix travels through L2 once per row: 17 x 2 x 18.6 MB x 4/3 = 842 MB per fold at ~3 TB/s = 0.27 ms.  Here a CTA
// walks ONE 192-vertex record range for 8-9 regressor rows at once: thread = blend feature k, run[row][slot] in registers
// (a record slot keeps its joint over long runs of the joint-sorted packing), and when a slot changes its joint the run sums
// leave as a FLUSH EVENT -- the very events, in the very order, of the skinning backward's joint-transform gradients
// (range_flush_base_s / flush_ptr_s / flush_idx_s, built with the records) -- so a gather over each joint's event list
// finishes the fold in a fixed order.  The blend matrix is read twice per fold instead of 17 times.
// Measured on a B200: 495 us against 270 us for fold_kernel (same 421 M DFMAs, 14 instead of 35 warps per SM): the fold is
// paced by the double-precision FMAs, ~6 SM cycles per warp-DFMA in fold_kernel and ~10 here -- not by the L2 traffic this
// variant removes.  Kept behind the switch as the record of that experiment; same results (parity suite green with it).
constexpr int FR_ROWS = 9;       // regressor rows per CTA: [0, 9) and [9, 17)

__global__ void __launch_bounds__(KA)
fold_runs_kernel(const VtxRec* __restrict__ vrec, const double* __restrict__ wj, const float* __restrict__ Pt_hi,
                 const float* __restrict__ Pt_lo, const int* __restrict__ range_base, double* __restrict__ part_ev) {
  const int c = blockIdx.x, rg = blockIdx.y, i0 = blockIdx.z * FR_ROWS, k = threadIdx.x;
  const int nrow = min(FR_ROWS, NH - i0);
  double run[FR_ROWS][4];
#pragma unroll
  for (int r = 0; r < FR_ROWS; r++)
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) run[r][s4] = 0.0;
  int ev = range_base[rg];
  const int p0 = rg * VS_S;
  auto flush = [&](int s4sel) {
    // run sums of slot s4sel (compile-time after unrolling) -> event ev, rows i0 .. i0 + nrow
#pragma unroll
    for (int r = 0; r < FR_ROWS; r++) {
      if (r < nrow) {
        const double v = s4sel == 0 ? run[r][0] : s4sel == 1 ? run[r][1] : s4sel == 2 ? run[r][2] : run[r][3];
        part_ev[(((int64_t)ev * NH + i0 + r) * 4 + c) * KA + k] = v;
      }
      if (s4sel == 0) run[r][0] = 0.0; else if (s4sel == 1) run[r][1] = 0.0; else if (s4sel == 2) run[r][2] = 0.0; else run[r][3] = 0.0;
    }
    ev++;
  };
#pragma unroll 2
  for (int q = 0; q < VS_S; q++) {
    const int pi = p0 + q;
    const uint32_t meta = __ldg(&vrec[pi].meta);
    const double p = c < 3 ? (double)__ldg(Pt_hi + (int64_t)(3 * pi + c) * KA + k) + (double)__ldg(Pt_lo + (int64_t)(3 * pi + c) * KA + k)
                           : (k == 0 ? 1.0 : 0.0);
    if (((meta >> 20) & 0xFu) && !((meta >> 25) & 1u)) {       // a slot changes its joint before this vertex (uniform)
      if ((meta >> 20) & 1u) flush(0);
      if ((meta >> 21) & 1u) flush(1);
      if ((meta >> 22) & 1u) flush(2);
      if ((meta >> 23) & 1u) flush(3);
    }
#pragma unroll
    for (int r = 0; r < FR_ROWS; r++) {
      if (r < nrow) {
        const double2* w2 = reinterpret_cast<const double2*>(wj + ((int64_t)(i0 + r) * VP + pi) * 4);
        const double2 wa = __ldg(w2), wb = __ldg(w2 + 1);
        run[r][0] = fma(wa.x, p, run[r][0]);
        run[r][1] = fma(wa.y, p, run[r][1]);
        run[r][2] = fma(wb.x, p, run[r][2]);
        run[r][3] = fma(wb.y, p, run[r][3]);
      }
    }
  }
  flush(0); flush(1); flush(2); flush(3);       // the range's last four events
}

// T[(j,i,c)][k] (and c_ji) = the sum of joint j's flush events, in the order of its list; later skinning passes add
__global__ void fold_gather_kernel(const double* __restrict__ part_ev, const int* __restrict__ fptr, const int* __restrict__ fidx,
                                   double* __restrict__ Tacc, int accumulate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)FOLD_N * KA + NJ * NH) return;
  int j, i, c, k;
  if (idx < (int64_t)FOLD_N * KA) {
    const int n = (int)(idx / KA);
    k = (int)(idx % KA); c = n % 3; i = (n / 3) % NH; j = n / (3 * NH);
  } else {
    const int e = (int)(idx - (int64_t)FOLD_N * KA);
    j = e / NH; i = e % NH; c = 3; k = 0;
  }
  double a = accumulate ? Tacc[idx] : 0.0;
  for (int q = fptr[j]; q < fptr[j + 1]; q++) a += part_ev[(((int64_t)fidx[q] * NH + i) * 4 + c) * KA + k];
  Tacc[idx] = a;
}

__global__ void fold_split_kernel(const double* __restrict__ Tacc, float* __restrict__ T_hi, float* __restrict__ T_lo,
                                  float* __restrict__ Tt_hi, float* __restrict__ Tt_lo, float* __restrict__ Tc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)FOLD_N * KA) {
    const int n = (int)(idx / KA), k = (int)(idx % KA);
    const float x = (float)Tacc[idx];
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    const float hi = __uint_as_float(r);
    const float d = x - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
    const float lo = __uint_as_float(r);
    T_hi[idx] = hi;
    T_lo[idx] = lo;
    Tt_hi[(int64_t)k * FOLD_NP + n] = hi;
    Tt_lo[(int64_t)k * FOLD_NP + n] = lo;
  } else if (idx < (int64_t)FOLD_N * KA + NJ * NH) {
    Tc[idx - (int64_t)FOLD_N * KA] = (float)Tacc[idx];
  }
}

static int launch_fold_runs(JrrModel* m, cudaStream_t st) {
  int max_ev = 0;
  for (int pass = 0; pass < m->n_pass; pass++) max_ev = std::max(max_ev, m->passes[pass].n_flush_s);
  if (max_ev > m->fold_ev_cap) {
    // (only when the packing changed: jrr_set_regressor with a new support re-packs the vertices on the host anyway)
    JRR_CUDA(cudaStreamSynchronize(st));
    if (m->fold_ev) JRR_CUDA(cudaFree(m->fold_ev));
    m->fold_ev = nullptr;
    m->fold_ev_cap = max_ev + 64;
    JRR_CUDA(cudaMalloc((void**)&m->fold_ev, (size_t)m->fold_ev_cap * NH * 4 * KA * sizeof(double)));
  }
  const int64_t n = (int64_t)FOLD_N * KA + NJ * NH;
  for (int pass = 0; pass < m->n_pass; pass++) {
    const PassTab& t = m->passes[pass];
    fold_prep_kernel<<<(NH * VP + 255) / 256, 256, 0, st>>>(t.vrec_s, t.vrec, m->fold_wj);
    JRR_LAUNCH_CHECK();
    fold_runs_kernel<<<dim3(4, NSPLIT_S, (NH + FR_ROWS - 1) / FR_ROWS), KA, 0, st>>>(t.vrec_s, m->fold_wj, m->Pt_hi, m->Pt_lo,
                                                                                  t.range_flush_base_s, m->fold_ev);
    JRR_LAUNCH_CHECK();
    fold_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->fold_ev, t.flush_ptr_s, t.flush_idx_s, m->fold_acc, pass > 0 ? 1 : 0);
    JRR_LAUNCH_CHECK();
  }
  fold_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->fold_acc, m->T_hi, m->T_lo, m->Tt_hi, m->Tt_lo, m->Tc);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// ---- the fold as three tensor-core GEMMs (default; JRR_FOLD_GEMM=0 selects fold_kernel) ------------------------------------
// T[(j,i,c)][k] = sum_p WJ[(j,i)][p] P3[c][k][p] with WJ[(j,i)][p] = w_pj Jhat_ip: per coordinate c one [512 x 6912] x
// [6912 x 224] GEMM over the packed vertices -- the mirror image of the refit's unfold GEMM above.  fold_kernel does the same
// sum with 421 M double-precision FMAs (0.27 ms: the DFMAs pace it, three restructurings measured); the tensor cores do it in
// 3xTF32 over 36 K splits of 192 vertices (truncation grows with the length of a split's accumulation chain: 1e-6 relative at
// 192-224, the blend GEMM's own figure), and the splits are summed in double in a fixed order.  WJ is 408 x 6912 with
// 4 x 17 non-zeros per vertex: dense on purpose -- 11.6 GFLOP of 3xTF32 is 20 us of tensor time.
constexpr int FG_M = 512;                  // (j, i) rows, 408 padded to the 128-row tile
constexpr int FG_KSPLIT = 36;              // 6912 vertices = 36 x 192

// (one launch per skinning pass, in order: thread (p, i) owns WJ[(., i)][p], so the passes of a vertex add up without races)
__global__ void fold_wj_kernel(const VtxRec* __restrict__ vrec, float* __restrict__ WJ) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NH * VP) return;
  const int i = idx / VP, p = idx % VP;
  const float jh = vrec[p].jh[i];
  const uint32_t meta = vrec[p].meta;
#pragma unroll
  for (int s4 = 0; s4 < 4; s4++) {
    const float wgt = vrec[p].w[s4];
    if (wgt != 0.f && jh != 0.f) WJ[(int64_t)(((meta >> (5 * s4)) & 31u) * NH + i) * VP + p] += wgt * jh;
  }
}

__global__ void fold_wj_split_kernel(const float* __restrict__ WJ, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)FG_M * VP) return;
  const float x = WJ[idx];
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  const float h = __uint_as_float(r);
  const float d = x - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  hi[idx] = h;
  lo[idx] = __uint_as_float(r);
}

// c_ji = sum_p WJ[(j,i)][p]: one warp per row, lane-strided partial sums in double + a butterfly (fixed order)
__global__ void fold_rowsum_kernel(const float* __restrict__ WJ, float* __restrict__ Tc) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= NJ * NH) return;
  double a = 0.0;
  for (int p = lane; p < VP; p += 32) a += (double)WJ[(int64_t)row * VP + p];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) Tc[row] = (float)a;
}

// K splits summed in double, rounded once, split into the tf32 pair, both majors
__global__ void fold_gemm_finish_kernel(const float* __restrict__ part, float* __restrict__ T_hi, float* __restrict__ T_lo,
                                        float* __restrict__ Tt_hi, float* __restrict__ Tt_lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)FOLD_N * KA) return;
  const int n = (int)(idx / KA), k = (int)(idx % KA);
  const int c = n % 3, ji = n / 3;           // n = (j * 17 + i) * 3 + c
  const float* src = part + (int64_t)c * FG_KSPLIT * FG_M * KA + (int64_t)ji * KA + k;
  double a = 0.0;
  for (int sp = 0; sp < FG_KSPLIT; sp++) a += (double)src[(int64_t)sp * FG_M * KA];
  const float x = (float)a;
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  const float hi = __uint_as_float(r);
  const float d = x - hi;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  const float lo = __uint_as_float(r);
  T_hi[idx] = hi;
  T_lo[idx] = lo;
  Tt_hi[(int64_t)k * FOLD_NP + n] = hi;
  Tt_lo[(int64_t)k * FOLD_NP + n] = lo;
}

static int launch_fold_gemm(JrrModel* m, cudaStream_t st) {
  if (!m->fg_wj) {
    if (int rc = dalloc(m, &m->fg_wj, (size_t)FG_M * VP)) return rc;                    // zero-filled: the padding rows stay zero
    if (int rc = dalloc(m, &m->fg_wj_hi, (size_t)FG_M * VP, false)) return rc;
    if (int rc = dalloc(m, &m->fg_wj_lo, (size_t)FG_M * VP, false)) return rc;
    if (int rc = dalloc(m, &m->fg_part, (size_t)3 * FG_KSPLIT * FG_M * KA, false)) return rc;
  }
  JRR_CUDA(cudaMemsetAsync(m->fg_wj, 0, sizeof(float) * (size_t)NJ * NH * VP, st));
  for (int pass = 0; pass < m->n_pass; pass++) {
    fold_wj_kernel<<<(NH * VP + 255) / 256, 256, 0, st>>>(m->passes[pass].vrec, m->fg_wj);
    JRR_LAUNCH_CHECK();
  }
  fold_wj_split_kernel<<<(unsigned)(((int64_t)FG_M * VP + 255) / 256), 256, 0, st>>>(m->fg_wj, m->fg_wj_hi, m->fg_wj_lo);
  JRR_LAUNCH_CHECK();
  fold_rowsum_kernel<<<(NJ * NH * 32 + 255) / 256, 256, 0, st>>>(m->fg_wj, m->Tc);
  JRR_LAUNCH_CHECK();
  for (int c = 0; c < 3; c++) {
    GemmDesc g{};                       // part_c[split][(j,i)][k] = sum_{p in split} WJ[(j,i)][p] P3[c][k][p]
    g.A_hi = m->fg_wj_hi; g.A_lo = m->fg_wj_lo; g.lda = VP;
    g.B_hi = m->P3_hi + (size_t)c * KA * VP; g.B_lo = m->P3_lo + (size_t)c * KA * VP; g.ldb = VP;
    g.M = FG_M; g.N = KA; g.K = VP / FG_KSPLIT; g.ksplit = FG_KSPLIT; g.epi = EPI_STORE_SPLITK;
    g.out0 = m->fg_part + (size_t)c * FG_KSPLIT * FG_M * KA; g.ldo = KA;
    if (int rc = launch_gemm(m, g, st)) return rc;
  }
  fold_gemm_finish_kernel<<<(unsigned)(((int64_t)FOLD_N * KA + 255) / 256), 256, 0, st>>>(m->fg_part, m->T_hi, m->T_lo, m->Tt_hi, m->Tt_lo);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_fold(JrrModel* m, cudaStream_t st) {
  if (!m->T_hi) {
    if (int rc = dalloc(m, &m->T_hi, (size_t)FOLD_NP * KA)) return rc;      // zero-filled: the padding rows stay zero
    if (int rc = dalloc(m, &m->T_lo, (size_t)FOLD_NP * KA)) return rc;
    if (int rc = dalloc(m, &m->Tt_hi, (size_t)FOLD_NP * KA)) return rc;
    if (int rc = dalloc(m, &m->Tt_lo, (size_t)FOLD_NP * KA)) return rc;
    if (int rc = dalloc(m, &m->Tc, (size_t)NJ * NH)) return rc;
    if (int rc = dalloc(m, &m->fold_part, (size_t)FOLD_CH * NH * 4 * NJ * KA)) return rc;
    if (int rc = dalloc(m, &m->fold_wj, (size_t)NH * VP * 4)) return rc;
    if (int rc = dalloc(m, &m->fold_acc, (size_t)FOLD_N * KA + NJ * NH)) return rc;
  }
  static const bool runs = [] { const char* e = getenv("JRR_FOLD_RUNS"); return e && e[0] == '1'; }();
  if (runs) return launch_fold_runs(m, st);
  static const bool fgemm = [] { const char* e = getenv("JRR_FOLD_GEMM"); return !(e && e[0] == '0'); }();
  if (fgemm) return launch_fold_gemm(m, st);
  for (int pass = 0; pass < m->n_pass; pass++) {      // linear in the skinning weights: passes add up in the fp64 partials
    const PassTab& t = m->passes[pass];
    fold_prep_kernel<<<(NH * VP + 255) / 256, 256, 0, st>>>(t.vrec, t.vrec, m->fold_wj);
    JRR_LAUNCH_CHECK();
    JRR_CUDA(cudaFuncSetAttribute(fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FOLD_SMEM));
    fold_kernel<<<dim3((NH + FOLD_R - 1) / FOLD_R, 4, FOLD_CH), KA, FOLD_SMEM, st>>>(t.vrec, m->fold_wj, m->Pt_hi, m->Pt_lo, m->fold_part, pass > 0 ? 1 : 0);
    JRR_LAUNCH_CHECK();
  }
  const int64_t n = (int64_t)FOLD_N * KA + NJ * NH;
  fold_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->fold_part, m->T_hi, m->T_lo, m->Tt_hi, m->Tt_lo, m->Tc);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// --------------------------------------------------------------------- critic load kernels
__global__ void split_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ hi,
                             float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r;
  float x = src[i];
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  float h = __uint_as_float(r);
  hi[i] = h;
  float d = x - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(d));
  lo[i] = __uint_as_float(r);
}

__global__ void split_transpose_kernel(const float* __restrict__ src, int rows, int cols,
                                       float* __restrict__ hi, float* __restrict__ lo,
                                       const float* __restrict__ row_scale = nullptr) {
  // src [rows][cols] (row r optionally scaled by row_scale[r]) -> hi/lo [cols][rows]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int c = (int)(i / rows), r = (int)(i % rows);
  uint32_t q;
  float x = src[(int64_t)r * cols + c];
  if (row_scale != nullptr) x *= row_scale[r];
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(x));
  float h = __uint_as_float(q);
  hi[i] = h;
  float d = x - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(d));
  lo[i] = __uint_as_float(q);
}

// K-range size of the fused backward: the largest of 768/384/192 vertices that still yields >= 9
// ranges over the active prefix (enough CTAs per 256-pose block), 192 for very sparse regressors
static int loss_range_size(int n_active) {
  for (int vs : {768, 384, 192})
    if (round_up(n_active, vs) / vs >= 9) return vs;
  return 192;
}

// (Re)builds everything that depends on the packed vertex order.  Vertices with a non-zero
// regressor column ("active") come first, so the loss path only walks the first nv_act packed
// vertices; inside each class vertices are sorted by joint set (long runs for the skinning
// kernels).  With a dense regressor every vertex is active and nothing is skipped.
int build_packing(JrrModel* m, const std::vector<uint8_t>& active) {
  auto lowest_first = [](uint32_t a, uint32_t b) {
    const uint32_t diff = a ^ b;
    const uint32_t low = diff & (~diff + 1u);
    return (a & low) != 0;
  };
  std::vector<int> order(V);
  for (int v = 0; v < V; v++) order[v] = v;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    if (active[x] != active[y]) return active[x] > active[y];
    if (m->h_key[x] == m->h_key[y]) return false;
    return lowest_first(m->h_key[x], m->h_key[y]);
  });
  int n_active = 0;
  for (int v = 0; v < V; v++) n_active += active[v] ? 1 : 0;
  if (n_active == 0) return fail(JRR_ERR_INVALID, "the normalised regressor has no non-zero column");
  std::vector<int> perm(VP, -1);
  for (int i = 0; i < V; i++) perm[i] = order[i];
  m->vs_l = loss_range_size(n_active);
  m->nv_act = (int)std::min<int64_t>(VP, round_up(n_active, m->vs_l));
  m->nsplit_act = m->nv_act / m->vs_l;
  m->packed_active = active;

  std::vector<int> vx_src;
  std::vector<float> vx_coef;
  std::vector<int> xptr(VP, 0), xcnt(VP, 0);
  for (int i = 0; i < VP; i++) {
    xptr[i] = (int)vx_src.size();
    if (perm[i] >= 0)
      for (auto& e : m->h_vx[perm[i]]) { vx_src.push_back(e.first); vx_coef.push_back(e.second); }
    xcnt[i] = (int)vx_src.size() - xptr[i];
  }
  auto build = [&](int pass, int VS, std::vector<VtxRec>& rec, std::vector<int>* flush_joint, std::vector<int>* range_base) {
    rec.assign(VP, VtxRec{});
    int cur[4] = {0, 0, 0, 0};
    for (int i = 0; i < VP; i++) {
      const bool first = (i % VS) == 0;
      if (first && range_base) (*range_base)[i / VS] = (int)flush_joint->size();
      // this pass's (up to four) weights of the vertex: entries [4*pass, 4*pass + 4) of its list
      std::vector<std::pair<int, float>> nz;
      if (perm[i] >= 0) {
        const auto& all = m->h_lbs[perm[i]];
        for (size_t e = (size_t)4 * pass; e < all.size() && e < (size_t)4 * pass + 4; e++) nz.push_back(all[e]);
      }
      int nj[4];
      float nw[4] = {0.f, 0.f, 0.f, 0.f};
      bool used[4] = {false, false, false, false};
      for (int k = 0; k < 4; k++) nj[k] = first ? 0 : cur[k];
      std::vector<std::pair<int, float>> rest;
      for (auto& e : nz) {
        int hit = -1;
        if (!first)
          for (int k = 0; k < 4; k++)
            if (!used[k] && cur[k] == e.first) { hit = k; break; }
        if (hit >= 0) { used[hit] = true; nw[hit] = e.second; }
        else rest.push_back(e);
      }
      for (auto& e : rest)
        for (int k = 0; k < 4; k++)
          if (!used[k]) { used[k] = true; nj[k] = e.first; nw[k] = e.second; break; }
      uint32_t meta = 0;
      for (int k = 0; k < 4; k++) {
        meta |= (uint32_t)nj[k] << (5 * k);
        const bool reload = first || nj[k] != cur[k];
        if (reload) {
          meta |= 1u << (20 + k);
          if (!first && flush_joint) flush_joint->push_back(cur[k]);
        }
        cur[k] = nj[k];
      }
      if (first) meta |= 1u << 25;
      if (xcnt[i] > 0 && pass == 0) meta |= 1u << 26;
      rec[i].meta = meta;
      for (int k = 0; k < 4; k++) rec[i].w[k] = nw[k];
      rec[i].xptr = xptr[i];
      rec[i].xcnt = pass == 0 ? xcnt[i] : 0;
      if (flush_joint && (i % VS) == VS - 1)
        for (int k = 0; k < 4; k++) flush_joint->push_back(cur[k]);
    }
  };
  // CSR joint -> flush ids (ids ascend inside a joint's list)
  auto csr = [](const std::vector<int>& fj, std::vector<int>& fptr, std::vector<int>& fidx) {
    fptr.assign(NJ + 1, 0);
    fidx.assign(fj.size(), 0);
    for (int f : fj) fptr[f + 1]++;
    for (int j = 0; j < NJ; j++) fptr[j + 1] += fptr[j];
    std::vector<int> fill(fptr.begin(), fptr.end() - 1);
    for (int f = 0; f < (int)fj.size(); f++) fidx[fill[fj[f]]++] = f;
  };
  JRR_CUDA(cudaMemcpy(m->perm, perm.data(), sizeof(int) * VP, cudaMemcpyHostToDevice));
  {
    std::vector<int> inv(V, 0);
    for (int i = 0; i < VP; i++)
      if (perm[i] >= 0) inv[perm[i]] = i;
    JRR_CUDA(cudaMemcpy(m->inv_perm, inv.data(), sizeof(int) * V, cudaMemcpyHostToDevice));
  }
  if (!vx_src.empty()) {
    JRR_CUDA(cudaMemcpy(m->vx_src, vx_src.data(), sizeof(int) * vx_src.size(), cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(m->vx_coef, vx_coef.data(), sizeof(float) * vx_coef.size(), cudaMemcpyHostToDevice));
  }
  m->flush_off[0] = 0;
  for (int pass = 0; pass < m->n_pass; pass++) {
    PassTab& t = m->passes[pass];
    std::vector<VtxRec> rec_f, rec_b, rec_l, rec_s;
    std::vector<int> flush_joint, range_base(NSPLIT_B + 1), flush_joint_l, range_base_l(VP / 192 + 1, 0);
    std::vector<int> flush_joint_s, range_base_s(NSPLIT_S + 1, 0);
    build(pass, VS_F, rec_f, nullptr, nullptr);
    build(pass, VS_B, rec_b, &flush_joint, &range_base);            // module backward: every vertex, 768-vertex ranges
    build(pass, m->vs_l, rec_l, &flush_joint_l, &range_base_l);     // fused backward: vs_l-vertex ranges
    build(pass, VS_S, rec_s, &flush_joint_s, &range_base_s);        // module backward, small batches: every vertex, 192-vertex ranges
    range_base_s[NSPLIT_S] = (int)flush_joint_s.size();
    t.n_flush_s = (int)flush_joint_s.size();
    range_base[NSPLIT_B] = (int)flush_joint.size();
    t.n_flush = (int)flush_joint.size();
    // the fused backward only walks the active ranges: its flush ids are a prefix
    flush_joint_l.resize(range_base_l[m->nsplit_act] ? range_base_l[m->nsplit_act] : flush_joint_l.size());
    t.n_flush_l = (int)flush_joint_l.size();
    std::vector<int> fptr, fidx, fptr_l, fidx_l, fptr_s, fidx_s;
    csr(flush_joint, fptr, fidx);
    csr(flush_joint_l, fptr_l, fidx_l);
    csr(flush_joint_s, fptr_s, fidx_s);
    if (fidx_s.size() > (size_t)4 * VP + 4 * NSPLIT_S) return fail(JRR_ERR_INVALID, "flush list overflow");
    if (fidx.size() > (size_t)4 * VP + 4 * NSPLIT_B || fidx_l.size() > (size_t)4 * VP + 4 * 36)
      return fail(JRR_ERR_INVALID, "flush list overflow");
    JRR_CUDA(cudaMemcpy(t.vrec, rec_f.data(), sizeof(VtxRec) * VP, cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.vrec_b, rec_b.data(), sizeof(VtxRec) * VP, cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.flush_ptr, fptr.data(), sizeof(int) * (NJ + 1), cudaMemcpyHostToDevice));
    if (!fidx.empty()) JRR_CUDA(cudaMemcpy(t.flush_idx, fidx.data(), sizeof(int) * fidx.size(), cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.range_flush_base, range_base.data(), sizeof(int) * (NSPLIT_B + 1), cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.vrec_l, rec_l.data(), sizeof(VtxRec) * VP, cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.flush_ptr_l, fptr_l.data(), sizeof(int) * (NJ + 1), cudaMemcpyHostToDevice));
    if (!fidx_l.empty()) JRR_CUDA(cudaMemcpy(t.flush_idx_l, fidx_l.data(), sizeof(int) * fidx_l.size(), cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.range_flush_base_l, range_base_l.data(), sizeof(int) * 37, cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.vrec_s, rec_s.data(), sizeof(VtxRec) * VP, cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.flush_ptr_s, fptr_s.data(), sizeof(int) * (NJ + 1), cudaMemcpyHostToDevice));
    if (!fidx_s.empty()) JRR_CUDA(cudaMemcpy(t.flush_idx_s, fidx_s.data(), sizeof(int) * fidx_s.size(), cudaMemcpyHostToDevice));
    JRR_CUDA(cudaMemcpy(t.range_flush_base_s, range_base_s.data(), sizeof(int) * (NSPLIT_S + 1), cudaMemcpyHostToDevice));
    m->flush_off[pass + 1] = m->flush_off[pass] + std::max(std::max(t.n_flush, t.n_flush_l), t.n_flush_s);
  }
  m->select_pass(0);
  const int64_t n = (int64_t)KA * NP;
  repack_blend_kernel<<<(unsigned)((n + 255) / 256), 256>>>(m->Pn, m->perm, m->P_hi, m->P_lo, m->Pt_hi, m->Pt_lo, m->P3_hi, m->P3_lo);
  JRR_CUDA(cudaGetLastError());
  JRR_CUDA(cudaDeviceSynchronize());
  return JRR_OK;
}

}  // namespace jrr

using namespace jrr;

extern "C" const char* jrr_last_error(void) { return g_err.c_str(); }
extern "C" int jrr_abi_version(void) { return JRR_ABI_VERSION; }
extern "C" int64_t jrr_last_launch_count(void) { return g_launches; }

extern "C" int jrr_model_destroy(JrrModel* m) {
  if (!m) return JRR_OK;
  cudaSetDevice(m->device);
  for (void* p : m->allocs) cudaFree(p);
  if (m->fold_ev) cudaFree(m->fold_ev);
  if (m->side) cudaStreamDestroy(m->side);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->ev_seed) cudaEventDestroy(m->ev_seed);
  if (m->ev_join2) cudaEventDestroy(m->ev_join2);
  delete m;
  return JRR_OK;
}

static int model_create_impl(const JrrModelDesc* d, JrrModel* m) {
  m->device = d->device;
  m->gemm_impl = d->gemm_impl;
  JRR_CUDA(cudaSetDevice(d->device));
  JRR_CUDA(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, d->device));
  // kernel-validation switches (the product configuration is all defaults)
  if (const char* e = getenv("JRR_OVERLAP_CRITIC")) m->overlap_critic = (e[0] != '0');
  if (const char* e = getenv("JRR_FUSED_FWD")) m->fused_fwd = (e[0] != '0');
  if (const char* e = getenv("JRR_FUSED_BWD")) m->fused_bwd = (e[0] != '0');
  if (const char* e = getenv("JRR_COMPACT_ACTIVE")) m->compact_active = (e[0] != '0');
  if (const char* e = getenv("JRR_CRITIC_HEAD_FUSED")) m->critic_head_fused = (e[0] != '0');
  if (const char* e = getenv("JRR_CRITIC_TS")) m->critic_ts = (e[0] != '0');
  if (const char* e = getenv("JRR_FOLD_TS")) m->fold_ts = (e[0] != '0');
  if (const char* e = getenv("JRR_SPLIT_ADAM")) m->split_adam = (e[0] != '0');
  if (const char* e = getenv("JRR_CRITIC_HEADLESS")) m->critic_headless = (e[0] != '0');
  if (const char* e = getenv("JRR_LOSS_PATH")) m->folded = (e[0] == 'f' || e[0] == '1') && m->gemm_impl == 0;
  if (m->gemm_impl != 0) { m->fused_fwd = false; m->fused_bwd = false; m->critic_head_fused = false; }
  if (!m->critic_head_fused) m->critic_ts = false;
  if (!m->critic_ts) m->critic_headless = false;
  // the active-vertex prefix is only walked by the two fused kernels; the stand-alone skinning
  // kernels always process every vertex, so compaction is tied to the fused configuration
  if (!(m->fused_fwd && m->fused_bwd)) m->compact_active = false;

  // ---- kinematic tree
  ChainTab& ct = m->chain;
  ct.max_depth = 0;
  for (int j = 0; j < NJ; j++)
    for (int c = 0; c < MAXCH; c++) ct.child[j][c] = -1;
  for (int j = 0; j < NJ; j++) {
    int p = (int)d->parents_host[j];
    if (j == 0) p = -1;
    if (j > 0 && (p < 0 || p >= j)) return fail(JRR_ERR_INVALID, "parents must precede their children");
    ct.parent[j] = p;
    ct.depth[j] = p < 0 ? 0 : ct.depth[p] + 1;
    ct.max_depth = std::max(ct.max_depth, ct.depth[j]);
    if (p >= 0) {
      int c = 0;
      while (c < MAXCH && ct.child[p][c] >= 0) c++;
      if (c == MAXCH) return fail(JRR_ERR_INVALID, "more than 4 children per joint not supported");
      ct.child[p][c] = j;
    }
  }
  for (int dd = 0; dd < NJ; dd++) ct.maxch[dd] = 0;
  for (int j = 0; j < NJ; j++) {
    int n = 0;
    for (int c = 0; c < MAXCH; c++) n += ct.child[j][c] >= 0;
    if (ct.depth[j] + 1 < NJ) ct.maxch[ct.depth[j] + 1] = std::max(ct.maxch[ct.depth[j] + 1], n);
  }

  // ---- host copies the (re)packing needs: joint-set key and <= 4 (joint, weight) pairs per vertex
  m->h_key.assign(V, 0);
  m->h_lbs.assign(V, {});
  for (int v = 0; v < V; v++) {
    uint32_t k = 0;
    for (int j = 0; j < NJ; j++) {
      const float x = d->lbs_weights_host[(size_t)v * NJ + j];
      if (x != 0.f) { k |= 1u << j; m->h_lbs[v].push_back({j, x}); }
    }
    // largest weights first: pass 0 then carries most of every vertex (and a 4-weight model is exactly one pass)
    std::stable_sort(m->h_lbs[v].begin(), m->h_lbs[v].end(),
                     [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return std::fabs(a.second) > std::fabs(b.second); });
    m->n_pass = std::max(m->n_pass, (int)(m->h_lbs[v].size() + 3) / 4);
    m->h_key[v] = k;
  }
  if (m->n_pass > 1 && m->gemm_impl != 0)
    return fail(JRR_ERR_INVALID, "lbs_weights rows with more than 4 non-zeros need the tensor-core build (gemm_impl 0): "
                                 "the SIMT validation kernels are single-pass");
  if (m->n_pass > 1) { m->fused_fwd = true; m->fused_bwd = true; }      // the unfused validation kernels are single-pass

  // ---- natural-order master of the augmented blend matrix [224][3*6890] (rows: posedirs, shapedirs^T, template)
  {
    std::vector<float> Pn((size_t)KA * 3 * V, 0.f);
    for (int k = 0; k < NF; k++)
      std::memcpy(&Pn[(size_t)k * 3 * V], d->posedirs_host + (size_t)k * 3 * V, sizeof(float) * 3 * V);
    for (int v = 0; v < V; v++)
      for (int c = 0; c < 3; c++) {
        for (int l = 0; l < NB; l++) Pn[(size_t)(FEAT_BETA + l) * 3 * V + 3 * v + c] = d->shapedirs_host[((size_t)v * 3 + c) * NB + l];
        Pn[(size_t)FEAT_ONE * 3 * V + 3 * v + c] = d->v_template_host[(size_t)v * 3 + c];
      }
    if (int rc = upload(m, &m->Pn, Pn)) return rc;
    if (int rc = dalloc(m, &m->P_hi, (size_t)KA * NP, false)) return rc;
    if (int rc = dalloc(m, &m->P_lo, (size_t)KA * NP, false)) return rc;
    if (int rc = dalloc(m, &m->Pt_hi, (size_t)KA * NP, false)) return rc;
    if (int rc = dalloc(m, &m->P3_hi, (size_t)KA * NP, false)) return rc;
    if (int rc = dalloc(m, &m->P3_lo, (size_t)KA * NP, false)) return rc;
    if (int rc = dalloc(m, &m->Pt_lo, (size_t)KA * NP, false)) return rc;
  }

  // ---- rest-joint pre-contraction
  {
    std::vector<float> J0(NJ * 3), JS(NJ * 3 * NB);
    for (int j = 0; j < NJ; j++)
      for (int c = 0; c < 3; c++) {
        double a = 0.0;
        double s[NB] = {0};
        for (int v = 0; v < V; v++) {
          const double r = d->J_regressor_host[(size_t)j * V + v];
          if (r == 0.0) continue;
          a += r * d->v_template_host[(size_t)v * 3 + c];
          for (int l = 0; l < NB; l++) s[l] += r * d->shapedirs_host[((size_t)v * 3 + c) * NB + l];
        }
        J0[j * 3 + c] = (float)a;
        for (int l = 0; l < NB; l++) JS[(j * 3 + c) * NB + l] = (float)s[l];
      }
    if (int rc = upload(m, &m->J0, J0)) return rc;
    if (int rc = upload(m, &m->JS, JS)) return rc;
  }

  // ---- joints49 tables
  m->h_vx.assign(V, {});      // per ORIGINAL vertex: (source-24, coef)
  {
    std::vector<int> picks(JRR_NUM_PICKS), jm(JRR_NUM_OUT_JOINTS);
    for (int p = 0; p < JRR_NUM_PICKS; p++) {
      int64_t v = d->vertex_picks_host[p];
      if (v < 0 || v >= V) return fail(JRR_ERR_INVALID, "vertex pick out of range");
      picks[p] = (int)v;
      m->h_vx[v].push_back({p, 1.f});
    }
    for (int o = 0; o < JRR_NUM_OUT_JOINTS; o++) {
      int64_t s2 = d->joint_map_host[o];
      if (s2 < 0 || s2 >= 54) return fail(JRR_ERR_INVALID, "joint_map entry out of range");
      jm[o] = (int)s2;
    }
    std::vector<int> ptr(JRR_NUM_EXTRA + 1, 0), col;
    std::vector<float> val;
    for (int e = 0; e < JRR_NUM_EXTRA; e++) {
      for (int v = 0; v < V; v++) {
        const float x = d->J_regressor_extra_host[(size_t)e * V + v];
        if (x != 0.f) {
          col.push_back(v);
          val.push_back(x);
          m->h_vx[v].push_back({JRR_NUM_PICKS + e, x});
        }
      }
      ptr[e + 1] = (int)col.size();
    }
    m->extra.rows = JRR_NUM_EXTRA;
    if (int rc = upload(m, &m->extra.ptr, ptr)) return rc;
    if (int rc = upload(m, &m->extra.col, col)) return rc;
    if (int rc = upload(m, &m->extra.val, val)) return rc;
    if (int rc = upload(m, &m->picks, picks)) return rc;
    if (int rc = upload(m, &m->joint_map, jm)) return rc;
    size_t nvx = 0;
    for (auto& l : m->h_vx) nvx += l.size();
    if (int rc = dalloc(m, &m->vx_src, nvx)) return rc;
    if (int rc = dalloc(m, &m->vx_coef, nvx)) return rc;
  }

  // ---- device buffers of the packing (filled by build_packing; sizes do not depend on the order)
  if (int rc = dalloc(m, &m->perm, VP)) return rc;
  if (int rc = dalloc(m, &m->inv_perm, V)) return rc;
  if (int rc = dalloc(m, &m->small_counter, 1)) return rc;
  if (m->n_pass > MAX_PASS) return fail(JRR_ERR_INVALID, "more than 24 skinning weights per vertex");
  for (int pass = 0; pass < m->n_pass; pass++) {
    PassTab& t = m->passes[pass];
    if (int rc = dalloc(m, &t.vrec, VP)) return rc;
    if (int rc = dalloc(m, &t.vrec_b, VP)) return rc;
    if (int rc = dalloc(m, &t.flush_ptr, NJ + 1)) return rc;
    if (int rc = dalloc(m, &t.flush_idx, (size_t)4 * VP + 4 * NSPLIT_B)) return rc;
    if (int rc = dalloc(m, &t.range_flush_base, NSPLIT_B + 1)) return rc;
    if (int rc = dalloc(m, &t.vrec_l, VP)) return rc;
    if (int rc = dalloc(m, &t.flush_ptr_l, NJ + 1)) return rc;
    if (int rc = dalloc(m, &t.flush_idx_l, (size_t)4 * VP + 4 * 36)) return rc;
    if (int rc = dalloc(m, &t.range_flush_base_l, 37)) return rc;
    if (int rc = dalloc(m, &t.vrec_s, VP)) return rc;
    if (int rc = dalloc(m, &t.flush_ptr_s, NJ + 1)) return rc;
    if (int rc = dalloc(m, &t.flush_idx_s, (size_t)4 * VP + 4 * NSPLIT_S)) return rc;
    if (int rc = dalloc(m, &t.range_flush_base_s, NSPLIT_S + 1)) return rc;
  }
  m->select_pass(0);
  if (int rc = dalloc(m, &m->active_dev, V)) return rc;
  if (int rc = build_packing(m, std::vector<uint8_t>(V, 1))) return rc;

  // ---- regressor state
  if (int rc = dalloc(m, &m->Jhat, (size_t)NH * V)) return rc;
  if (int rc = dalloc(m, &m->rowsum, NH)) return rc;
  if (int rc = dalloc(m, &m->regdot, NH)) return rc;
  // ---- critic state
  if (int rc = dalloc(m, &m->critic_small, 8192)) return rc;
  if (int rc = dalloc(m, &m->shape_critic, 256)) return rc;
  if (int rc = dalloc(m, &m->W1_hi, (size_t)C_Z * C_H)) return rc;
  if (int rc = dalloc(m, &m->W1_lo, (size_t)C_Z * C_H)) return rc;
  if (int rc = dalloc(m, &m->W1t_hi, (size_t)C_Z * C_H)) return rc;
  if (int rc = dalloc(m, &m->W1t_lo, (size_t)C_Z * C_H)) return rc;
  if (int rc = dalloc(m, &m->W2_hi, (size_t)C_Z * C_Z)) return rc;
  if (int rc = dalloc(m, &m->W2_lo, (size_t)C_Z * C_Z)) return rc;
  if (int rc = dalloc(m, &m->W2t_hi, (size_t)C_Z * C_Z)) return rc;
  if (int rc = dalloc(m, &m->W2t_lo, (size_t)C_Z * C_Z)) return rc;
  if (int rc = dalloc(m, &m->W2tw_hi, (size_t)C_Z * C_Z)) return rc;
  if (int rc = dalloc(m, &m->W2tw_lo, (size_t)C_Z * C_Z)) return rc;
  JRR_CUDA(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
  JRR_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  JRR_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
  JRR_CUDA(cudaEventCreateWithFlags(&m->ev_seed, cudaEventDisableTiming));
  JRR_CUDA(cudaEventCreateWithFlags(&m->ev_join2, cudaEventDisableTiming));
  JRR_CUDA(cudaDeviceSynchronize());
  return JRR_OK;
}

extern "C" int jrr_model_create(const JrrModelDesc* d, JrrModel** out) {
  if (!d || !out) return fail(JRR_ERR_INVALID, "null argument");
  if (!d->v_template_host || !d->shapedirs_host || !d->posedirs_host || !d->J_regressor_host ||
      !d->parents_host || !d->lbs_weights_host || !d->J_regressor_extra_host || !d->joint_map_host ||
      !d->vertex_picks_host)
    return fail(JRR_ERR_INVALID, "JrrModelDesc has a null array");
  JrrModel* m = new JrrModel();
  int rc = model_create_impl(d, m);
  if (rc != JRR_OK) {
    std::string keep = g_err;
    jrr_model_destroy(m);
    g_err = keep;
    return rc;
  }
  *out = m;
  return JRR_OK;
}

extern "C" int jrr_set_regressor(JrrModel* m, const float* J17_raw, const float* mask, void* stream) {
  if (!m || !J17_raw) return fail(JRR_ERR_INVALID, "null argument");
  reset_launch_count();
  cudaStream_t st = (cudaStream_t)stream;
  // Normalise first (fills Jhat), then make sure every vertex with a non-zero column sits in the
  // "active" prefix of the packing.  This entry point is outside the hot loop and SYNCHRONISES:
  // the support is read back and, if it is not covered by the current packing, the packed
  // constants are rebuilt (the refit itself can only shrink the support, so it never repacks).
  reg_rowsum_kernel<<<NH, 256, 0, st>>>(J17_raw, mask, m->rowsum);
  JRR_LAUNCH_CHECK();
  reg_normalise_kernel<<<(NH * V + 255) / 256, 256, 0, st>>>(J17_raw, mask, m->rowsum, m->Jhat);
  JRR_LAUNCH_CHECK();
  if (m->compact_active) {
    reg_active_kernel<<<(V + 255) / 256, 256, 0, st>>>(m->Jhat, m->active_dev);
    JRR_LAUNCH_CHECK();
    std::vector<uint8_t> act(V);
    JRR_CUDA(cudaMemcpyAsync(act.data(), m->active_dev, V, cudaMemcpyDeviceToHost, st));
    JRR_CUDA(cudaStreamSynchronize(st));
    bool covered = true;
    int n_act = 0;
    for (int v = 0; v < V; v++) { n_act += act[v]; if (act[v] && !m->packed_active[v]) covered = false; }
    // repack when the support is not covered, or when it shrank enough to drop a whole range
    const int vs = loss_range_size(std::max(n_act, 1));
    const int want = (int)std::min<int64_t>(VP, round_up(std::max(n_act, 1), vs));
    if (!covered || want < m->nv_act) {
      if (int rc = build_packing(m, act)) return rc;
    }
  }
  if (int rc = launch_regressor_normalise(m, J17_raw, mask, st)) return rc;
  return m->folded ? launch_fold(m, st) : JRR_OK;
}

extern "C" int jrr_critic_load(JrrModel* m, const float* p, void* stream) {
  if (!m || !p) return fail(JRR_ERR_INVALID, "null argument");
  reset_launch_count();
  return critic_load_impl(m, p, (cudaStream_t)stream);
}

int jrr::critic_load_impl(JrrModel* m, const float* p, cudaStream_t st) {
  // state_dict order (see jrr.h)
  const float* c1w = p;               // 192
  const float* c1b = c1w + 192;       // 32
  const float* c2w = c1b + 32;        // 1024
  const float* c2b = c2w + 1024;      // 32
  const float* heads = c2b + 32;      // 24 x (32 + 1)
  const float* W1 = heads + 24 * 33;  // 1024*768
  const float* b1 = W1 + (size_t)C_Z * C_H;
  const float* W2 = b1 + C_Z;
  const float* b2 = W2 + (size_t)C_Z * C_Z;
  const float* w3 = b2 + C_Z;
  const float* b3 = w3 + C_Z;
  float* cs = m->critic_small;
  JRR_CUDA(cudaMemcpyAsync(cs + 0, c1w, 192 * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 192, c1b, 32 * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 224, c2w, 1024 * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 1248, c2b, 32 * 4, cudaMemcpyDeviceToDevice, st));
  // heads: weight[32], bias[1] interleaved per joint -> [24][32] and [24]
  JRR_CUDA(cudaMemcpy2DAsync(cs + 1280, 32 * 4, heads, 33 * 4, 32 * 4, 24, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpy2DAsync(cs + 2048, 4, heads + 32, 33 * 4, 4, 24, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 2072, b1, C_Z * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 3096, b2, C_Z * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 4120, w3, C_Z * 4, cudaMemcpyDeviceToDevice, st));
  JRR_CUDA(cudaMemcpyAsync(cs + 5144, b3, 4, cudaMemcpyDeviceToDevice, st));
  const int64_t n1 = (int64_t)C_Z * C_H, n2 = (int64_t)C_Z * C_Z;
  split_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(W1, n1, m->W1_hi, m->W1_lo);
  JRR_LAUNCH_CHECK();
  split_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(W2, n2, m->W2_hi, m->W2_lo);
  JRR_LAUNCH_CHECK();
  split_transpose_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(W1, C_Z, C_H, m->W1t_hi, m->W1t_lo);
  JRR_LAUNCH_CHECK();
  split_transpose_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(W2, C_Z, C_Z, m->W2t_hi, m->W2t_lo);
  JRR_LAUNCH_CHECK();
  // (diag(w3) W2)^T: with it the layer-2 backward reads the ReLU mask itself as its A operand (critic_backward_gemms)
  split_transpose_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(W2, C_Z, C_Z, m->W2tw_hi, m->W2tw_lo, w3);
  JRR_LAUNCH_CHECK();
  m->has_critic = true;
  return JRR_OK;
}

extern "C" int jrr_shape_critic_load(JrrModel* m, const float* p, float w_shape, void* stream) {
  if (!m) return fail(JRR_ERR_INVALID, "null argument");
  reset_launch_count();
  if (p == nullptr) {   // switch the term off
    m->has_shape_critic = false;
    m->w_shape = 0.f;
    return JRR_OK;
  }
  // state_dict order is already the kernel's layout: W0[10][10] b0[10] W1[5][10] b1[5] W2[1][5] b2[1]
  JRR_CUDA(cudaMemcpyAsync(m->shape_critic, p, JRR_SHAPE_CRITIC_PARAMS * sizeof(float), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  m->has_shape_critic = true;
  m->w_shape = w_shape;
  return JRR_OK;
}

extern "C" int jrr_set_loss_path(JrrModel* m, int mode, void* stream) {
  if (!m) return fail(JRR_ERR_INVALID, "null argument");
  if (mode != JRR_LOSS_PATH_VERTEX && mode != JRR_LOSS_PATH_FOLDED) return fail(JRR_ERR_INVALID, "unknown loss path");
  if (mode == JRR_LOSS_PATH_FOLDED && m->gemm_impl != 0) return fail(JRR_ERR_INVALID, "the folded loss path needs the tcgen05 GEMM");
  reset_launch_count();
  m->folded = mode == JRR_LOSS_PATH_FOLDED;
  if (m->folded && m->has_regressor) return launch_fold(m, (cudaStream_t)stream);
  return JRR_OK;
}
