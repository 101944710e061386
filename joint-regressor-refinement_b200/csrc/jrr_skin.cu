// Linear blend skinning over the sparse skinning weights, fused with the 17x6890
// J-regressor reduction (forward) and with the regressor-transpose seed, the dA reduction
// and the pose-blend gradient (backward).  One thread owns one pose and walks a range of
// packed vertices; every per-vertex constant (joint ids, weights, regressor column) is a
// warp-uniform broadcast load, the four joint transforms a vertex needs are cached in
// registers and only re-fetched when the joint id of a slot changes ("runs"), and all
// per-pose arrays are laid out pose-contiguous so every global access is a coalesced 128 B
// line.  Reductions over vertices stay inside a thread (fixed order, no float atomics).
//
// Replaces (file:line under /root/reference): the W.A / T.v part of smplx.lbs.lbs reached
// through scripts/smpl.py:72-74, vertex_joint_selector + J_regressor_extra + joint_map
// (scripts/smpl.py:75-78), the regressor contraction of utils.find_joints
// (scripts/utils.py:96-98), move_pelvis + MSELoss (scripts/utils.py:106-114,
// scripts/optimize.py:238-239) and their autograd backward (scripts/optimize.py:264).
#include "jrr_internal.cuh"
#include "jrr_f32x2.cuh"

namespace jrr {

constexpr int SK_THREADS = 128;
constexpr int VT = 32;               // vertices per tile (constants staging + transposition)
constexpr int TILE_LD = 3 * VT + 1;  // 97: conflict-free column access
constexpr int GV = 4;                // vertices per register-prefetch group
constexpr int TILE_F4 = VT * REC_WORDS / 4;  // 224 float4 of vertex records per tile

__device__ __forceinline__ float tf32_hi_k(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// the tile's vertex records live in shared memory; every read below is a warp-wide broadcast
struct RecView {
  const float4* p;
  __device__ __forceinline__ void head(int lv, uint32_t& meta, float w[4], int& xptr, int& xcnt) const {
    const float4 a = p[lv * 7], c = p[lv * 7 + 1];
    meta = __float_as_uint(a.x);
    w[0] = a.y; w[1] = a.z; w[2] = a.w; w[3] = c.x;
    xptr = __float_as_int(c.y); xcnt = __float_as_int(c.z);
  }
  __device__ __forceinline__ void jh(int lv, float out[JH_STRIDE]) const {
#pragma unroll
    for (int q = 0; q < JH_STRIDE / 4; q++) {
      const float4 t = p[lv * 7 + 2 + q];
      out[q * 4 + 0] = t.x; out[q * 4 + 1] = t.y; out[q * 4 + 2] = t.z; out[q * 4 + 3] = t.w;
    }
  }
};

__device__ __forceinline__ void load_vp_group(const float* __restrict__ vpT, int64_t BP, int64_t b, int i,
                                              float out[3 * GV]) {
  const float* src = vpT + (int64_t)(3 * i) * BP + b;
#pragma unroll
  for (int q = 0; q < 3 * GV; q++) out[q] = src[(int64_t)q * BP];
}

// ---------------------------------------------------------------------------- forward
template <bool WRITE_V, bool WRITE_VT, bool PART>
__global__ void __launch_bounds__(SK_THREADS, 3)
skin_fwd_kernel(const VtxRec* __restrict__ vrec, const int* __restrict__ perm,
                const float* __restrict__ AT, const float* __restrict__ vpT, int64_t B, int64_t BP,
                float* __restrict__ vertices_out, float* __restrict__ vT_out, float* __restrict__ part) {
  extern __shared__ float4 smem4[];
  float4* sconst = smem4;                                         // [2][224]
  float* tile = reinterpret_cast<float*>(smem4 + 2 * TILE_F4);    // [128][97] when WRITE_V
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * SK_THREADS;
  const int64_t b = b0 + tid;
  const int s = blockIdx.y;
  const int i0 = s * VS_F;
  constexpr int NT = VS_F / VT;
  float A[4][12];
  float acc[PART ? NACC : 1];
#pragma unroll
  for (int a = 0; a < (PART ? NACC : 1); a++) acc[a] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int e = 0; e < 12; e++) A[k][e] = 0.f;

  {
    const float4* g = reinterpret_cast<const float4*>(vrec + i0);
    for (int e = tid; e < TILE_F4; e += SK_THREADS) sconst[e] = __ldg(g + e);
  }
  float nx[3 * GV];
  load_vp_group(vpT, BP, b, i0, nx);
  __syncthreads();

#pragma unroll 1
  for (int t = 0; t < NT; t++) {
    const RecView rv{sconst + (t & 1) * TILE_F4};
    const bool has_next = t + 1 < NT;
    float4 pf0 = make_float4(0, 0, 0, 0), pf1 = pf0;
    if (has_next) {
      const float4* g = reinterpret_cast<const float4*>(vrec + i0 + (t + 1) * VT);
      pf0 = __ldg(g + tid);
      if (tid + SK_THREADS < TILE_F4) pf1 = __ldg(g + tid + SK_THREADS);
    }
#pragma unroll 1
    for (int sub = 0; sub < VT / GV; sub++) {
      float cur[3 * GV];
#pragma unroll
      for (int q = 0; q < 3 * GV; q++) cur[q] = nx[q];
      const int inext = i0 + t * VT + (sub + 1) * GV;
      if (inext < i0 + VS_F) load_vp_group(vpT, BP, b, inext, nx);
#pragma unroll
      for (int ii = 0; ii < GV; ii++) {
        const int lv = sub * GV + ii;
        const int i = i0 + t * VT + lv;
        uint32_t meta; float w[4]; int xptr, xcnt;
        rv.head(lv, meta, w, xptr, xcnt);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if ((meta >> (20 + k)) & 1u) {
            const int j = (meta >> (5 * k)) & 31u;
            const float* src = AT + (int64_t)(j * 12) * BP + b;
#pragma unroll
            for (int e = 0; e < 12; e++) A[k][e] = src[(int64_t)e * BP];
          }
        }
        const float x = cur[ii * 3 + 0], y = cur[ii * 3 + 1], z = cur[ii * 3 + 2];
        float v[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const float t0 = w[0] * A[0][r * 4 + 0] + w[1] * A[1][r * 4 + 0] + w[2] * A[2][r * 4 + 0] + w[3] * A[3][r * 4 + 0];
          const float t1 = w[0] * A[0][r * 4 + 1] + w[1] * A[1][r * 4 + 1] + w[2] * A[2][r * 4 + 1] + w[3] * A[3][r * 4 + 1];
          const float t2 = w[0] * A[0][r * 4 + 2] + w[1] * A[1][r * 4 + 2] + w[2] * A[2][r * 4 + 2] + w[3] * A[3][r * 4 + 2];
          const float t3 = w[0] * A[0][r * 4 + 3] + w[1] * A[1][r * 4 + 3] + w[2] * A[2][r * 4 + 3] + w[3] * A[3][r * 4 + 3];
          v[r] = t0 * x + t1 * y + t2 * z + t3;
        }
        if (WRITE_VT) {
#pragma unroll
          for (int r = 0; r < 3; r++) vT_out[(int64_t)(3 * i + r) * BP + b] = v[r];
        }
        if (WRITE_V) {
#pragma unroll
          for (int r = 0; r < 3; r++) tile[tid * TILE_LD + lv * 3 + r] = v[r];
        }
        if (PART && ((meta >> 24) & 1u)) {
          float jh[JH_STRIDE];
          rv.jh(lv, jh);
#pragma unroll
          for (int j = 0; j < NH; j++) {
            acc[PART ? j * 3 + 0 : 0] = fmaf(jh[j], v[0], acc[PART ? j * 3 + 0 : 0]);
            acc[PART ? j * 3 + 1 : 0] = fmaf(jh[j], v[1], acc[PART ? j * 3 + 1 : 0]);
            acc[PART ? j * 3 + 2 : 0] = fmaf(jh[j], v[2], acc[PART ? j * 3 + 2 : 0]);
          }
        }
      }
    }
    if (WRITE_V) {
      // natural-order vertices [B][6890][3]: rows leave through smem, scattered by the packing
      // permutation in 12-byte pieces
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      const int it0 = i0 + t * VT;
      int dstoff[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int f = lane + 32 * q;
        const int v = perm[it0 + f / 3];
        dstoff[q] = v >= 0 ? v * 3 + f % 3 : -1;
      }
      for (int r = warp; r < SK_THREADS; r += SK_THREADS / 32) {
        const int64_t bb = b0 + r;
        if (bb >= B) break;
        float* dst = vertices_out + bb * (int64_t)(V * 3);
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (dstoff[q] >= 0) dst[dstoff[q]] = tile[r * TILE_LD + lane + 32 * q];
      }
    }
    if (has_next) {
      float4* dstc = sconst + ((t + 1) & 1) * TILE_F4;
      dstc[tid] = pf0;
      if (tid + SK_THREADS < TILE_F4) dstc[tid + SK_THREADS] = pf1;
    }
    __syncthreads();
  }
  if (PART) {
#pragma unroll
    for (int a = 0; a < (PART ? NACC : 1); a++) part[((int64_t)s * NACC + a) * BP + b] = acc[a];
  }
}

// ------------------------------------------------------------------ loss seed (joint term)
// pred = sum of the partial slots (fixed order); pc = pred - pred[0]; diff = pc - gt/1000;
// g = w_joint * 2 diff / (51 B_logical); pelvis adjustment; per-CTA loss partial.
// CTA = one pose block of 128.  Phase 1: 8 warps sum the slots row by row (4 independent
// 128-byte loads per slot and lane, slots unrolled) into smem; phase 2: 128 threads do the
// per-pose maths; phase 3: all threads write gT coalesced.
constexpr int LS_THREADS = 512;

// 2-D reprojection of the 17 regressed joints (scripts/renderer.py:35-49 with pytorch3d 0.3.0
// PerspectiveCameras, R = I, focal 5000/224, principal point 0, 224x224 screen):
//   P = 2*(-x, -y, z) + T;  ndc = f*P.xy/P.z;  screen = 111.5*(1 - ndc)
// Accumulates d loss/d pred (into g) and d loss/d T for residual scale `sc`; returns sum of squares.
__device__ __forceinline__ float proj2d_grad(const float pred[NACC], const float T[3], const float* __restrict__ gt2d,
                                             float sc, float g[NACC], float dT[3]) {
  const float f = 5000.f / 224.f, half = 111.5f;
  float loss = 0.f;
#pragma unroll
  for (int j = 0; j < NH; j++) {
    const float Px = -2.f * pred[j * 3 + 0] + T[0], Py = -2.f * pred[j * 3 + 1] + T[1], Pz = 2.f * pred[j * 3 + 2] + T[2];
    const float iz = 1.f / Pz;
    const float xn = f * Px * iz, yn = f * Py * iz;
    const float dx = half * (1.f - xn) - gt2d[j * 2 + 0], dy = half * (1.f - yn) - gt2d[j * 2 + 1];
    loss += dx * dx + dy * dy;
    const float dxn = -half * sc * dx, dyn = -half * sc * dy;
    const float dPx = dxn * f * iz, dPy = dyn * f * iz, dPz = -(dxn * xn + dyn * yn) * iz;
    dT[0] += dPx; dT[1] += dPy; dT[2] += dPz;
    if (g != nullptr) { g[j * 3 + 0] += -2.f * dPx; g[j * 3 + 1] += -2.f * dPy; g[j * 3 + 2] += 2.f * dPz; }
  }
  return loss;
}

__device__ __forceinline__ void adam_update3(float T[3], const float dT[3], float m[3], float v[3], int t, float lr) {
  const float bc2s = (float)sqrt(1.0 - pow(0.999, (double)t));
  const float step = (float)((double)lr / (1.0 - pow(0.9, (double)t)));
#pragma unroll
  for (int c = 0; c < 3; c++) {
    m[c] = 0.9f * m[c] + 0.1f * dT[c];
    v[c] = 0.999f * v[c] + 0.001f * dT[c] * dT[c];
    T[c] -= step * (m[c] / (sqrtf(v[c]) / bc2s + 1e-8f));
  }
}

// PPB poses per CTA: 32 (more CTAs in flight; the kernel is latency-bound on the partial-slot loads) or 128
template <int PPB>
__global__ void __launch_bounds__(LS_THREADS)
loss_seed_kernel(const float* __restrict__ part, int nslots, int n_tiles, int T, int G, int mdiv,
                 const float* __restrict__ gt_mm, int64_t B, int64_t BP, float scale,
                 float* __restrict__ gT, float* __restrict__ joints17_out, float* __restrict__ loss_part,
                 const Proj2D p2d, int n_pass, int64_t pass_stride) {
  constexpr int NC = PPB / 32;
  __shared__ float sp[NACC][PPB];
  __shared__ float red[NC], red2[NC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * PPB;
  if (nslots <= 0) {
    // partials written by the fused forward kernel: two per CTA segment of this pose block
    const int mb = (int)(b0 / 128) / mdiv;      // the fused forward's tile row: a 128-pose block, or a CTA pair's two
    const int c0 = (int)(((int64_t)mb * n_tiles * G) / T);
    const int c1 = (int)((((int64_t)(mb + 1) * n_tiles - 1) * G) / T);
    nslots = 2 * (c1 - c0 + 1);
  }
  const int64_t slot_stride = (int64_t)NACC * BP;
  for (int a = warp; a < NACC; a += LS_THREADS / 32) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0.f;
    for (int ps = 0; ps < n_pass; ps++) {          // one region of partial sums per skinning pass (SMPL: one)
      const float* src = part + ps * pass_stride + (int64_t)a * BP + b0 + lane;
      // 8 slots per round trip (predicated): the kernel is bound by the latency of these loads
      for (int s = 0; s < nslots; s += 8) {
        float v[8][NC];
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int c = 0; c < NC; c++) v[u][c] = (s + u < nslots) ? src[(int64_t)(s + u) * slot_stride + 32 * c] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int c = 0; c < NC; c++) acc[c] += v[u][c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) sp[a][lane + 32 * c] = acc[c];
  }
  __syncthreads();
  float loss = 0.f, loss2 = 0.f;
  if (tid < PPB) {
    const int64_t b = b0 + tid;
    float pred[NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) pred[a] = sp[a][tid];
    if (b < B && joints17_out != nullptr) {
#pragma unroll
      for (int a = 0; a < NACC; a++) joints17_out[b * NACC + a] = pred[a];
    }
    if (gT != nullptr) {
      float g[NACC];
      float sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < NACC; a++) {
        float d = 0.f;
        if (b < B) d = (pred[a] - pred[a % 3]) - gt_mm[b * NACC + a] / 1000.f;
        loss += d * d;
        g[a] = scale * d;
        sum[a % 3] += g[a];
      }
#pragma unroll
      for (int c = 0; c < 3; c++) g[c] -= sum[c];
      if (p2d.gt2d != nullptr) {
        // 2-D term: seeds the same backward through the (not pelvis-centred) joints, and its
        // camera gradient is complete here, so the camera's Adam step happens in place
        float l2 = 0.f;
        if (b < B) {
          float Tc[3], dT[3] = {0.f, 0.f, 0.f}, m3[3], v3[3];
          if (p2d.dcam_ext != nullptr) { dT[0] = p2d.dcam_ext[b * 3 + 0]; dT[1] = p2d.dcam_ext[b * 3 + 1]; dT[2] = p2d.dcam_ext[b * 3 + 2]; }
#pragma unroll
          for (int c = 0; c < 3; c++) { Tc[c] = p2d.cam[b * 3 + c]; m3[c] = p2d.cam_m[b * 3 + c]; v3[c] = p2d.cam_v[b * 3 + c]; }
          l2 = proj2d_grad(pred, Tc, p2d.gt2d + b * 34, p2d.scale, g, dT);
          adam_update3(Tc, dT, m3, v3, *p2d.step_count + 1, p2d.lr);
#pragma unroll
          for (int c = 0; c < 3; c++) { p2d.cam[b * 3 + c] = Tc[c]; p2d.cam_m[b * 3 + c] = m3[c]; p2d.cam_v[b * 3 + c] = v3[c]; }
        }
        loss2 = l2;
      }
#pragma unroll
      for (int a = 0; a < NACC; a++) sp[a][tid] = g[a];
    }
  }
  if (gT == nullptr) return;
  // deterministic block reduction of the loss (warps 0..3 hold the poses)
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    loss2 += __shfl_xor_sync(0xffffffffu, loss2, o);
  }
  if (tid < PPB && lane == 0) { red[warp] = loss; red2[warp] = loss2; }
  __syncthreads();
  if (tid == 0) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; c++) { t1 += red[c]; t2 += red2[c]; }
    loss_part[blockIdx.x] = t1;
    loss_part[LOSS_PART_2D + blockIdx.x] = t2;
  }
  for (int idx = tid; idx < NACC * PPB; idx += LS_THREADS) {
    const int a = idx / PPB, bl = idx % PPB;
    gT[(int64_t)a * BP + b0 + bl] = sp[a][bl];
  }
}

// ---------------------------------------------------------------------------- folded loss path
// The regressed joints are linear in the blended vertices and the blended vertices are linear in the
// blend features, with the joint transforms as the only pose-dependent factors in between:
//   joints17_i(b) = sum_j  A_j^R(b) . (T_ji . feat_b)  +  A_j^t(b) . c_ji
//   T_ji[c][k] = sum_v Jhat_iv w_vj P[3v+c][k]   (3 x 224),    c_ji = sum_v Jhat_iv w_vj
// T and c depend on the model and the regressor only (fold_kernel, once per regressor version), so
// the whole per-vertex pass (blend GEMM N = 20736, skinning, 17x6890 reduction and their backward)
// becomes Q = feat . T^T (N = 1224), this kernel, and dfeat = dQ . T.  Thread = pose:
//   forward   pred_i = sum_j A_j^R q_ji + A_j^t c_ji            (q from QT, pose-contiguous)
//   loss      exactly loss_seed_kernel's (pelvis-centred MSE seed, optional 2-D term + camera Adam)
//   backward  dA_j^R = sum_i g_i (x) q_ji,  dA_j^t = sum_i g_i c_ji,  dq_ji = A_j^R^T g_i
// dQ rows leave through shared memory so that each pose's 51 values per joint are written contiguously.
// Work split: 32 poses per CTA (lane = pose, so every load/store of the pose-contiguous arrays is one
// 128-byte line), 8 warps x 3 joints each; the partial joints meet in shared memory in a fixed order.
constexpr int FS_POSES = 32;
constexpr int FS_WARPS = 8;
constexpr int FS_JPW = NJ / FS_WARPS;          // 3 joints per warp
constexpr int FS_LD = 53;                      // dq staging row (odd: conflict-free transposed access)
__global__ void __launch_bounds__(FS_WARPS * 32)
folded_seed_kernel(const float* __restrict__ QT, const float* __restrict__ AT, const float* __restrict__ Tc,
                   const float* __restrict__ gt_mm, int64_t B, int64_t BP, float scale, const Proj2D p2d,
                   float* __restrict__ loss_part, float* __restrict__ joints17_out, float* __restrict__ dAT,
                   float* __restrict__ dQ_hi, float* __restrict__ dQ_lo, float* __restrict__ dc_part) {
  extern __shared__ float fs_smem[];
  float* sbuf = fs_smem;                                   // forward: partial joints [warp][51][32]; backward: dq staging
  float* sg = sbuf + FS_WARPS * FS_POSES * FS_LD;          // joints, then the loss seed g, [a][pose]
  float* sTc = sg + NACC * FS_POSES;                       // [24][17]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * FS_POSES, b = b0 + lane;
  for (int i = tid; i < NJ * NH; i += FS_WARPS * 32) sTc[i] = Tc[i];
  __syncthreads();
  {
    float pred[NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) pred[a] = 0.f;
#pragma unroll 1
    for (int jj = 0; jj < FS_JPW; jj++) {
      const int j = warp * FS_JPW + jj;
      float A[12];
#pragma unroll
      for (int e = 0; e < 12; e++) A[e] = AT[(int64_t)(j * 12 + e) * BP + b];
      const float* q = QT + (int64_t)(j * NH * 3) * BP + b;
#pragma unroll
      for (int i = 0; i < NH; i++) {
        const float q0 = q[(int64_t)(i * 3 + 0) * BP], q1 = q[(int64_t)(i * 3 + 1) * BP], q2 = q[(int64_t)(i * 3 + 2) * BP];
        const float tc = sTc[j * NH + i];
#pragma unroll
        for (int r = 0; r < 3; r++)
          pred[i * 3 + r] += fmaf(A[r * 4 + 0], q0, fmaf(A[r * 4 + 1], q1, fmaf(A[r * 4 + 2], q2, A[r * 4 + 3] * tc)));
      }
    }
#pragma unroll
    for (int a = 0; a < NACC; a++) sbuf[(warp * NACC + a) * FS_POSES + lane] = pred[a];
  }
  __syncthreads();
  // joints (fixed-order sum of the eight partial sets), spread over the warps: warp w owns components a = w, w + 8, ...
  for (int a = warp; a < NACC; a += FS_WARPS) {
    float t = sbuf[a * FS_POSES + lane];
#pragma unroll
    for (int w2 = 1; w2 < FS_WARPS; w2++) t += sbuf[(w2 * NACC + a) * FS_POSES + lane];
    sg[a * FS_POSES + lane] = t;
  }
  __syncthreads();
  if (joints17_out != nullptr) {       // [32 poses][51] is one contiguous block of the output: coalesced
    for (int e = tid; e < FS_POSES * NACC; e += FS_WARPS * 32) {
      const int p = e / NACC, a = e - p * NACC;
      if (b0 + p < B) joints17_out[b0 * NACC + e] = sg[a * FS_POSES + p];
    }
  }
  if (dAT != nullptr) {
    // Loss seed on all warps (it was one warp's serial section -- 51 strided loads, 51 divisions -- with the other seven
    // parked at the barrier: 44 % of the kernel's warp time).  Thread (w, pose) handles a = w, w + 8, ...; the ground
    // truth block [32][51] is contiguous and comes in through shared memory; pelvis sums and the loss meet in shared memory
    // in a fixed order.
    float* sgt = sbuf;                                      // [32][51] (the partial joints are consumed)
    float* ssum = sbuf + FS_POSES * NACC;                   // [8 warps][3][32]
    float* sred = ssum + FS_WARPS * 3 * FS_POSES;           // [8] loss
    float* spred = sred + 32;                               // [51][32] the joints, kept for the 2-D term
    for (int e = tid; e < FS_POSES * NACC; e += FS_WARPS * 32)
      sgt[e] = (b0 + e / NACC < B) ? gt_mm[b0 * NACC + e] : 0.f;
    __syncthreads();
    float loss = 0.f;
    float sum[3] = {0.f, 0.f, 0.f};
    float dloc[(NACC + FS_WARPS - 1) / FS_WARPS];
#pragma unroll
    for (int k = 0; k < (NACC + FS_WARPS - 1) / FS_WARPS; k++) {
      const int a = warp + k * FS_WARPS;
      float d = 0.f;
      if (a < NACC && b < B) d = (sg[a * FS_POSES + lane] - sg[(a % 3) * FS_POSES + lane]) - sgt[lane * NACC + a] / 1000.f;
      if (a < NACC && p2d.gt2d != nullptr) spred[a * FS_POSES + lane] = sg[a * FS_POSES + lane];
      loss += d * d;
      dloc[k] = scale * d;
      if (a < NACC) sum[a % 3] += dloc[k];
    }
    __syncthreads();                                         // every warp has read the joints it needs from sg
#pragma unroll
    for (int k = 0; k < (NACC + FS_WARPS - 1) / FS_WARPS; k++) {
      const int a = warp + k * FS_WARPS;
      if (a < NACC) sg[a * FS_POSES + lane] = dloc[k];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) ssum[(warp * 3 + c) * FS_POSES + lane] = sum[c];
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) sred[warp] = loss;
    __syncthreads();
    if (warp < 3) {                                          // pelvis adjustment g[c] -= sum over all components of g[. % 3 == c]
      float t = ssum[warp * FS_POSES + lane];
#pragma unroll
      for (int w2 = 1; w2 < FS_WARPS; w2++) t += ssum[(w2 * 3 + warp) * FS_POSES + lane];
      sg[warp * FS_POSES + lane] -= t;
    }
    float loss2 = 0.f;
    if (p2d.gt2d != nullptr) {                               // 2-D reprojection term + camera Adam: one warp, whole poses
      __syncthreads();
      if (warp == 0) {
        float pred[NACC], g[NACC];
#pragma unroll
        for (int a = 0; a < NACC; a++) { pred[a] = spred[a * FS_POSES + lane]; g[a] = sg[a * FS_POSES + lane]; }
        if (b < B) {
          float Tcam[3], dT[3] = {0.f, 0.f, 0.f}, m3[3], v3[3];
          if (p2d.dcam_ext != nullptr) { dT[0] = p2d.dcam_ext[b * 3 + 0]; dT[1] = p2d.dcam_ext[b * 3 + 1]; dT[2] = p2d.dcam_ext[b * 3 + 2]; }
#pragma unroll
          for (int c = 0; c < 3; c++) { Tcam[c] = p2d.cam[b * 3 + c]; m3[c] = p2d.cam_m[b * 3 + c]; v3[c] = p2d.cam_v[b * 3 + c]; }
          loss2 = proj2d_grad(pred, Tcam, p2d.gt2d + b * 34, p2d.scale, g, dT);
          adam_update3(Tcam, dT, m3, v3, *p2d.step_count + 1, p2d.lr);
#pragma unroll
          for (int c = 0; c < 3; c++) { p2d.cam[b * 3 + c] = Tcam[c]; p2d.cam_m[b * 3 + c] = m3[c]; p2d.cam_v[b * 3 + c] = v3[c]; }
        }
        for (int o = 16; o > 0; o >>= 1) loss2 += __shfl_xor_sync(0xffffffffu, loss2, o);
#pragma unroll
        for (int a = 0; a < NACC; a++) sg[a * FS_POSES + lane] = g[a];
      }
    }
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < FS_WARPS; w2++) t += sred[w2];
      loss_part[blockIdx.x] = t;
      loss_part[LOSS_PART_2D + blockIdx.x] = loss2;
    }
  }
  if (dAT == nullptr) return;      // joints only (uniform)
  __syncthreads();
  float* sdq = sbuf + warp * FS_POSES * FS_LD;
#pragma unroll 1
  for (int jj = 0; jj < FS_JPW; jj++) {
    const int j = warp * FS_JPW + jj;
    float A[12], dA[12];
#pragma unroll
    for (int e = 0; e < 12; e++) { A[e] = AT[(int64_t)(j * 12 + e) * BP + b]; dA[e] = 0.f; }
    const float* q = QT + (int64_t)(j * NH * 3) * BP + b;
#pragma unroll
    for (int i = 0; i < NH; i++) {
      const float q0 = q[(int64_t)(i * 3 + 0) * BP], q1 = q[(int64_t)(i * 3 + 1) * BP], q2 = q[(int64_t)(i * 3 + 2) * BP];
      const float tc = sTc[j * NH + i];
      const float g0 = sg[(i * 3 + 0) * FS_POSES + lane], g1 = sg[(i * 3 + 1) * FS_POSES + lane], g2 = sg[(i * 3 + 2) * FS_POSES + lane];
      dA[0] = fmaf(g0, q0, dA[0]); dA[1] = fmaf(g0, q1, dA[1]); dA[2] = fmaf(g0, q2, dA[2]); dA[3] = fmaf(g0, tc, dA[3]);
      dA[4] = fmaf(g1, q0, dA[4]); dA[5] = fmaf(g1, q1, dA[5]); dA[6] = fmaf(g1, q2, dA[6]); dA[7] = fmaf(g1, tc, dA[7]);
      dA[8] = fmaf(g2, q0, dA[8]); dA[9] = fmaf(g2, q1, dA[9]); dA[10] = fmaf(g2, q2, dA[10]); dA[11] = fmaf(g2, tc, dA[11]);
#pragma unroll
      for (int c = 0; c < 3; c++)
        sdq[lane * FS_LD + i * 3 + c] = fmaf(A[c], g0, fmaf(A[4 + c], g1, A[8 + c] * g2));
      if (dc_part != nullptr) {
        // regressor refit (folded form): d loss / d c_ji = sum_b g_i . A_j^t, this CTA's 32 poses (uniform branch)
        float t = fmaf(g0, A[3], fmaf(g1, A[7], g2 * A[11]));
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) dc_part[(int64_t)blockIdx.x * (NJ * NH) + j * NH + i] = t;
      }
    }
#pragma unroll
    for (int e = 0; e < 12; e++) dAT[(int64_t)(j * 12 + e) * BP + b] = dA[e];
    __syncwarp();
    // each pose's 51 values of this joint are contiguous in its dQ row (K-major A operand of the second GEMM)
#pragma unroll 4
    for (int p = 0; p < FS_POSES; p++) {
      const int64_t o = (b0 + p) * FOLD_NP + j * NACC;
      if (dQ_lo == nullptr) {        // plain fp32 rows: the GEMM splits them on their way into tensor memory (uniform)
        dQ_hi[o + lane] = sdq[p * FS_LD + lane];
        if (lane < NACC - 32) dQ_hi[o + 32 + lane] = sdq[p * FS_LD + 32 + lane];
        continue;
      }
      {
        const float x = sdq[p * FS_LD + lane];
        const float hi = tf32_hi_k(x);
        dQ_hi[o + lane] = hi;
        dQ_lo[o + lane] = tf32_hi_k(x - hi);
      }
      if (lane < NACC - 32) {
        const float x = sdq[p * FS_LD + 32 + lane];
        const float hi = tf32_hi_k(x);
        dQ_hi[o + 32 + lane] = hi;
        dQ_lo[o + 32 + lane] = tf32_hi_k(x - hi);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------- backward
// dv_i = Jhat^T g (USE_G) + dvertices (USE_DV, natural layout staged through smem) +
//        picks / extra-regressor rows of the joints49 gradient (USE_X)
// dvp_i = (sum_k w_k AR_k)^T dv_i                        -> dvp_hi/lo [BP][NP] (A operand of the
//                                                           backward blend GEMM, tf32 split)
// dA_k += w_k dv_i (x) [vp_i ; 1]                        -> flushed per run to dAflush
template <bool USE_G, bool USE_DV, bool USE_X>
__global__ void __launch_bounds__(SK_THREADS, 2)
skin_bwd_kernel(const VtxRec* __restrict__ vrec, const int* __restrict__ perm,
                const int* __restrict__ range_flush_base, const int* __restrict__ vx_src,
                const float* __restrict__ vx_coef, const float* __restrict__ AT,
                const float* __restrict__ vpT, const float* __restrict__ gT,
                const float* __restrict__ dvertices, const float* __restrict__ d30T, int64_t B,
                int64_t BP, float* __restrict__ dvp_hi, float* __restrict__ dvp_lo,
                float* __restrict__ dAflush) {
  extern __shared__ float4 smem4[];
  float4* sconst = smem4;                                             // [2][224]
  float* tile_out = reinterpret_cast<float*>(smem4 + 2 * TILE_F4);    // [128][97]
  float* tile_in = tile_out + SK_THREADS * TILE_LD;                   // [128][97] (USE_DV)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * SK_THREADS;
  const int64_t b = b0 + tid;
  const int s = blockIdx.y;
  const int i0 = s * VS_B;
  constexpr int NT = VS_B / VT;
  float* flush_dst = dAflush + (int64_t)range_flush_base[s] * 12 * BP + b;

  // loss seed, packed over joint pairs: gp[c][p] = (g[2p][c], g[2p+1][c])
  f32x2 gp[USE_G ? 3 : 1][USE_G ? 9 : 1];
  if (USE_G) {
#pragma unroll
    for (int p = 0; p < 9; p++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float lo = gT[(int64_t)((2 * p) * 3 + c) * BP + b];
        const float hi = (2 * p + 1 < NH) ? gT[(int64_t)((2 * p + 1) * 3 + c) * BP + b] : 0.f;
        gp[USE_G ? c : 0][USE_G ? p : 0] = pk2(lo, hi);
      }
  }
  // cached joint rotations (rows packed over columns 0/1, column 2 scalar) and the dA slot
  // accumulators (row r: columns (0,1) and (2,3))
  f32x2 AR01[4][3], dA01[4][3], dA23[4][3];
  float AR2[4][3];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      AR01[k][r] = pk2(0.f, 0.f); AR2[k][r] = 0.f;
      dA01[k][r] = pk2(0.f, 0.f); dA23[k][r] = pk2(0.f, 0.f);
    }
  {
    const float4* gsrc = reinterpret_cast<const float4*>(vrec + i0);
    for (int e = tid; e < TILE_F4; e += SK_THREADS) sconst[e] = __ldg(gsrc + e);
  }
  float nx[3 * GV];
  const float* vsrc = vpT + (int64_t)(3 * i0) * BP + b;      // walks 3 rows per vertex
#pragma unroll
  for (int q = 0; q < 3 * GV; q++) { nx[q] = *vsrc; vsrc += BP; }
  __syncthreads();

#pragma unroll 1
  for (int t = 0; t < NT; t++) {
    const float4* rt = sconst + (t & 1) * TILE_F4;
    const int it0 = i0 + t * VT;
    const bool has_next = t + 1 < NT;
    float4 pf0 = make_float4(0, 0, 0, 0), pf1 = pf0;
    if (has_next) {
      const float4* gsrc = reinterpret_cast<const float4*>(vrec + it0 + VT);
      pf0 = __ldg(gsrc + tid);
      if (tid + SK_THREADS < TILE_F4) pf1 = __ldg(gsrc + tid + SK_THREADS);
    }
    int voff[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int f = lane + 32 * q;
      const int v = USE_DV ? perm[it0 + f / 3] : -1;
      voff[q] = v >= 0 ? v * 3 + f % 3 : -1;
    }
    if (USE_DV) {
      // gather the natural-order dvertices rows of this tile into smem (pose-major -> thread-major)
      for (int r = warp; r < SK_THREADS; r += SK_THREADS / 32) {
        const int64_t bb = b0 + r;
        const float* src = dvertices + bb * (int64_t)(V * 3);
#pragma unroll
        for (int q = 0; q < 3; q++)
          tile_in[r * TILE_LD + lane + 32 * q] = (bb < B && voff[q] >= 0) ? src[voff[q]] : 0.f;
      }
      __syncthreads();
    }
#pragma unroll 1
    for (int sub = 0; sub < VT / GV; sub++) {
      float cur[3 * GV];
#pragma unroll
      for (int q = 0; q < 3 * GV; q++) cur[q] = nx[q];
      if (it0 + (sub + 1) * GV < i0 + VS_B) {
#pragma unroll
        for (int q = 0; q < 3 * GV; q++) { nx[q] = *vsrc; vsrc += BP; }
      }
      const float4* rh = rt + (sub * GV) * 7;
      float4 r0[GV];
      float w3[GV];
      uint32_t meta[GV], many = 0;
#pragma unroll
      for (int ii = 0; ii < GV; ii++) {
        r0[ii] = rh[ii * 7];
        const float4 r1 = rh[ii * 7 + 1];
        w3[ii] = r1.x;
        meta[ii] = __float_as_uint(r0[ii].x);
        many |= meta[ii];
      }
      const bool any_reload = (many >> 20) & 0xFu;
      // ---- per-vertex gradient dv
      float dv[GV][3];
#pragma unroll
      for (int ii = 0; ii < GV; ii++) { dv[ii][0] = 0.f; dv[ii][1] = 0.f; dv[ii][2] = 0.f; }
      if (USE_G && ((many >> 24) & 1u)) {
#pragma unroll
        for (int ii = 0; ii < GV; ii++) {
          f32x2 a[3] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
          for (int qq = 0; qq < JH_STRIDE / 4; qq++) {
            const float4 tt = rh[ii * 7 + 2 + qq];
            const f32x2 ja = pk2(tt.x, tt.y), jb = pk2(tt.z, tt.w);
#pragma unroll
            for (int c = 0; c < 3; c++) {
              if (2 * qq < 9) a[c] = fma2(ja, gp[USE_G ? c : 0][USE_G ? 2 * qq : 0], a[c]);
              if (2 * qq + 1 < 9) a[c] = fma2(jb, gp[USE_G ? c : 0][USE_G ? 2 * qq + 1 : 0], a[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float lo, hi;
            upk2(a[c], lo, hi);
            dv[ii][c] = lo + hi;
          }
        }
      }
      if (USE_DV) {
#pragma unroll
        for (int ii = 0; ii < GV; ii++)
#pragma unroll
          for (int c = 0; c < 3; c++) dv[ii][c] += tile_in[tid * TILE_LD + (sub * GV + ii) * 3 + c];
      }
      if (USE_X && ((many >> 26) & 1u)) {
#pragma unroll
        for (int ii = 0; ii < GV; ii++) {
          const float4 r1 = rh[ii * 7 + 1];
          const int xptr = __float_as_int(r1.y), xcnt = __float_as_int(r1.z);
          for (int q = 0; q < xcnt; q++) {
            const int src = __ldg(vx_src + xptr + q);
            const float cf = __ldg(vx_coef + xptr + q);
#pragma unroll
            for (int c = 0; c < 3; c++) dv[ii][c] = fmaf(cf, d30T[(int64_t)(src * 3 + c) * BP + b], dv[ii][c]);
          }
        }
      }
      // ---- dvp = sum_k w_k AR_k^T dv   and   dA_k += w_k dv (x) [vp;1]
#define JRR_BWD_VERTEX(ii)                                                                            \
      {                                                                                               \
        const float wk[4] = {r0[ii].y, r0[ii].z, r0[ii].w, w3[ii]};                                   \
        const f32x2 d0 = pk2(dv[ii][0], dv[ii][0]), d1 = pk2(dv[ii][1], dv[ii][1]),                   \
                    d2 = pk2(dv[ii][2], dv[ii][2]);                                                   \
        const f32x2 vp01 = pk2(cur[ii * 3 + 0], cur[ii * 3 + 1]), vp21 = pk2(cur[ii * 3 + 2], 1.f);   \
        f32x2 o01 = pk2(0.f, 0.f);                                                                    \
        float o2 = 0.f;                                                                               \
        _Pragma("unroll") for (int k = 0; k < 4; k++) {                                               \
          const f32x2 u01 = fma2(AR01[k][2], d2, fma2(AR01[k][1], d1, mul2(AR01[k][0], d0)));         \
          const float u2 = fmaf(AR2[k][2], dv[ii][2], fmaf(AR2[k][1], dv[ii][1], AR2[k][0] * dv[ii][0])); \
          const f32x2 ww = pk2(wk[k], wk[k]);                                                         \
          o01 = fma2(ww, u01, o01);                                                                   \
          o2 = fmaf(wk[k], u2, o2);                                                                   \
          _Pragma("unroll") for (int r = 0; r < 3; r++) {                                             \
            const float wd = wk[k] * dv[ii][r];                                                       \
            const f32x2 wdd = pk2(wd, wd);                                                            \
            dA01[k][r] = fma2(wdd, vp01, dA01[k][r]);                                                 \
            dA23[k][r] = fma2(wdd, vp21, dA23[k][r]);                                                 \
          }                                                                                           \
        }                                                                                             \
        float o0, o1;                                                                                 \
        upk2(o01, o0, o1);                                                                            \
        float* to = tile_out + tid * TILE_LD + (sub * GV + ii) * 3;                                   \
        to[0] = o0; to[1] = o1; to[2] = o2;                                                           \
      }
      if (!any_reload) {
#pragma unroll
        for (int ii = 0; ii < GV; ii++) JRR_BWD_VERTEX(ii)
      } else {
#pragma unroll
        for (int ii = 0; ii < GV; ii++) {
          const bool first = (meta[ii] >> 25) & 1u;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if ((meta[ii] >> (20 + k)) & 1u) {
              if (!first) {
                // this slot's joint changes: its accumulated dA leaves as one flush event
#pragma unroll
                for (int r = 0; r < 3; r++) {
                  float a0, a1, a2, a3;
                  upk2(dA01[k][r], a0, a1);
                  upk2(dA23[k][r], a2, a3);
                  flush_dst[(int64_t)(r * 4 + 0) * BP] = a0;
                  flush_dst[(int64_t)(r * 4 + 1) * BP] = a1;
                  flush_dst[(int64_t)(r * 4 + 2) * BP] = a2;
                  flush_dst[(int64_t)(r * 4 + 3) * BP] = a3;
                  dA01[k][r] = pk2(0.f, 0.f);
                  dA23[k][r] = pk2(0.f, 0.f);
                }
                flush_dst += 12 * BP;
              }
              const int j = (meta[ii] >> (5 * k)) & 31u;
              const float* src = AT + (int64_t)(j * 12) * BP + b;
#pragma unroll
              for (int r = 0; r < 3; r++) {
                const float a0 = src[(int64_t)(r * 4 + 0) * BP], a1 = src[(int64_t)(r * 4 + 1) * BP];
                AR01[k][r] = pk2(a0, a1);
                AR2[k][r] = src[(int64_t)(r * 4 + 2) * BP];
              }
            }
          }
          JRR_BWD_VERTEX(ii)
        }
      }
#undef JRR_BWD_VERTEX
    }
    __syncthreads();
    // dvp rows (A operand of the backward blend GEMM, K-major, tf32 hi/lo), coalesced
    for (int r = warp; r < SK_THREADS; r += SK_THREADS / 32) {
      const int64_t bb = b0 + r;
      float* dh = dvp_hi + bb * (int64_t)NP + (int64_t)it0 * 3;
      float* dl = dvp_lo + bb * (int64_t)NP + (int64_t)it0 * 3;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int f = lane + 32 * q;
        const float v = tile_out[r * TILE_LD + f];
        const float hi = tf32_hi_k(v);
        dh[f] = hi;
        dl[f] = tf32_hi_k(v - hi);
      }
    }
    if (has_next) {
      float4* dstc = sconst + ((t + 1) & 1) * TILE_F4;
      dstc[tid] = pf0;
      if (tid + SK_THREADS < TILE_F4) dstc[tid + SK_THREADS] = pf1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      float a0, a1, a2, a3;
      upk2(dA01[k][r], a0, a1);
      upk2(dA23[k][r], a2, a3);
      flush_dst[(int64_t)(r * 4 + 0) * BP] = a0;
      flush_dst[(int64_t)(r * 4 + 1) * BP] = a1;
      flush_dst[(int64_t)(r * 4 + 2) * BP] = a2;
      flush_dst[(int64_t)(r * 4 + 3) * BP] = a3;
    }
    flush_dst += 12 * BP;
  }
}

// dAT[k][e][b] = sum over the flush events of joint k (fixed order)
__global__ void __launch_bounds__(SK_THREADS)
dA_reduce_kernel(const int* __restrict__ flush_ptr, const int* __restrict__ flush_idx, int flush_limit,
                 const float* __restrict__ dAflush, int64_t BP, float* __restrict__ dAT, int accumulate) {
  const int64_t b = (int64_t)blockIdx.x * SK_THREADS + threadIdx.x;
  const int k = blockIdx.y;
  float acc[12];
#pragma unroll
  for (int e = 0; e < 12; e++) acc[e] = 0.f;
  const int p0 = flush_ptr[k], p1 = flush_ptr[k + 1];
  for (int p = p0; p < p1; p++) {
    const int f = __ldg(flush_idx + p);
    if (f >= flush_limit) break;      // ids ascend inside a joint's list; later ranges were not walked
    const float* src = dAflush + (int64_t)f * 12 * BP + b;
#pragma unroll
    for (int e = 0; e < 12; e++) acc[e] += src[(int64_t)e * BP];
  }
#pragma unroll
  for (int e = 0; e < 12; e++) {
    float* d = dAT + (int64_t)(k * 12 + e) * BP + b;
    *d = accumulate ? *d + acc[e] : acc[e];      // (later skinning passes add to the first)
  }
}

// ----------------------------------------------------------------------- module path: vertex un-packing
// The fused forward leaves the skinned vertices packed (sorted by joint set) and pose-contiguous, vT[3p+c][b]: that is the
// layout in which a warp of 32 poses writes full 128-byte lines.  The drop-in returns [B][6890][3] in the model's own vertex
// order.  This kernel walks the NATURAL order: a CTA takes 32 poses x 32 consecutive vertices, GATHERS their three rows each
// through inv_perm (every read is one full line of 32 poses), transposes in shared memory and writes 384 contiguous bytes per
// pose -- both directions fully coalesced, which a scatter by `perm` from the packed side cannot be (12-byte pieces).
constexpr int UP_V = 64;
__global__ void __launch_bounds__(256)
unpack_vertices_kernel(const int* __restrict__ inv_perm, const float* __restrict__ vT, int64_t B, int64_t BP,
                       float* __restrict__ out) {
  __shared__ float tile[UP_V * 3][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = blockIdx.x * UP_V;
  const int64_t b0 = (int64_t)blockIdx.y * 32;
  constexpr int RPW = UP_V * 3 / 8;               // 24 rows per warp: all gathers of a warp are in flight together
  float x[RPW];
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    const int v = v0 + r / 3;
    x[i] = v < V ? vT[(int64_t)(3 * __ldg(inv_perm + v) + r % 3) * BP + b0 + lane] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) tile[warp + 8 * i][lane] = x[i];
  __syncthreads();
  const int nv = min(UP_V, V - v0) * 3;          // floats of this vertex group per pose
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int pp = warp + 8 * q;
    const int64_t b = b0 + pp;
    if (b >= B) continue;
    float* o = out + b * (int64_t)(V * 3) + (int64_t)v0 * 3;
#pragma unroll
    for (int j = 0; j < UP_V * 3 / 32; j++) {
      const int e = j * 32 + lane;
      if (e < nv) o[e] = tile[e][pp];
    }
  }
}

int launch_unpack_vertices(const JrrModel* m, const Workspace& w, const float* vT, float* vertices_out, cudaStream_t st) {
  dim3 grid((V + UP_V - 1) / UP_V, (unsigned)((w.B + 31) / 32));
  unpack_vertices_kernel<<<grid, 256, 0, st>>>(m->inv_perm, vT, w.B, w.BP, vertices_out);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// Inverse of unpack_vertices_kernel for the module backward: the caller's d loss / d vertices [B][6890][3] is read in the
// natural order (coalesced, 768 bytes per pose), transposed in shared memory and written as full pose-contiguous lines
// into the packed layout dvT[3p+c][b] the fused backward walks.  The joints49 gradient that reaches vertices (21 vertex
// picks, 9 extra-regressor rows; d30T from joints49_bwd_kernel) is added here, so the backward kernel needs no special case.
__global__ void __launch_bounds__(256)
pack_dvertices_kernel(const int* __restrict__ inv_perm, const VtxRec* __restrict__ vrec, const int* __restrict__ vx_src,
                      const float* __restrict__ vx_coef, const float* __restrict__ dverts, const float* __restrict__ d30T,
                      int64_t B, int64_t BP, float* __restrict__ dvT) {
  __shared__ float tile[UP_V * 3][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v0 = blockIdx.x * UP_V;
  const int64_t b0 = (int64_t)blockIdx.y * 32;
  const int nv = min(UP_V, V - v0) * 3;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int pp = warp + 8 * q;
    const int64_t b = b0 + pp;
    const float* src = dverts != nullptr ? dverts + b * (int64_t)(V * 3) + (int64_t)v0 * 3 : nullptr;
#pragma unroll
    for (int j = 0; j < UP_V * 3 / 32; j++) {
      const int e = j * 32 + lane;
      tile[e][pp] = (src != nullptr && b < B && e < nv) ? src[e] : 0.f;
    }
  }
  __syncthreads();
  constexpr int RPW = UP_V * 3 / 8;
#pragma unroll 4
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    const int v = v0 + r / 3, c = r % 3;
    if (v >= V) continue;
    const int p = __ldg(inv_perm + v);
    float x = tile[r][lane];
    if (d30T != nullptr) {
      const int xp = vrec[p].xptr, xc = vrec[p].xcnt;
      for (int q = 0; q < xc; q++) x = fmaf(vx_coef[xp + q], d30T[(int64_t)(vx_src[xp + q] * 3 + c) * BP + b0 + lane], x);
    }
    dvT[(int64_t)(3 * p + c) * BP + b0 + lane] = x;
  }
}

int launch_pack_dvertices(const JrrModel* m, const Workspace& w, const float* dvertices, bool use_x, float* dvT, cudaStream_t st) {
  // the 22 padding vertices sit at the end of the packed order: their rows must be finite (their weights are zero)
  JRR_CUDA(cudaMemsetAsync(dvT + (int64_t)(3 * V) * w.BP, 0, sizeof(float) * (size_t)(3 * (VP - V)) * w.BP, st));
  dim3 grid((V + UP_V - 1) / UP_V, (unsigned)(w.BP / 32));
  pack_dvertices_kernel<<<grid, 256, 0, st>>>(m->inv_perm, m->vrec_b, m->vx_src, m->vx_coef, dvertices, use_x ? w.d30T : nullptr,
                                              w.B, w.BP, dvT);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// joints49 straight from the packed pose-contiguous vertices (same arithmetic as joints49_fwd_kernel)
__global__ void joints49_fwd_packed_kernel(const int* __restrict__ joint_map, const int* __restrict__ picks, Csr extra,
                                           const int* __restrict__ inv_perm, const float* __restrict__ Jp,
                                           const float* __restrict__ vT, int64_t B, int64_t BP, float* __restrict__ joints49) {
  // thread = (output joint o, pose b) with b fastest: the reads of one vertex row are pose-contiguous
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BP * JRR_NUM_OUT_JOINTS) return;
  const int o = (int)(idx / BP);
  const int64_t b = idx % BP;
  if (b >= B) return;
  const int src = joint_map[o];
  float out[3] = {0.f, 0.f, 0.f};
  if (src < NJ) {
    for (int c = 0; c < 3; c++) out[c] = Jp[b * 72 + src * 3 + c];
  } else if (src < NJ + JRR_NUM_PICKS) {
    const int p = inv_perm[picks[src - NJ]];
    for (int c = 0; c < 3; c++) out[c] = vT[(int64_t)(3 * p + c) * BP + b];
  } else {
    // eight entries at a time: their index -> permutation -> vertex loads are three dependent round trips per BATCH, not per
    // entry (a row of J_regressor_extra was a 16-deep chain of them: 21 us even for one pose); the sum keeps its order
    const int e = src - NJ - JRR_NUM_PICKS;
    const int q1 = extra.ptr[e + 1];
    for (int q = extra.ptr[e]; q < q1; q += 8) {
      int p[8];
      float cf[8], v[8][3];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const bool on = q + u < q1;
        cf[u] = on ? extra.val[q + u] : 0.f;
        p[u] = on ? extra.col[q + u] : extra.col[q];
      }
#pragma unroll
      for (int u = 0; u < 8; u++) p[u] = inv_perm[p[u]];
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int c = 0; c < 3; c++) v[u][c] = vT[(int64_t)(3 * p[u] + c) * BP + b];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (q + u < q1)
#pragma unroll
          for (int c = 0; c < 3; c++) out[c] = fmaf(cf[u], v[u][c], out[c]);
    }
  }
  for (int c = 0; c < 3; c++) joints49[(b * JRR_NUM_OUT_JOINTS + o) * 3 + c] = out[c];
}

int launch_joints49_fwd_packed(const JrrModel* m, const Workspace& w, const float* vT, float* joints49_out, cudaStream_t st) {
  const int64_t n = w.BP * JRR_NUM_OUT_JOINTS;
  joints49_fwd_packed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->joint_map, m->picks, m->extra, m->inv_perm, w.Jp,
                                                                         vT, w.B, w.BP, joints49_out);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// ----------------------------------------------------------------------- joints49 (module)
// smplx vertex_joint_selector + scripts/smpl.py:75-78
__global__ void joints49_fwd_kernel(const int* __restrict__ joint_map, const int* __restrict__ picks,
                                    Csr extra, const float* __restrict__ Jp,
                                    const float* __restrict__ vertices, int64_t B,
                                    float* __restrict__ joints49) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * JRR_NUM_OUT_JOINTS) return;
  const int64_t b = idx / JRR_NUM_OUT_JOINTS;
  const int o = (int)(idx % JRR_NUM_OUT_JOINTS);
  const int src = joint_map[o];
  float out[3] = {0.f, 0.f, 0.f};
  const float* vb = vertices + b * (int64_t)(V * 3);
  if (src < NJ) {
    for (int c = 0; c < 3; c++) out[c] = Jp[b * 72 + src * 3 + c];
  } else if (src < NJ + JRR_NUM_PICKS) {
    const int v = picks[src - NJ];
    for (int c = 0; c < 3; c++) out[c] = vb[v * 3 + c];
  } else {
    const int e = src - NJ - JRR_NUM_PICKS;
    for (int p = extra.ptr[e]; p < extra.ptr[e + 1]; p++) {
      const int v = extra.col[p];
      const float cf = extra.val[p];
      for (int c = 0; c < 3; c++) out[c] = fmaf(cf, vb[v * 3 + c], out[c]);
    }
  }
  for (int c = 0; c < 3; c++) joints49[idx * 3 + c] = out[c];
}

// gather the joints49 gradient back onto its 54 sources: first 24 -> dJp [BP][72], the 30
// vertex-borne sources -> d30T [90][BP] (consumed by skin_bwd USE_X)
__global__ void joints49_bwd_kernel(const int* __restrict__ joint_map, const float* __restrict__ dj49,
                                    int64_t B, int64_t BP, float* __restrict__ dJp,
                                    float* __restrict__ d30T) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BP * 54) return;
  const int src = (int)(idx / BP);
  const int64_t b = idx % BP;
  float d[3] = {0.f, 0.f, 0.f};
  if (b < B && dj49 != nullptr)
    for (int o = 0; o < JRR_NUM_OUT_JOINTS; o++)
      if (joint_map[o] == src)
        for (int c = 0; c < 3; c++) d[c] += dj49[(b * JRR_NUM_OUT_JOINTS + o) * 3 + c];
  if (src < NJ) {
    for (int c = 0; c < 3; c++) dJp[b * 72 + src * 3 + c] = d[c];
  } else {
    for (int c = 0; c < 3; c++) d30T[(int64_t)((src - NJ) * 3 + c) * BP + b] = d[c];
  }
}

// ---------------------------------------------------------------------------- loss finish
__global__ void loss_finish_kernel(const LossFinishArgs a) { loss_finish_warp(a, threadIdx.x); }

// ---------------------------------------------------------------------------- camera fit
// optimize.py:187-199: Adam(lr) on the camera translation alone against the 2-D joints.  The 3-D
// joints do not depend on the camera, so they are computed once and each frame iterates privately.
__global__ void camera_fit_kernel(const float* __restrict__ joints17, const float* __restrict__ gt2d,
                                  float* __restrict__ cam, int64_t B, int iters, float lr, float scale,
                                  float* __restrict__ loss_part) {
  __shared__ float red[4];
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  float last = 0.f;
  if (b < B) {
    float pred[NACC], T[3], m3[3] = {0.f, 0.f, 0.f}, v3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < NACC; a++) pred[a] = joints17[b * NACC + a];
#pragma unroll
    for (int c = 0; c < 3; c++) T[c] = cam[b * 3 + c];
    for (int t = 1; t <= iters; t++) {
      float dT[3] = {0.f, 0.f, 0.f};
      last = proj2d_grad(pred, T, gt2d + b * 34, scale, nullptr, dT);
      adam_update3(T, dT, m3, v3, t, lr);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) cam[b * 3 + c] = T[c];
  }
  for (int o = 16; o > 0; o >>= 1) last += __shfl_xor_sync(0xffffffffu, last, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = last;
  __syncthreads();
  if (threadIdx.x == 0 && loss_part != nullptr) loss_part[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void sum_scale_kernel(const float* __restrict__ parts, int n, float scale, float* __restrict__ out) {
  const int lane = threadIdx.x;
  float a = 0.f;
  for (int i = lane; i < n; i += 32) a += parts[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[0] = a * scale;
}

// ---------------------------------------------------------------------------- host side
int launch_skin_fwd(const JrrModel* m, const Workspace& w, float* vertices_out, float* vT_out,
                    bool want_part, cudaStream_t st) {
  dim3 grid((unsigned)(w.BP / SK_THREADS), NSPLIT), block(SK_THREADS);
  const size_t cbytes = 2 * TILE_F4 * sizeof(float4);
#define JRR_SF(WV, WVT, P)                                                                       \
  do {                                                                                           \
    auto kern = skin_fwd_kernel<WV, WVT, P>;                                                     \
    const size_t smem = cbytes + (WV ? (size_t)SK_THREADS * TILE_LD * sizeof(float) : 0);        \
    cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e_ != cudaSuccess) return fail(JRR_ERR_CUDA, cudaGetErrorString(e_));                    \
    kern<<<grid, block, smem, st>>>(m->vrec, m->perm, w.AT, w.vpT, w.B, w.BP, vertices_out,      \
                                    vT_out, w.part);                                             \
  } while (0)
  const bool wv = vertices_out != nullptr, wvt = vT_out != nullptr;
  if (wv && !wvt && !want_part) JRR_SF(true, false, false);
  else if (wv && !wvt && want_part) JRR_SF(true, false, true);
  else if (!wv && !wvt && want_part) JRR_SF(false, false, true);
  else if (!wv && wvt && want_part) JRR_SF(false, true, true);
  else if (!wv && wvt && !want_part) JRR_SF(false, true, false);
  else return fail(JRR_ERR_INVALID, "skin_fwd: unsupported output combination");
#undef JRR_SF
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// poses per CTA of the loss-seed kernel: 32 while the partial count fits its region of the partial buffer
static inline int loss_seed_ppb(int64_t BP) { return BP / 32 <= LOSS_PART_2D ? 32 : 128; }

// upstream gradient of the 17 regressed joints [B,17,3] -> the pose-contiguous seed gT [51][BP] the skinning backward reads
// (rows of padding poses are zero): the backward of find_joints on its own (jrr_find_joints_backward)
__global__ void seed_from_dpred_kernel(const float* __restrict__ dpred, int64_t B, int64_t BP, float* __restrict__ gT) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NACC * BP) return;
  const int a = (int)(idx / BP);
  const int64_t b = idx % BP;
  gT[idx] = b < B ? dpred[b * NACC + a] : 0.f;
}

int launch_seed_from_dpred(const Workspace& w, const float* dpred, cudaStream_t st) {
  const int64_t n = NACC * w.BP;
  seed_from_dpred_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dpred, w.B, w.BP, w.gT);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_loss_seed(const JrrModel* m, Workspace& w, bool fused_partials, const float* gt_mm,
                     int64_t B_logical, float w_joint, float* joints17_out, const Proj2D& p2d, cudaStream_t st) {
  const int ppb = loss_seed_ppb(w.BP);
  w.n_joint_part = (int)(w.BP / ppb);
  dim3 grid((unsigned)(w.BP / ppb)), block(LS_THREADS);
  const float scale = gt_mm != nullptr ? w_joint * 2.f / (51.f * (float)B_logical) : 0.f;
  const FusedSched sc = fused_fwd_sched(m, w.BP, m->nv_act);
  const int n_tiles = sc.n_tiles, T = sc.T, G = sc.G;
  if (ppb == 32)
    loss_seed_kernel<32><<<grid, block, 0, st>>>(w.part, fused_partials ? 0 : NSPLIT, n_tiles, T, G, sc.mdiv, gt_mm, w.B, w.BP, scale,
                                                 gt_mm != nullptr ? w.gT : nullptr, joints17_out, w.loss_part, p2d,
                                                 fused_partials ? m->n_pass : 1, w.part_stride);
  else
    loss_seed_kernel<128><<<grid, block, 0, st>>>(w.part, fused_partials ? 0 : NSPLIT, n_tiles, T, G, sc.mdiv, gt_mm, w.B, w.BP, scale,
                                                  gt_mm != nullptr ? w.gT : nullptr, joints17_out, w.loss_part, p2d,
                                                  fused_partials ? m->n_pass : 1, w.part_stride);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_skin_bwd(const JrrModel* m, const Workspace& w, const float* dvertices, bool use_g,
                    bool use_x, cudaStream_t st) {
  dim3 grid((unsigned)(w.BP / SK_THREADS), NSPLIT_B), block(SK_THREADS);
  const bool use_dv = dvertices != nullptr;
  const size_t smem = 2 * TILE_F4 * sizeof(float4) + (size_t)SK_THREADS * TILE_LD * sizeof(float) * (use_dv ? 2 : 1);
#define JRR_SB(G, DV, X)                                                                          \
  do {                                                                                            \
    auto kern = skin_bwd_kernel<G, DV, X>;                                                        \
    cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e_ != cudaSuccess) return fail(JRR_ERR_CUDA, cudaGetErrorString(e_));                     \
    kern<<<grid, block, smem, st>>>(m->vrec_b, m->perm, m->range_flush_base, m->vx_src,           \
                                    m->vx_coef, w.AT, w.vpT, w.gT, dvertices, w.d30T, w.B, w.BP,  \
                                    w.dvp_hi, w.dvp_lo, w.dAflush);                               \
  } while (0)
  if (use_g && !use_dv && !use_x) JRR_SB(true, false, false);
  else if (!use_g && use_dv && use_x) JRR_SB(false, true, true);
  else if (!use_g && use_dv && !use_x) JRR_SB(false, true, false);
  else if (!use_g && !use_dv && use_x) JRR_SB(false, false, true);
  else return fail(JRR_ERR_INVALID, "skin_bwd: unsupported input combination");
#undef JRR_SB
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_dA_reduce(const JrrModel* m, const Workspace& w, int lists, cudaStream_t st) {
  dim3 grid((unsigned)(w.BP / SK_THREADS), NJ), block(SK_THREADS);
  const float* src = w.dAflush + (int64_t)m->flush_off[m->cur_pass] * 12 * w.BP;      // this pass's flush region
  const int acc = m->cur_pass > 0 ? 1 : 0;
  if (lists == 1)
    dA_reduce_kernel<<<grid, block, 0, st>>>(m->flush_ptr_l, m->flush_idx_l, m->n_flush_l, src, w.BP, w.dAT, acc);
  else if (lists == 2)
    dA_reduce_kernel<<<grid, block, 0, st>>>(m->flush_ptr_s, m->flush_idx_s, m->n_flush_s, src, w.BP, w.dAT, acc);
  else
    dA_reduce_kernel<<<grid, block, 0, st>>>(m->flush_ptr, m->flush_idx, m->n_flush, src, w.BP, w.dAT, acc);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_joints49_fwd(const JrrModel* m, const Workspace& w, const float* vertices,
                        float* joints49_out, cudaStream_t st) {
  const int64_t n = w.B * JRR_NUM_OUT_JOINTS;
  joints49_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->joint_map, m->picks, m->extra,
                                                                  w.Jp, vertices, w.B, joints49_out);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_joints49_bwd(const JrrModel* m, const Workspace& w, const float* djoints49,
                        cudaStream_t st) {
  const int64_t n = w.BP * 54;
  joints49_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->joint_map, djoints49, w.B, w.BP,
                                                                  w.dJp, w.d30T);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// Folded loss path: joints (+ loss seed, dA, dQ when gt_mm is given) from Q = feat . T^T.
int launch_folded_seed(const JrrModel* m, Workspace& w, const float* gt_mm, int64_t B_logical, float w_joint,
                       float* joints17_out, const Proj2D& p2d, cudaStream_t st, float* dc_part, bool plain_dq) {
  const float scale = gt_mm != nullptr ? w_joint * 2.f / (51.f * (float)B_logical) : 0.f;
  const bool grad = gt_mm != nullptr;
  w.n_joint_part = (int)(w.BP / FS_POSES);
  constexpr int smem = (FS_WARPS * FS_POSES * FS_LD + NACC * FS_POSES + NJ * NH) * (int)sizeof(float);
  JRR_CUDA(cudaFuncSetAttribute(folded_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  folded_seed_kernel<<<(unsigned)(w.BP / FS_POSES), FS_WARPS * 32, smem, st>>>(
      w.vpT, w.AT, m->Tc, gt_mm, w.B, w.BP, scale, p2d, w.loss_part, joints17_out, grad ? w.dAT : nullptr,
      grad ? w.dvp_hi : nullptr, (grad && !plain_dq) ? w.dvp_lo : nullptr, grad ? dc_part : nullptr);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

LossFinishArgs make_loss_finish_args(const Workspace& w, int64_t B_logical, float w_joint, float w_pose, bool have_pose,
                                     float w_2d, float w_shape, float* loss_out, float* loss_accum) {
  LossFinishArgs a{};
  a.lp_joint = w.loss_part; a.n_joint = w.n_joint_part; a.sj = 1.f / (51.f * (float)B_logical);
  a.lp_pose = w.loss_part + LOSS_PART_POSE; a.n_pose = have_pose ? w.n_pose_part : 0; a.sp = 1.f / (25.f * (float)B_logical);
  a.lp_2d = w.loss_part + LOSS_PART_2D; a.s2 = 1.f / (34.f * (float)B_logical);
  a.lp_shape = w.shape_part; a.n_shape = (int)(w.BP / 128); a.ss = 1.f / (float)B_logical;
  a.wj = w_joint; a.wp = w_pose; a.w2 = w_2d; a.wsh = w_shape;
  a.loss_out = loss_out; a.loss_accum = loss_accum;
  return a;
}

int launch_loss_finish(const Workspace& w, int64_t B_logical, float w_joint, float w_pose,
                       bool have_pose, float w_2d, float w_shape, float* loss_out, float* loss_accum, cudaStream_t st) {
  loss_finish_kernel<<<1, 32, 0, st>>>(make_loss_finish_args(w, B_logical, w_joint, w_pose, have_pose, w_2d, w_shape,
                                                            loss_out, loss_accum));
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_camera_fit(const Workspace& w, const float* joints17, const float* gt2d, float* cam, int iters, float lr,
                      int64_t B_logical, float* loss_out, cudaStream_t st) {
  const unsigned nblk = (unsigned)((w.B + 127) / 128);
  camera_fit_kernel<<<nblk, 128, 0, st>>>(joints17, gt2d, cam, w.B, iters, lr, 2.f / (34.f * (float)B_logical),
                                          loss_out ? w.loss_part + LOSS_PART_2D : nullptr);
  JRR_LAUNCH_CHECK();
  if (loss_out) {
    sum_scale_kernel<<<1, 32, 0, st>>>(w.loss_part + LOSS_PART_2D, (int)nblk, 1.f / (34.f * (float)B_logical), loss_out);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

}  // namespace jrr
