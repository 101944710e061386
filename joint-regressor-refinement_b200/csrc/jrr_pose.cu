// Warp-per-pose kernels: rotation decode (rot6d / Rodrigues / rotmat), rest joints
// J = J0 + (J24.S) beta, the 24-joint kinematic chain walked level by level with warp
// shuffles (lane = joint), and the analytic reverse walk fused with the per-pose Adam update.
//
// Replaces (file:line under /root/reference): scripts/utils.py:190-204 rot6d_to_rotmat,
// smplx.lbs.{batch_rodrigues, vertices2joints, batch_rigid_transform} reached through
// scripts/smpl.py:72-74, their autograd backward (scripts/optimize.py:264) and
// torch.optim.Adam.step (scripts/optimize.py:201-202,265).
#include <algorithm>
#include <cstdlib>

#include "jrr_internal.cuh"

namespace jrr {

constexpr int POSES_PER_CTA = 8;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- rotation decoders -------------------------------------------------------------------
__device__ __forceinline__ void rot6d_decode(const float x[6], float R[9]) {
  // utils.py:198-204: a1 = even entries, a2 = odd entries, columns b1,b2,b3
  float a1x = x[0], a1y = x[2], a1z = x[4];
  float a2x = x[1], a2y = x[3], a2z = x[5];
  float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  float s = b1x * a2x + b1y * a2y + b1z * a2z;
  float ux = a2x - s * b1x, uy = a2y - s * b1y, uz = a2z - s * b1z;
  float n2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);
  float b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  float b3x = b1y * b2z - b1z * b2y;
  float b3y = b1z * b2x - b1x * b2z;
  float b3z = b1x * b2y - b1y * b2x;
  R[0] = b1x; R[1] = b2x; R[2] = b3x;
  R[3] = b1y; R[4] = b2y; R[5] = b3y;
  R[6] = b1z; R[7] = b2z; R[8] = b3z;
}

// backward of rot6d_decode: dR (row-major 3x3) -> dx[6]
__device__ __forceinline__ void rot6d_backward(const float x[6], const float dR[9], float dx[6]) {
  float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
  float n1r = sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
  float n1 = fmaxf(n1r, 1e-12f);
  float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  float s = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  float u[3] = {a2[0] - s * b1[0], a2[1] - s * b1[1], a2[2] - s * b1[2]};
  float n2r = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  float n2 = fmaxf(n2r, 1e-12f);
  float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  float db1[3] = {dR[0], dR[3], dR[6]};
  float db2[3] = {dR[1], dR[4], dR[7]};
  float db3[3] = {dR[2], dR[5], dR[8]};
  // b3 = b1 x b2
  db1[0] += b2[1] * db3[2] - b2[2] * db3[1];
  db1[1] += b2[2] * db3[0] - b2[0] * db3[2];
  db1[2] += b2[0] * db3[1] - b2[1] * db3[0];
  db2[0] += db3[1] * b1[2] - db3[2] * b1[1];
  db2[1] += db3[2] * b1[0] - db3[0] * b1[2];
  db2[2] += db3[0] * b1[1] - db3[1] * b1[0];
  // b2 = u / max(|u|, eps)
  float du[3];
  if (n2r > 1e-12f) {
    float d = b2[0] * db2[0] + b2[1] * db2[1] + b2[2] * db2[2];
    for (int i = 0; i < 3; i++) du[i] = (db2[i] - b2[i] * d) / n2;
  } else {
    for (int i = 0; i < 3; i++) du[i] = db2[i] / n2;
  }
  // u = a2 - s*b1, s = b1.a2
  float ds = -(b1[0] * du[0] + b1[1] * du[1] + b1[2] * du[2]);
  float da2[3];
  for (int i = 0; i < 3; i++) {
    da2[i] = du[i] + ds * b1[i];
    db1[i] += -s * du[i] + ds * a2[i];
  }
  float da1[3];
  if (n1r > 1e-12f) {
    float d = b1[0] * db1[0] + b1[1] * db1[1] + b1[2] * db1[2];
    for (int i = 0; i < 3; i++) da1[i] = (db1[i] - b1[i] * d) / n1;
  } else {
    for (int i = 0; i < 3; i++) da1[i] = db1[i] / n1;
  }
  dx[0] = da1[0]; dx[2] = da1[1]; dx[4] = da1[2];
  dx[1] = da2[0]; dx[3] = da2[1]; dx[5] = da2[2];
}

__device__ __forceinline__ void rodrigues_decode(const float r[3], float R[9]) {
  // smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, n = r/angle
  float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  float th = sqrtf(ex * ex + ey * ey + ez * ez);
  float x = r[0] / th, y = r[1] / th, z = r[2] / th;
  float s, c;
  sincosf(th, &s, &c);
  float oc = 1.f - c;
  // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K^2 = n n^T - |n|^2 I
  float nn = x * x + y * y + z * z;
  R[0] = 1.f + oc * (x * x - nn); R[1] = -s * z + oc * x * y;     R[2] = s * y + oc * x * z;
  R[3] = s * z + oc * x * y;      R[4] = 1.f + oc * (y * y - nn); R[5] = -s * x + oc * y * z;
  R[6] = -s * y + oc * x * z;     R[7] = s * x + oc * y * z;      R[8] = 1.f + oc * (z * z - nn);
}

__device__ __forceinline__ void rodrigues_backward(const float r[3], const float dR[9], float dr[3]) {
  float e[3] = {r[0] + 1e-8f, r[1] + 1e-8f, r[2] + 1e-8f};
  float th = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  float n[3] = {r[0] / th, r[1] / th, r[2] / th};
  float s, c;
  sincosf(th, &s, &c);
  float oc = 1.f - c;
  float K[9] = {0.f, -n[2], n[1], n[2], 0.f, -n[0], -n[1], n[0], 0.f};
  float K2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      K2[i * 3 + j] = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
  // R = I + s K + oc K^2
  float dth = 0.f;
  for (int i = 0; i < 9; i++) dth += dR[i] * (c * K[i] + s * K2[i]);
  // dK = s dR + oc (dR K^T + K^T dR)
  float dK[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < 3; k++) {
        a += dR[i * 3 + k] * K[j * 3 + k];   // dR K^T
        b += K[k * 3 + i] * dR[k * 3 + j];   // K^T dR
      }
      dK[i * 3 + j] = s * dR[i * 3 + j] + oc * (a + b);
    }
  float dn[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
  // n = r/th, th = |r + eps|
  float ndn = n[0] * dn[0] + n[1] * dn[1] + n[2] * dn[2];
  float dth_tot = dth - ndn / th;
  for (int i = 0; i < 3; i++) dr[i] = dn[i] / th + dth_tot * e[i] / th;
}

template <int KIND>
__device__ __forceinline__ void decode_rot(const float* __restrict__ pose, int64_t b, int j, bool valid,
                                           float raw[9], float R[9]) {
  if (!valid) {
    for (int i = 0; i < 9; i++) { raw[i] = 0.f; R[i] = (i % 4 == 0) ? 1.f : 0.f; }
    if (KIND == JRR_POSE_ROT6D) { raw[0] = 1.f; raw[3] = 1.f; }
    if (KIND == JRR_POSE_ROTMAT) for (int i = 0; i < 9; i++) raw[i] = R[i];
    return;
  }
  if (KIND == JRR_POSE_ROT6D) {
    const float* p = pose + (b * NJ + j) * 6;
    for (int i = 0; i < 6; i++) raw[i] = p[i];
    rot6d_decode(raw, R);
  } else if (KIND == JRR_POSE_AXIS_ANGLE) {
    const float* p = pose + (b * NJ + j) * 3;
    for (int i = 0; i < 3; i++) raw[i] = p[i];
    rodrigues_decode(raw, R);
  } else {
    const float* p = pose + (b * NJ + j) * 9;
    for (int i = 0; i < 9; i++) { raw[i] = p[i]; R[i] = raw[i]; }
  }
}

// Forward chain for lane j: fills GR (world rotation), Gt (world translation = posed joint),
// Jr (rest joint), and GRp (parent's world rotation; identity for the root).
__device__ __forceinline__ void chain_forward(const ChainTab& tab, int j, const float R[9],
                                              const float Jr[3], float GR[9], float Gt[3],
                                              float GRp[9], float rel[3]) {
  const int par = tab.parent[j];
  const int src = par < 0 ? 0 : par;
  const int dep = tab.depth[j];
  float Jp[3];
  for (int i = 0; i < 3; i++) Jp[i] = __shfl_sync(FULL, Jr[i], src);
  for (int i = 0; i < 3; i++) rel[i] = par < 0 ? Jr[i] : Jr[i] - Jp[i];
  for (int i = 0; i < 9; i++) { GR[i] = R[i]; GRp[i] = (i % 4 == 0) ? 1.f : 0.f; }
  for (int i = 0; i < 3; i++) Gt[i] = rel[i];
  for (int d = 1; d <= tab.max_depth; d++) {
    float PR[9], Pt[3];
    for (int i = 0; i < 9; i++) PR[i] = __shfl_sync(FULL, GR[i], src);
    for (int i = 0; i < 3; i++) Pt[i] = __shfl_sync(FULL, Gt[i], src);
    if (dep == d) {
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
          GR[r * 3 + c] = PR[r * 3 + 0] * R[0 * 3 + c] + PR[r * 3 + 1] * R[1 * 3 + c] + PR[r * 3 + 2] * R[2 * 3 + c];
        Gt[r] = PR[r * 3 + 0] * rel[0] + PR[r * 3 + 1] * rel[1] + PR[r * 3 + 2] * rel[2] + Pt[r];
      }
      for (int i = 0; i < 9; i++) GRp[i] = PR[i];
    }
  }
}

__device__ __forceinline__ void rest_joint(const float* __restrict__ J0, const float* __restrict__ JS,
                                           const float beta[NB], int j, float Jr[3]) {
  for (int c = 0; c < 3; c++) {
    float a = J0[j * 3 + c];
    const float* s = JS + (j * 3 + c) * NB;
    for (int l = 0; l < NB; l++) a = fmaf(s[l], beta[l], a);
    Jr[c] = a;
  }
}

template <int KIND>
__global__ void __launch_bounds__(POSES_PER_CTA * 32)
pose_fwd_kernel(const __grid_constant__ ChainTab tab, const float* __restrict__ J0,
                const float* __restrict__ JS, const float* __restrict__ betas,
                const float* __restrict__ pose, int64_t B, int64_t BP, float* __restrict__ AT,
                float* __restrict__ feat_hi, float* __restrict__ feat_lo, float* __restrict__ Jp_out) {
  __shared__ float sA[NJ * 12][POSES_PER_CTA];
  __shared__ float sF[POSES_PER_CTA][KA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * POSES_PER_CTA + warp;
  const bool valid = b < B;
  const int j = lane < NJ ? lane : NJ - 1;

  float beta[NB];
  for (int l = 0; l < NB; l++) beta[l] = valid ? betas[b * NB + l] : 0.f;
  float raw[9], R[9], Jr[3], GR[9], Gt[3], GRp[9], rel[3];
  decode_rot<KIND>(pose, b, j, valid, raw, R);
  rest_joint(J0, JS, beta, j, Jr);
  chain_forward(tab, j, R, Jr, GR, Gt, GRp, rel);

  for (int i = lane; i < KA; i += 32) sF[warp][i] = 0.f;
  __syncwarp();
  if (lane < NJ) {
    for (int r = 0; r < 3; r++) {
      float t = Gt[r] - (GR[r * 3 + 0] * Jr[0] + GR[r * 3 + 1] * Jr[1] + GR[r * 3 + 2] * Jr[2]);
      sA[lane * 12 + r * 4 + 0][warp] = GR[r * 3 + 0];
      sA[lane * 12 + r * 4 + 1][warp] = GR[r * 3 + 1];
      sA[lane * 12 + r * 4 + 2][warp] = GR[r * 3 + 2];
      sA[lane * 12 + r * 4 + 3][warp] = t;
    }
    if (lane >= 1)
      for (int i = 0; i < 9; i++) sF[warp][(lane - 1) * 9 + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
    if (Jp_out != nullptr && valid)
      for (int r = 0; r < 3; r++) Jp_out[b * 72 + lane * 3 + r] = Gt[r];
  }
  if (lane < NB) sF[warp][FEAT_BETA + lane] = beta[lane];
  if (lane == 0) sF[warp][FEAT_ONE] = 1.f;
  __syncthreads();
  // feature rows (hi/lo split for the 3xTF32 blend GEMM), coalesced
  if (b < BP) {
    for (int i = lane; i < KA; i += 32) {
      float f = sF[warp][i];
      if (feat_lo == nullptr) { feat_hi[b * KA + i] = f; continue; }     // plain fp32 (A-through-TMEM GEMM)
      float hi = tf32_hi(f);
      feat_hi[b * KA + i] = hi;
      feat_lo[b * KA + i] = tf32_hi(f - hi);
    }
  }
  // transforms, pose contiguous: each thread stores one full 32-byte sector
  const int64_t b0 = (int64_t)blockIdx.x * POSES_PER_CTA;
  for (int e = threadIdx.x; e < NJ * 12; e += blockDim.x) {
    float4 v0 = make_float4(sA[e][0], sA[e][1], sA[e][2], sA[e][3]);
    float4 v1 = make_float4(sA[e][4], sA[e][5], sA[e][6], sA[e][7]);
    float4* dst = reinterpret_cast<float4*>(AT + (int64_t)e * BP + b0);
    dst[0] = v0;
    dst[1] = v1;
  }
}

// ---- backward + Adam ----------------------------------------------------------------------
// One warp per pose, lane = joint.  Inputs (all produced earlier in the same step):
//   dAT   [288][BP]           d loss / d relative transform (from the skinning backward)
//   dfeat [KSPLIT][BP][224]   split-K partials of d loss / d blend features
//   dJp   [BP][72]            d loss / d posed joints (module path only)
//   dx6c  [BP][144]           critic gradient w.r.t. rot6d (refine path only)
template <int KIND, bool ADAM>
__global__ void __launch_bounds__(POSES_PER_CTA * 32, 4)   // <= 64 registers: all 4096 warps of a 4096-pose step in one wave
pose_bwd_kernel(const __grid_constant__ ChainTab tab, const float* __restrict__ J0,
                const float* __restrict__ JS, const float* betas /* may alias betas_rw */,
                const float* pose /* may alias x6_rw */, int64_t B, int64_t BP, const float* __restrict__ dAT,
                const float* __restrict__ dfeat, int ksplit, const float* __restrict__ dJp,
                const float* __restrict__ dx6c, const float* __restrict__ dbeta_s, float* __restrict__ dbetas_out,
                float* __restrict__ dpose_out, float* x6_rw, float* betas_rw,
                float* __restrict__ adam_m, float* __restrict__ adam_v,
                const int32_t* __restrict__ step_count, float lr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * POSES_PER_CTA + warp;
  if (b >= B) return;  // whole warp exits together
  const int j = lane < NJ ? lane : NJ - 1;
  const bool act = lane < NJ;

  float beta[NB];
  for (int l = 0; l < NB; l++) beta[l] = betas[b * NB + l];
  float raw[9], R[9], Jr[3], GR[9], Gt[3], GRp[9], rel[3];
  decode_rot<KIND>(pose, b, j, true, raw, R);
  rest_joint(J0, JS, beta, j, Jr);
  chain_forward(tab, j, R, Jr, GR, Gt, GRp, rel);

  // upstream grads of the relative transform A_j = [GR_j | Gt_j - GR_j Jr_j]
  float dGR[9], dGt[3], dJr[3];
  {
    float dA[12];
    for (int e = 0; e < 12; e++) dA[e] = act ? dAT[(int64_t)(j * 12 + e) * BP + b] : 0.f;
    for (int r = 0; r < 3; r++) {
      float dt = dA[r * 4 + 3];
      dGt[r] = dt + ((dJp != nullptr && act) ? dJp[b * 72 + j * 3 + r] : 0.f);
      for (int c = 0; c < 3; c++) dGR[r * 3 + c] = dA[r * 4 + c] - dt * Jr[c];
    }
    // At = Gt - GR Jr  ->  dJr += -GR^T dAt
    for (int c = 0; c < 3; c++)
      dJr[c] = -(GR[0 * 3 + c] * dA[3] + GR[1 * 3 + c] * dA[7] + GR[2 * 3 + c] * dA[11]);
  }
  // reverse walk, deepest level first.  A joint at depth d sends its parent
  //   mR = dGR R^T + dGt rel^T (9), mt = dGt (3), mJ = -GRp^T dGt (3)
  const int dep = tab.depth[j];
  int chl[MAXCH];      // this lane's children, read once (a lane-indexed constant load serialises per address)
#pragma unroll
  for (int ci = 0; ci < MAXCH; ci++) chl[ci] = tab.child[j][ci];
  for (int d = tab.max_depth; d >= 1; d--) {
    float msg[15];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        msg[r * 3 + c] = dGR[r * 3 + 0] * R[c * 3 + 0] + dGR[r * 3 + 1] * R[c * 3 + 1] +
                         dGR[r * 3 + 2] * R[c * 3 + 2] + dGt[r] * rel[c];
    float gpt[3];
    for (int c = 0; c < 3; c++)
      gpt[c] = GRp[0 * 3 + c] * dGt[0] + GRp[1 * 3 + c] * dGt[1] + GRp[2 * 3 + c] * dGt[2];
    for (int r = 0; r < 3; r++) { msg[9 + r] = dGt[r]; msg[12 + r] = -gpt[r]; }
    if (act && dep == d)
      for (int c = 0; c < 3; c++) dJr[c] += gpt[c];   // rel_j = Jr_j - Jr_parent
    const int rounds = tab.maxch[d];   // children of one joint sit in its first slots
#pragma unroll
    for (int ci = 0; ci < MAXCH; ci++) {
      if (ci < rounds) {               // uniform
        const int ch = chl[ci];
        const int src = ch < 0 ? 0 : ch;
        const bool take = act && ch >= 0 && dep + 1 == d;   // a child sits one level below its parent
#pragma unroll
        for (int i = 0; i < 15; i++) {
          float v = __shfl_sync(FULL, msg[i], src);
          if (take) {
            if (i < 9) dGR[i] += v;
            else if (i < 12) dGt[i - 9] += v;
            else dJr[i - 12] += v;
          }
        }
      }
    }
  }
  // root: Gt_0 = Jr_0
  if (lane == 0)
    for (int c = 0; c < 3; c++) dJr[c] += dGt[c];
  // local rotation gradient dR_j = GRp^T dGR_j (+ blend-feature gradient for j >= 1)
  float dR[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      dR[r * 3 + c] = GRp[0 * 3 + r] * dGR[0 * 3 + c] + GRp[1 * 3 + r] * dGR[1 * 3 + c] + GRp[2 * 3 + r] * dGR[2 * 3 + c];
  if (act && lane >= 1)
    for (int s = 0; s < ksplit; s++) {
      const float* df = dfeat + ((int64_t)s * BP + b) * KA + (lane - 1) * 9;
      for (int i = 0; i < 9; i++) dR[i] += df[i];
    }
  // d beta_l = sum_j JS_j[:,l] . dJr_j  + dfeat[207+l]
  float dbeta = 0.f;
  for (int l = 0; l < NB; l++) {
    float p = 0.f;
    if (act)
      for (int c = 0; c < 3; c++) p += JS[(j * 3 + c) * NB + l] * dJr[c];
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
    if (lane == l) dbeta = p;
  }
  if (lane < NB)
    for (int s = 0; s < ksplit; s++) dbeta += dfeat[((int64_t)s * BP + b) * KA + FEAT_BETA + lane];
  if (lane < NB && dbeta_s != nullptr) dbeta += dbeta_s[b * NB + lane];   // shape-critic term

  // rotation parameter gradient
  float dp[9];
  int np;
  if (KIND == JRR_POSE_ROT6D) { rot6d_backward(raw, dR, dp); np = 6; }
  else if (KIND == JRR_POSE_AXIS_ANGLE) { rodrigues_backward(raw, dR, dp); np = 3; }
  else { for (int i = 0; i < 9; i++) dp[i] = dR[i]; np = 9; }

  if (!ADAM) {
    if (act)
      for (int i = 0; i < np; i++) dpose_out[(b * NJ + j) * np + i] = dp[i];
    if (lane < NB) dbetas_out[b * NB + lane] = dbeta;
  } else {
    // torch.optim.Adam (no weight decay / amsgrad), optimize.py:201-202,265
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const int t = *step_count + 1;
    // bias corrections in double, like the Python scalars of torch.optim.Adam
    const float bc2s = (float)sqrt(1.0 - pow(0.999, (double)t));
    const float step = (float)((double)lr / (1.0 - pow(0.9, (double)t)));
    if (act) {
      for (int i = 0; i < 6; i++) {
        float g = dp[i] + (dx6c != nullptr ? dx6c[b * 144 + j * 6 + i] : 0.f);
        int64_t pi = b * NPARAM + j * 6 + i;
        float m = b1 * adam_m[pi] + (1.f - b1) * g;
        float v = b2 * adam_v[pi] + (1.f - b2) * g * g;
        adam_m[pi] = m;
        adam_v[pi] = v;
        float denom = sqrtf(v) / bc2s + eps;
        x6_rw[(b * NJ + j) * 6 + i] = raw[i] - step * (m / denom);
      }
    }
    if (lane < NB) {
      int64_t pi = b * NPARAM + 144 + lane;
      float g = dbeta;
      float m = b1 * adam_m[pi] + (1.f - b1) * g;
      float v = b2 * adam_v[pi] + (1.f - b2) * g * g;
      adam_m[pi] = m;
      adam_v[pi] = v;
      float denom = sqrtf(v) / bc2s + eps;
      betas_rw[b * NB + lane] = beta[lane] - step * (m / denom);
    }
  }
}

__global__ void bump_step_kernel(int32_t* step_count) { *step_count += 1; }

// Adam step on the parameter gradients left by pose_bwd_kernel<ROT6D, false> plus the critics' input
// gradients: lets the chain backward run beside the critic branch, only this element-wise kernel
// waits for it.  Same arithmetic, in the same order, as the fused Adam of pose_bwd_kernel.
__global__ void __launch_bounds__(256)
adam_params_kernel(const float* __restrict__ gx6, const float* __restrict__ gbetas, const float* __restrict__ dx6c,
                   const float* __restrict__ dbeta_s, const float* __restrict__ ext_dx6,
                   const float* __restrict__ ext_dbetas, int64_t B, float* __restrict__ x6, float* __restrict__ betas,
                   float* __restrict__ adam_m, float* __restrict__ adam_v, const float* __restrict__ coef,
                   int32_t* __restrict__ step_count) {
  // thread = two consecutive parameters of one pose (NPARAM and 144 are even: a pair never straddles rot6d / betas or
  // two poses, and every array is 8-byte aligned at it): half the threads, twice the bytes in flight per thread
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the step counter advances here (nothing in this kernel reads it: the bias corrections of this step are in `coef`), not
  // in a one-thread kernel of its own at the very end of the step
  if (idx == 0) step_count[0] += 1;
  constexpr int HP = NPARAM / 2;
  if (idx >= B * HP) return;
  const int64_t b = idx / HP;
  const int p = 2 * (int)(idx - b * HP);
  float2 g;
  float2* prm;
  if (p < 144) {
    g = *reinterpret_cast<const float2*>(gx6 + b * 144 + p);
    if (dx6c != nullptr) {
      const float2 c = *reinterpret_cast<const float2*>(dx6c + b * 144 + p);
      g.x += c.x; g.y += c.y;
    }
    if (ext_dx6 != nullptr) {
      const float2 c = *reinterpret_cast<const float2*>(ext_dx6 + b * 144 + p);
      g.x += c.x; g.y += c.y;
    }
    prm = reinterpret_cast<float2*>(x6 + b * 144 + p);
  } else {
    g = *reinterpret_cast<const float2*>(gbetas + b * NB + (p - 144));
    if (dbeta_s != nullptr) {
      const float2 c = *reinterpret_cast<const float2*>(dbeta_s + b * NB + (p - 144));
      g.x += c.x; g.y += c.y;
    }
    if (ext_dbetas != nullptr) {
      const float2 c = *reinterpret_cast<const float2*>(ext_dbetas + b * NB + (p - 144));
      g.x += c.x; g.y += c.y;
    }
    prm = reinterpret_cast<float2*>(betas + b * NB + (p - 144));
  }
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc2s = coef[0], step = coef[1];      // adam_coef_kernel
  float2* pm = reinterpret_cast<float2*>(adam_m + b * NPARAM + p);
  float2* pv = reinterpret_cast<float2*>(adam_v + b * NPARAM + p);
  float2 m = *pm, v = *pv, x = *prm;
  m.x = b1 * m.x + (1.f - b1) * g.x;
  m.y = b1 * m.y + (1.f - b1) * g.y;
  v.x = b2 * v.x + (1.f - b2) * g.x * g.x;
  v.y = b2 * v.y + (1.f - b2) * g.y * g.y;
  *pm = m;
  *pv = v;
  x.x = x.x - step * (m.x / (sqrtf(v.x) / bc2s + eps));
  x.y = x.y - step * (m.y / (sqrtf(v.y) / bc2s + eps));
  *prm = x;
}

// Bias corrections of this step in double, like torch.optim.Adam's Python scalars (a double-precision pow is a
// few microseconds of dependent latency: one thread does it early in the step, off the critical path).
__global__ void adam_coef_kernel(const int32_t* __restrict__ step_count, float lr, float* __restrict__ coef) {
  const int t = *step_count + 1;
  coef[0] = (float)sqrt(1.0 - pow(0.999, (double)t));
  coef[1] = (float)((double)lr / (1.0 - pow(0.9, (double)t)));
}

int launch_adam_coef(const Workspace& w, const int32_t* step_count, float lr, cudaStream_t st) {
  adam_coef_kernel<<<1, 1, 0, st>>>(step_count, lr, w.adam_coef);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// ---- small batches: the whole module forward in ONE launch -------------------------------------
// Up to SMALL_MAX (<= 32: lane = pose when skinning) poses the tensor-core path pads the batch to a 128-row tile and a single epilogue lane skins every
// vertex of the pose (26 us at one pose) behind three more launches.  Here every block repeats the kinematic chain of the
// batch (warp per pose, cheap) into shared memory, then WARP = PACKED VERTEX: the lanes stream the vertex's three rows of
// the augmented blend matrix (K contiguous, hi + lo = the fp32 weights to 2^-22) with 16-byte loads, all twelve of them in
// flight at once, contract them with the features of four poses at a time, butterfly-reduce, and lanes 0..3 skin the
// vertex for their pose and store it in the model's vertex order.  The blend matrix is read once per block whatever the
// batch.  The block that finishes last (device counter, self-resetting) gathers the 49 joints from the stored vertices.
// Replaces smplx.lbs.lbs + vertex_joint_selector + scripts/smpl.py:75-78 for B <= SMALL_MAX.
constexpr int SMALL_MAX = 8;        // beyond this the tensor-core path is faster (measured: 44 us at 4 poses, 67 us at 64)
constexpr int SMALL_WARPS = 8;
constexpr int SMALL_CHUNK = 4;      // poses contracted per pass over the register-held weights

template <int KIND>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 2)
smpl_small_fwd_kernel(const __grid_constant__ ChainTab tab, const float* __restrict__ J0, const float* __restrict__ JS,
                      const float* __restrict__ betas, const float* __restrict__ pose, int B, int BS /* B rounded up to 4 */,
                      const float* __restrict__ Pt_hi, const float* __restrict__ Pt_lo, const VtxRec* __restrict__ vrec,
                      const int* __restrict__ perm, const int* __restrict__ joint_map, const int* __restrict__ picks,
                      Csr extra, float* __restrict__ vertices, float* __restrict__ joints49, unsigned* __restrict__ counter) {
  extern __shared__ float small_smem[];
  float* sA = small_smem;                    // [288][BS]
  float* sF = sA + NJ * 12 * BS;             // [BS][224]
  float* sJp = sF + BS * KA;                 // [BS][72]
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const bool tail = lane < (KA - 128) / 4;     // lanes that own a second 16-byte piece of a 224-float row
  const int gw = blockIdx.x * SMALL_WARPS + warp, GW = gridDim.x * SMALL_WARPS;
  // the three blend-matrix rows of packed vertex i, this lane's K slice (hi + lo): issued before they are needed
  float p[3][8];
  auto load_rows = [&](int i) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int64_t row = (int64_t)(3 * i + c) * KA;
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(Pt_hi + row) + lane);
      const float4 l0 = __ldg(reinterpret_cast<const float4*>(Pt_lo + row) + lane);
      float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f), l1 = h1;
      if (tail) {
        h1 = __ldg(reinterpret_cast<const float4*>(Pt_hi + row + 128) + lane);
        l1 = __ldg(reinterpret_cast<const float4*>(Pt_lo + row + 128) + lane);
      }
      p[c][0] = h0.x + l0.x; p[c][1] = h0.y + l0.y; p[c][2] = h0.z + l0.z; p[c][3] = h0.w + l0.w;
      p[c][4] = h1.x + l1.x; p[c][5] = h1.y + l1.y; p[c][6] = h1.z + l1.z; p[c][7] = h1.w + l1.w;
    }
  };
  if (gw < VP) load_rows(gw);                  // in flight while the chain runs

  // ---- kinematic chain of every pose of the batch (warp per pose, lane = joint)
  for (int b = warp; b < BS; b += SMALL_WARPS) {
    const bool valid = b < B;
    const int j = lane < NJ ? lane : NJ - 1;
    float beta[NB];
    for (int l = 0; l < NB; l++) beta[l] = valid ? betas[b * NB + l] : 0.f;
    float raw[9], R[9], Jr[3], GR[9], Gt[3], GRp[9], rel[3];
    decode_rot<KIND>(pose, b, j, valid, raw, R);
    rest_joint(J0, JS, beta, j, Jr);
    chain_forward(tab, j, R, Jr, GR, Gt, GRp, rel);
    for (int i = lane; i < KA; i += 32) sF[b * KA + i] = 0.f;
    __syncwarp();
    if (lane < NJ) {
      for (int r = 0; r < 3; r++) {
        const float t = Gt[r] - (GR[r * 3 + 0] * Jr[0] + GR[r * 3 + 1] * Jr[1] + GR[r * 3 + 2] * Jr[2]);
        sA[(lane * 12 + r * 4 + 0) * BS + b] = GR[r * 3 + 0];
        sA[(lane * 12 + r * 4 + 1) * BS + b] = GR[r * 3 + 1];
        sA[(lane * 12 + r * 4 + 2) * BS + b] = GR[r * 3 + 2];
        sA[(lane * 12 + r * 4 + 3) * BS + b] = t;
        sJp[b * 72 + lane * 3 + r] = Gt[r];
      }
      if (lane >= 1)
        for (int i = 0; i < 9; i++) sF[b * KA + (lane - 1) * 9 + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
    }
    if (lane < NB) sF[b * KA + FEAT_BETA + lane] = beta[lane];
    if (lane == 0) sF[b * KA + FEAT_ONE] = 1.f;
  }
  __syncthreads();

  // ---- warp = packed vertex
  for (int i = gw; i < VP; i += GW) {
    if (i != gw) load_rows(i);
    const int vid = perm[i];
    if (vid < 0) continue;                     // padding (warp-uniform)
    const VtxRec* rec = vrec + i;
    const uint32_t meta = __ldg(&rec->meta);
    float wk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) wk[k] = __ldg(&rec->w[k]);

    // blended vertex of pose b lands in lane b
    float x = 0.f, y = 0.f, z = 0.f;
    for (int b0 = 0; b0 < BS; b0 += SMALL_CHUNK) {
      float acc[SMALL_CHUNK][3];
#pragma unroll
      for (int bb = 0; bb < SMALL_CHUNK; bb++) {
        const float4 f0 = *reinterpret_cast<const float4*>(sF + (b0 + bb) * KA + 4 * lane);
        float4 f1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tail) f1 = *reinterpret_cast<const float4*>(sF + (b0 + bb) * KA + 128 + 4 * lane);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float a = p[c][0] * f0.x;
          a = fmaf(p[c][1], f0.y, a); a = fmaf(p[c][2], f0.z, a); a = fmaf(p[c][3], f0.w, a);
          a = fmaf(p[c][4], f1.x, a); a = fmaf(p[c][5], f1.y, a); a = fmaf(p[c][6], f1.z, a); a = fmaf(p[c][7], f1.w, a);
          acc[bb][c] = a;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int bb = 0; bb < SMALL_CHUNK; bb++)
#pragma unroll
          for (int c = 0; c < 3; c++) acc[bb][c] += __shfl_xor_sync(FULL, acc[bb][c], o);
#pragma unroll
      for (int bb = 0; bb < SMALL_CHUNK; bb++)
        if (lane == b0 + bb) { x = acc[bb][0]; y = acc[bb][1]; z = acc[bb][2]; }
    }
    if (lane < B) {
      const int b = lane;
      float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float* a = sA + (int)((meta >> (5 * k)) & 31u) * 12 * BS + b;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const float yk = fmaf(a[(r * 4 + 2) * BS], z, fmaf(a[(r * 4 + 1) * BS], y, fmaf(a[(r * 4 + 0) * BS], x, a[(r * 4 + 3) * BS])));
          v[r] = fmaf(wk[k], yk, v[r]);
        }
      }
      float* dst = vertices + ((int64_t)b * V + vid) * 3;
      dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
    }
  }
  if (joints49 == nullptr) return;

  // ---- the block that finishes last gathers the 49 joints (24 posed joints, 21 vertex picks, 9 extra-regressor rows)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // posed joints and vertex picks: one thread per (pose, joint), every dependent load chain in flight at once
  for (int t = threadIdx.x; t < B * JRR_NUM_OUT_JOINTS; t += SMALL_WARPS * 32) {
    const int b = t / JRR_NUM_OUT_JOINTS, src = joint_map[t % JRR_NUM_OUT_JOINTS];
    if (src >= NJ + JRR_NUM_PICKS) continue;
    float* dst = joints49 + (int64_t)t * 3;
    if (src < NJ) {
      for (int c = 0; c < 3; c++) dst[c] = sJp[b * 72 + src * 3 + c];
    } else {
      const float* vsrc = vertices + ((int64_t)b * V + picks[src - NJ]) * 3;
      for (int c = 0; c < 3; c++) dst[c] = __ldcg(vsrc + c);
    }
  }
  // extra-regressor rows: one warp per (pose, row), lanes stride the row's non-zeros, fixed-order butterfly
  for (int task = warp; task < B * JRR_NUM_OUT_JOINTS; task += SMALL_WARPS) {
    const int b = task / JRR_NUM_OUT_JOINTS, src = joint_map[task % JRR_NUM_OUT_JOINTS];
    if (src < NJ + JRR_NUM_PICKS) continue;
    const int e = src - NJ - JRR_NUM_PICKS;
    const float* vb = vertices + (int64_t)b * V * 3;
    float out[3] = {0.f, 0.f, 0.f};
    for (int q = extra.ptr[e] + lane; q < extra.ptr[e + 1]; q += 32) {
      const int v = extra.col[q];
      const float cf = extra.val[q];
      for (int c = 0; c < 3; c++) out[c] = fmaf(cf, __ldcg(vb + v * 3 + c), out[c]);
    }
    for (int sft = 16; sft > 0; sft >>= 1)
      for (int c = 0; c < 3; c++) out[c] += __shfl_xor_sync(FULL, out[c], sft);
    if (lane < 3) joints49[(int64_t)task * 3 + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
  }
  if (threadIdx.x == 0) *counter = 0u;
}

bool smpl_small_fwd_available(const JrrModel* m, int64_t B) {
  static const bool on = [] { const char* e = getenv("JRR_SMALL_FWD"); return !(e && e[0] == '0'); }();
  return on && B <= SMALL_MAX && m->n_pass == 1 && m->gemm_impl == 0 && m->small_counter != nullptr;
}

int launch_smpl_small_fwd(const JrrModel* m, int64_t B, const float* betas, const float* pose, int kind,
                          float* vertices, float* joints49, cudaStream_t st) {
  const int BS = (int)round_up(B, SMALL_CHUNK);
  const size_t smem = (size_t)(NJ * 12 * BS + BS * KA + BS * 72) * sizeof(float);
  const int grid = 2 * m->num_sms;
#define JRR_SF(KIND)                                                                                               \
  do {                                                                                                             \
    auto kern = smpl_small_fwd_kernel<KIND>;                                                                       \
    JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
    kern<<<grid, SMALL_WARPS * 32, smem, st>>>(m->chain, m->J0, m->JS, betas, pose, (int)B, BS, m->Pt_hi, m->Pt_lo, \
                                               m->passes[0].vrec, m->perm, m->joint_map, m->picks, m->extra, vertices, \
                                               joints49, m->small_counter);                                       \
  } while (0)
  switch (kind) {
    case JRR_POSE_ROTMAT: JRR_SF(JRR_POSE_ROTMAT); break;
    case JRR_POSE_AXIS_ANGLE: JRR_SF(JRR_POSE_AXIS_ANGLE); break;
    case JRR_POSE_ROT6D: JRR_SF(JRR_POSE_ROT6D); break;
    default: return fail(JRR_ERR_INVALID, "unknown pose kind");
  }
#undef JRR_SF
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// ---- small batches: the module BACKWARD in three launches ----------------------------------------------------------
// Same grain as smpl_small_fwd_kernel (warp = packed vertex, lanes = the vertex's blend-matrix rows split along K, lane b =
// pose b once the blended vertex is reduced): every block repeats the chain, recomputes the blended vertex, takes the
// vertex gradient (the caller's d loss / d vertices plus the joints49 gradient that reaches this vertex through a pick or an
// extra-regressor row), and leaves (i) its contribution to the joint-transform gradients in a per-WARP shared accumulator
// (lanes = poses: no two lanes touch one address, no atomics), (ii) its contribution to the blend-feature gradient in the
// lanes' registers (d blended vertex broadcast by shuffles against the rows the lane already holds).  Warps, then blocks,
// are summed in a fixed order (small_bwd_reduce_kernel), and the unchanged chain backward (pose_bwd_kernel) finishes.
// The tensor-core path pads to 256 poses and walks the vertices in 36 serial ranges: 180 us at one pose.
constexpr int SMALLB_DA = NJ * 12;        // 288 joint-transform entries

template <int KIND>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 2)
smpl_small_bwd_kernel(const __grid_constant__ ChainTab tab, const float* __restrict__ J0, const float* __restrict__ JS,
                      const float* __restrict__ betas, const float* __restrict__ pose, int B, int BS,
                      const float* __restrict__ Pt_hi, const float* __restrict__ Pt_lo, const VtxRec* __restrict__ vrec,
                      const int* __restrict__ perm, const int* __restrict__ joint_map, const int* __restrict__ vx_src,
                      const float* __restrict__ vx_coef, const float* __restrict__ dverts, const float* __restrict__ dj49,
                      float* __restrict__ part_dA /* [grid][288][BS] */, float* __restrict__ part_df /* [grid][BS][224] */) {
  extern __shared__ float small_smem[];
  float* sA = small_smem;                               // [288][BS]
  float* sF = sA + SMALLB_DA * BS;                      // [BS][224]
  float* sd30 = sF + BS * KA;                           // [BS][90]  joints49 gradient on its 30 vertex-borne sources
  float* wdA = sd30 + BS * 90;                          // [SMALL_WARPS][288][BS]; later [SMALL_WARPS][BS][224] (it is larger)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tail = lane < (KA - 128) / 4;
  const int gw = blockIdx.x * SMALL_WARPS + warp, GW = gridDim.x * SMALL_WARPS;
  float p[3][8];
  auto load_rows = [&](int i) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int64_t row = (int64_t)(3 * i + c) * KA;
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(Pt_hi + row) + lane);
      const float4 l0 = __ldg(reinterpret_cast<const float4*>(Pt_lo + row) + lane);
      float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f), l1 = h1;
      if (tail) {
        h1 = __ldg(reinterpret_cast<const float4*>(Pt_hi + row + 128) + lane);
        l1 = __ldg(reinterpret_cast<const float4*>(Pt_lo + row + 128) + lane);
      }
      p[c][0] = h0.x + l0.x; p[c][1] = h0.y + l0.y; p[c][2] = h0.z + l0.z; p[c][3] = h0.w + l0.w;
      p[c][4] = h1.x + l1.x; p[c][5] = h1.y + l1.y; p[c][6] = h1.z + l1.z; p[c][7] = h1.w + l1.w;
    }
  };
  if (gw < VP) load_rows(gw);
  for (int e = threadIdx.x; e < SMALL_WARPS * SMALLB_DA * BS; e += SMALL_WARPS * 32) wdA[e] = 0.f;
  for (int e = threadIdx.x; e < BS * 90; e += SMALL_WARPS * 32) {
    const int b = e / 90, sc = e % 90, src = NJ + sc / 3, c = sc % 3;
    float d = 0.f;
    if (b < B && dj49 != nullptr)
      for (int o = 0; o < JRR_NUM_OUT_JOINTS; o++)
        if (joint_map[o] == src) d += dj49[((int64_t)b * JRR_NUM_OUT_JOINTS + o) * 3 + c];
    sd30[e] = d;
  }
  for (int b = warp; b < BS; b += SMALL_WARPS) {
    const bool valid = b < B;
    const int j = lane < NJ ? lane : NJ - 1;
    float beta[NB];
    for (int l = 0; l < NB; l++) beta[l] = valid ? betas[b * NB + l] : 0.f;
    float raw[9], R[9], Jr[3], GR[9], Gt[3], GRp[9], rel[3];
    decode_rot<KIND>(pose, b, j, valid, raw, R);
    rest_joint(J0, JS, beta, j, Jr);
    chain_forward(tab, j, R, Jr, GR, Gt, GRp, rel);
    for (int i = lane; i < KA; i += 32) sF[b * KA + i] = 0.f;
    __syncwarp();
    if (lane < NJ) {
      for (int r = 0; r < 3; r++) {
        const float t = Gt[r] - (GR[r * 3 + 0] * Jr[0] + GR[r * 3 + 1] * Jr[1] + GR[r * 3 + 2] * Jr[2]);
        sA[(lane * 12 + r * 4 + 0) * BS + b] = GR[r * 3 + 0];
        sA[(lane * 12 + r * 4 + 1) * BS + b] = GR[r * 3 + 1];
        sA[(lane * 12 + r * 4 + 2) * BS + b] = GR[r * 3 + 2];
        sA[(lane * 12 + r * 4 + 3) * BS + b] = t;
      }
      if (lane >= 1)
        for (int i = 0; i < 9; i++) sF[b * KA + (lane - 1) * 9 + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
    }
    if (lane < NB) sF[b * KA + FEAT_BETA + lane] = beta[lane];
    if (lane == 0) sF[b * KA + FEAT_ONE] = 1.f;
  }
  __syncthreads();

  float dfp[8][8];                 // [pose][this lane's 8 blend features]: d loss / d feature, summed over the warp's vertices
#pragma unroll
  for (int b = 0; b < 8; b++)
#pragma unroll
    for (int e = 0; e < 8; e++) dfp[b][e] = 0.f;
  float* mydA = wdA + warp * SMALLB_DA * BS;
  for (int i = gw; i < VP; i += GW) {
    if (i != gw) load_rows(i);
    const int vid = perm[i];
    if (vid < 0) continue;
    const VtxRec* rec = vrec + i;
    const uint32_t meta = __ldg(&rec->meta);
    float wk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) wk[k] = __ldg(&rec->w[k]);
    const int xptr = __ldg(&rec->xptr), xcnt = __ldg(&rec->xcnt);
    float x = 0.f, y = 0.f, z = 0.f;
    for (int b0 = 0; b0 < BS; b0 += SMALL_CHUNK) {
      float acc[SMALL_CHUNK][3];
#pragma unroll
      for (int bb = 0; bb < SMALL_CHUNK; bb++) {
        const float4 f0 = *reinterpret_cast<const float4*>(sF + (b0 + bb) * KA + 4 * lane);
        float4 f1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tail) f1 = *reinterpret_cast<const float4*>(sF + (b0 + bb) * KA + 128 + 4 * lane);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float a = p[c][0] * f0.x;
          a = fmaf(p[c][1], f0.y, a); a = fmaf(p[c][2], f0.z, a); a = fmaf(p[c][3], f0.w, a);
          a = fmaf(p[c][4], f1.x, a); a = fmaf(p[c][5], f1.y, a); a = fmaf(p[c][6], f1.z, a); a = fmaf(p[c][7], f1.w, a);
          acc[bb][c] = a;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int bb = 0; bb < SMALL_CHUNK; bb++)
#pragma unroll
          for (int c = 0; c < 3; c++) acc[bb][c] += __shfl_xor_sync(FULL, acc[bb][c], o);
#pragma unroll
      for (int bb = 0; bb < SMALL_CHUNK; bb++)
        if (lane == b0 + bb) { x = acc[bb][0]; y = acc[bb][1]; z = acc[bb][2]; }
    }
    // lane b: the vertex gradient of pose b, its share of dA, and d blended vertex
    float dvp0 = 0.f, dvp1 = 0.f, dvp2 = 0.f;
    if (lane < B) {
      const int b = lane;
      float dv[3] = {0.f, 0.f, 0.f};
      if (dverts != nullptr) {
        const float* src = dverts + ((int64_t)b * V + vid) * 3;
        dv[0] = src[0]; dv[1] = src[1]; dv[2] = src[2];
      }
      for (int q = 0; q < xcnt; q++) {
        const float cf = vx_coef[xptr + q];
        const float* d3 = sd30 + b * 90 + vx_src[xptr + q] * 3;
        dv[0] = fmaf(cf, d3[0], dv[0]); dv[1] = fmaf(cf, d3[1], dv[1]); dv[2] = fmaf(cf, d3[2], dv[2]);
      }
      const float v4[4] = {x, y, z, 1.f};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int j = (int)((meta >> (5 * k)) & 31u);
        const float* a = sA + j * 12 * BS + b;
        float* da = mydA + j * 12 * BS + b;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const float u = wk[k] * dv[r];
          dvp0 = fmaf(a[(r * 4 + 0) * BS], u, dvp0);
          dvp1 = fmaf(a[(r * 4 + 1) * BS], u, dvp1);
          dvp2 = fmaf(a[(r * 4 + 2) * BS], u, dvp2);
#pragma unroll
          for (int cc = 0; cc < 4; cc++) da[(r * 4 + cc) * BS] += u * v4[cc];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (b < B) {                                   // (uniform)
        const float d0 = __shfl_sync(FULL, dvp0, b), d1 = __shfl_sync(FULL, dvp1, b), d2 = __shfl_sync(FULL, dvp2, b);
#pragma unroll
        for (int e = 0; e < 8; e++) dfp[b][e] = fmaf(d0, p[0][e], fmaf(d1, p[1][e], fmaf(d2, p[2][e], dfp[b][e])));
      }
    }
  }
  __syncthreads();
  // ---- joint-transform gradients: warps summed in order -> this block's partial
  for (int e = threadIdx.x; e < SMALLB_DA * BS; e += SMALL_WARPS * 32) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < SMALL_WARPS; w8++) a += wdA[w8 * SMALLB_DA * BS + e];
    part_dA[(int64_t)blockIdx.x * SMALLB_DA * BS + e] = a;
  }
  __syncthreads();
  // ---- blend-feature gradients: lanes' registers -> [warp][pose][224] (re-using the accumulator region), warps summed in order
  float* wdf = wdA;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    if (b < BS) {
      float* dst = wdf + (warp * BS + b) * KA;
      *reinterpret_cast<float4*>(dst + 4 * lane) = make_float4(dfp[b][0], dfp[b][1], dfp[b][2], dfp[b][3]);
      if (tail) *reinterpret_cast<float4*>(dst + 128 + 4 * lane) = make_float4(dfp[b][4], dfp[b][5], dfp[b][6], dfp[b][7]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < BS * KA; e += SMALL_WARPS * 32) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < SMALL_WARPS; w8++) a += wdf[w8 * BS * KA + e];
    part_df[(int64_t)blockIdx.x * BS * KA + e] = a;
  }
}

// fixed-order sum of `n` block partials `stride` floats apart: four interleaved chains (eight loads in flight per thread; one
// chain of dependent adds behind one load at a time made this kernel as long as the gradient kernel itself, 23 us)
__device__ __forceinline__ float small_sum_partials(const float* __restrict__ p, int n, int64_t stride) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
#pragma unroll 2
  for (; k + 3 < n; k += 4) {
    a0 += p[(int64_t)k * stride];
    a1 += p[(int64_t)(k + 1) * stride];
    a2 += p[(int64_t)(k + 2) * stride];
    a3 += p[(int64_t)(k + 3) * stride];
  }
  for (; k < n; k++) a0 += p[(int64_t)k * stride];
  return (a0 + a1) + (a2 + a3);
}

// block partials summed in order -> dAT [288][BP], dfeat [1][BP][224] (one split), dJp [BP][72] for the chain backward
// (b0: the first pose of this group of up to 8 inside the batch; dj49 already points at the group)
__global__ void small_bwd_reduce_kernel(const float* __restrict__ part_dA, const float* __restrict__ part_df, int nblk, int B,
                                        int BS, int64_t BP, int b0, const int* __restrict__ joint_map,
                                        const float* __restrict__ dj49, float* __restrict__ dAT, float* __restrict__ dfeat,
                                        float* __restrict__ dJp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nA = SMALLB_DA * BS, nF = BS * KA, nJ = BS * 72;
  if (idx < nA) {
    const int e = idx / BS, b = idx % BS;
    const float a = small_sum_partials(part_dA + idx, nblk, nA);
    dAT[(int64_t)e * BP + b0 + b] = b < B ? a : 0.f;
  } else if (idx < nA + nF) {
    const int q = idx - nA, b = q / KA, k = q % KA;
    const float a = small_sum_partials(part_df + q, nblk, nF);
    dfeat[(int64_t)(b0 + b) * KA + k] = b < B ? a : 0.f;
  } else if (idx < nA + nF + nJ) {
    const int q = idx - nA - nF, b = q / 72, src = (q % 72) / 3, c = q % 3;
    float d = 0.f;
    if (b < B && dj49 != nullptr)
      for (int o = 0; o < JRR_NUM_OUT_JOINTS; o++)
        if (joint_map[o] == src) d += dj49[((int64_t)b * JRR_NUM_OUT_JOINTS + o) * 3 + c];
    dJp[(int64_t)(b0 + b) * 72 + src * 3 + c] = d;
  }
}

constexpr int SMALLB_MAX = 16;        // two groups of 8 poses still beat the padded tensor-core path (127 us at 8, 250 us at 16)

bool smpl_small_bwd_available(const JrrModel* m, int64_t B) {
  static const bool on = [] { const char* e = getenv("JRR_SMALL_BWD"); return !(e && e[0] == '0'); }();
  return on && B <= SMALLB_MAX && m->n_pass == 1 && m->gemm_impl == 0;
}

// leaves dAT / dfeat (ksplit = 1) / dJp in the workspace for launch_pose_bwd; scratch: the (idle) blend-gradient buffer.
// Groups of up to 8 poses (the lanes' register budget for the feature gradients), each with its own pair of launches.
int launch_smpl_small_bwd(const JrrModel* m, Workspace& w, const float* betas, const float* pose, int kind,
                          const float* dverts, const float* dj49, cudaStream_t st) {
  const int grid = 2 * m->num_sms;
  const int pose_stride = kind == JRR_POSE_ROTMAT ? NJ * 9 : (kind == JRR_POSE_AXIS_ANGLE ? NJ * 3 : NJ * 6);
  if (kind != JRR_POSE_ROTMAT && kind != JRR_POSE_AXIS_ANGLE && kind != JRR_POSE_ROT6D) return fail(JRR_ERR_INVALID, "unknown pose kind");
  for (int b0 = 0; b0 < (int)w.B; b0 += SMALL_MAX) {
    const int B = std::min((int)w.B - b0, SMALL_MAX), BS = (int)round_up(B, SMALL_CHUNK);
    const size_t smem = (size_t)(SMALLB_DA * BS + BS * KA + BS * 90 + SMALL_WARPS * SMALLB_DA * BS) * sizeof(float);
    float* part_dA = w.dvp_hi;
    float* part_df = part_dA + (size_t)grid * SMALLB_DA * BS;
    const float* g_betas = betas + (size_t)b0 * NB;
    const float* g_pose = pose + (size_t)b0 * pose_stride;
    const float* g_dv = dverts ? dverts + (size_t)b0 * V * 3 : nullptr;
    const float* g_dj = dj49 ? dj49 + (size_t)b0 * JRR_NUM_OUT_JOINTS * 3 : nullptr;
#define JRR_SB(KIND)                                                                                               \
    do {                                                                                                           \
      auto kern = smpl_small_bwd_kernel<KIND>;                                                                     \
      JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
      kern<<<grid, SMALL_WARPS * 32, smem, st>>>(m->chain, m->J0, m->JS, g_betas, g_pose, B, BS, m->Pt_hi, m->Pt_lo, \
                                                 m->passes[0].vrec, m->perm, m->joint_map, m->vx_src, m->vx_coef, g_dv, \
                                                 g_dj, part_dA, part_df);                                          \
    } while (0)
    switch (kind) {
      case JRR_POSE_ROTMAT: JRR_SB(JRR_POSE_ROTMAT); break;
      case JRR_POSE_AXIS_ANGLE: JRR_SB(JRR_POSE_AXIS_ANGLE); break;
      default: JRR_SB(JRR_POSE_ROT6D); break;
    }
#undef JRR_SB
    JRR_LAUNCH_CHECK();
    const int n = SMALLB_DA * BS + BS * KA + BS * 72;
    small_bwd_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part_dA, part_df, grid, B, BS, w.BP, b0, m->joint_map, g_dj, w.dAT,
                                                            w.dfeat, w.dJp);
    JRR_LAUNCH_CHECK();
  }
  w.ksplit = 1;
  return JRR_OK;
}

// ---- host wrappers ---------------------------------------------------------------------------
int launch_adam_params(const Workspace& w, bool use_critic, bool use_shape, float* x6, float* betas, float* adam_m,
                       float* adam_v, int32_t* step_count, float lr, cudaStream_t st, const float* ext_dx6,
                       const float* ext_dbetas) {
  const int64_t n = w.B * (NPARAM / 2);
  adam_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.gx6, w.gbetas, use_critic ? w.dx6c : nullptr,
                                                                 use_shape ? w.dbeta_s : nullptr, ext_dx6, ext_dbetas, w.B, x6, betas, adam_m,
                                                                 adam_v, w.adam_coef, step_count);
  (void)lr;
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_pose_fwd(const JrrModel* m, int64_t B, int64_t BP, const float* betas, const float* pose,
                    int kind, float* AT, float* feat_hi, float* feat_lo, float* Jp, cudaStream_t st) {
  dim3 grid((unsigned)(BP / POSES_PER_CTA)), block(POSES_PER_CTA * 32);
  switch (kind) {
    case JRR_POSE_ROTMAT:
      pose_fwd_kernel<JRR_POSE_ROTMAT><<<grid, block, 0, st>>>(m->chain, m->J0, m->JS, betas, pose, B, BP, AT, feat_hi, feat_lo, Jp);
      break;
    case JRR_POSE_AXIS_ANGLE:
      pose_fwd_kernel<JRR_POSE_AXIS_ANGLE><<<grid, block, 0, st>>>(m->chain, m->J0, m->JS, betas, pose, B, BP, AT, feat_hi, feat_lo, Jp);
      break;
    case JRR_POSE_ROT6D:
      pose_fwd_kernel<JRR_POSE_ROT6D><<<grid, block, 0, st>>>(m->chain, m->J0, m->JS, betas, pose, B, BP, AT, feat_hi, feat_lo, Jp);
      break;
    default:
      return fail(JRR_ERR_INVALID, "unknown pose kind");
  }
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_pose_bwd(const JrrModel* m, const Workspace& w, const float* betas, const float* pose,
                    int kind, bool use_dJp, bool use_critic, bool use_shape, float* dbetas_out, float* dpose_out,
                    float* x6, float* betas_rw, float* adam_m, float* adam_v, int32_t* step_count,
                    float lr, cudaStream_t st) {
  dim3 grid((unsigned)((w.B + POSES_PER_CTA - 1) / POSES_PER_CTA)), block(POSES_PER_CTA * 32);
  const float* dJp = use_dJp ? w.dJp : nullptr;
  const float* dx6c = use_critic ? w.dx6c : nullptr;
  const float* dbs = use_shape ? w.dbeta_s : nullptr;
  const bool adam = x6 != nullptr;
#define JRR_PB(KIND, AD)                                                                          \
  pose_bwd_kernel<KIND, AD><<<grid, block, 0, st>>>(m->chain, m->J0, m->JS, betas, pose, w.B, w.BP, \
      w.dAT, w.dfeat, w.ksplit, dJp, dx6c, dbs, dbetas_out, dpose_out, x6, betas_rw, adam_m, adam_v, \
      step_count, lr)
  if (adam) {
    if (kind != JRR_POSE_ROT6D) return fail(JRR_ERR_INVALID, "Adam step needs rot6d parameters");
    JRR_PB(JRR_POSE_ROT6D, true);
    JRR_LAUNCH_CHECK();
    bump_step_kernel<<<1, 1, 0, st>>>(step_count);
    JRR_LAUNCH_CHECK();
  } else {
    switch (kind) {
      case JRR_POSE_ROTMAT: JRR_PB(JRR_POSE_ROTMAT, false); break;
      case JRR_POSE_AXIS_ANGLE: JRR_PB(JRR_POSE_AXIS_ANGLE, false); break;
      case JRR_POSE_ROT6D: JRR_PB(JRR_POSE_ROT6D, false); break;
      default: return fail(JRR_ERR_INVALID, "unknown pose kind");
    }
    JRR_LAUNCH_CHECK();
  }
#undef JRR_PB
  return JRR_OK;
}

}  // namespace jrr
