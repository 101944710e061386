// Pose critic (scripts/discriminator.py:7-54) on the refinement path: the per-joint 1x1
// conv stack, the 24 joint heads, the sigmoid + MSE-vs-ones loss (scripts/optimize.py:241-247)
// and the analytic INPUT gradient (weights are frozen inside the inner loop).  The two wide
// layers 768->1024->1024 (and their transposes in the backward) are 3xTF32 tensor-core GEMMs
// (jrr_gemm_tc.cu); this file holds the small fused pieces around them.
#include "jrr_internal.cuh"

namespace jrr {


__device__ __forceinline__ float tf32_hi_c(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// thread = (pose, joint): conv1x1 6->32, ReLU, conv1x1 32->32, ReLU -> h[b][j*32+c] (hi/lo)
// (<= 85 registers: three CTAs per SM, so the 384 CTAs of a 4096-frame step are one wave, not 1.3)
__global__ void __launch_bounds__(256, 3)
critic_pre_kernel(const float* __restrict__ cs, const float* __restrict__ x6, int64_t B, int64_t BP,
                  float* __restrict__ h_hi, float* __restrict__ h_lo, float* __restrict__ zj_out,
                  uint2* __restrict__ masks_out) {
  // conv weights TRANSPOSED in shared memory ([input][output]): the inner loops then run over 32 independent
  // accumulators (one broadcast LDS.128 feeds four of them) instead of one 32-deep dependent chain per output;
  // every output still sums bias, k = 0, 1, ... in the same order
  __shared__ __align__(16) float sw1t[6 * 32];
  __shared__ __align__(16) float sw2t[32 * 32];
  __shared__ float sb[64];
  __shared__ float shw[NJ * 33];     // joint-head weights, row stride 33: lanes hold different joints
  for (int i = threadIdx.x; i < 192; i += blockDim.x) sw1t[(i % 6) * 32 + i / 6] = cs[CS_C1W + i];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sw2t[(i & 31) * 32 + (i >> 5)] = cs[CS_C2W + i];
  if (threadIdx.x < 32) { sb[threadIdx.x] = cs[CS_C1B + threadIdx.x]; sb[32 + threadIdx.x] = cs[CS_C2B + threadIdx.x]; }
  if (zj_out != nullptr)
    for (int i = threadIdx.x; i < NJ * 32; i += blockDim.x) shw[(i >> 5) * 33 + (i & 31)] = cs[CS_HW + i];
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BP * NJ) return;
  const int64_t b = idx / NJ;
  float h2[32];
  uint32_t m1 = 0, m2 = 0;     // ReLU masks of the two convs, reused by critic_post_kernel
  if (b < B) {
    float x[6];
    for (int i = 0; i < 6; i++) x[i] = x6[idx * 6 + i];
    float h1[32];
#pragma unroll
    for (int k = 0; k < 32; k++) h1[k] = sb[k];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int k = 0; k < 32; k++) h1[k] = fmaf(sw1t[i * 32 + k], x[i], h1[k]);
#pragma unroll
    for (int k = 0; k < 32; k++) {
      m1 |= (h1[k] > 0.f ? 1u : 0u) << k;
      h1[k] = fmaxf(h1[k], 0.f);
    }
#pragma unroll
    for (int c = 0; c < 32; c++) h2[c] = sb[32 + c];
#pragma unroll
    for (int k = 0; k < 32; k++)
#pragma unroll
      for (int c = 0; c < 32; c++) h2[c] = fmaf(sw2t[k * 32 + c], h1[k], h2[c]);
#pragma unroll
    for (int c = 0; c < 32; c++) {
      m2 |= (h2[c] > 0.f ? 1u : 0u) << c;
      h2[c] = fmaxf(h2[c], 0.f);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; c++) h2[c] = 0.f;
  }
  if (masks_out != nullptr) masks_out[idx] = make_uint2(m1, m2);
  if (zj_out != nullptr) {   // joint head logit (linears.j), consumed by critic_head_light_kernel
    const int j = (int)(idx % NJ);
    float zj = __ldg(cs + CS_HB + j);
#pragma unroll
    for (int c = 0; c < 32; c++) zj = fmaf(shw[j * 33 + c], h2[c], zj);
    zj_out[idx] = zj;
  }
  float4* ph = reinterpret_cast<float4*>(h_hi + idx * 32);
  if (h_lo == nullptr) {   // plain fp32 features: the GEMM splits them in shared memory
#pragma unroll
    for (int q = 0; q < 8; q++) ph[q] = make_float4(h2[q * 4], h2[q * 4 + 1], h2[q * 4 + 2], h2[q * 4 + 3]);
    return;
  }
  float4* pl = reinterpret_cast<float4*>(h_lo + idx * 32);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    float hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      hi[t] = tf32_hi_c(h2[q * 4 + t]);
      lo[t] = tf32_hi_c(h2[q * 4 + t] - hi[t]);
    }
    ph[q] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    pl[q] = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// warp = pose: global head 1024->1, 24 joint heads 32->1, sigmoid, loss partial, logits grads
constexpr int HEAD_WARPS = 8;
__global__ void __launch_bounds__(HEAD_WARPS * 32)
critic_head_kernel(const float* __restrict__ cs, const float* __restrict__ h_hi,
                   const float* __restrict__ h_lo, const float* __restrict__ z2_hi,
                   const float* __restrict__ z2_lo, int64_t B, int64_t BP, float gscale, float target,
                   float* __restrict__ scores_out, float* __restrict__ dz2_hi,
                   float* __restrict__ dz2_lo, float* __restrict__ dzj, float* __restrict__ dzg,
                   float* __restrict__ loss_part) {
  __shared__ float red[HEAD_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * HEAD_WARPS + warp;
  float lsum = 0.f;
  if (b < BP) {
    float z2[32];
    float zg = 0.f;
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const int i = lane + 32 * q;
      z2[q] = z2_hi[b * C_Z + i] + z2_lo[b * C_Z + i];
      zg = fmaf(cs[CS_W3 + i], z2[q], zg);
    }
    for (int o = 16; o > 0; o >>= 1) zg += __shfl_xor_sync(0xffffffffu, zg, o);
    zg += cs[CS_B3];
    float zj = 0.f;
    if (lane < NJ) {
      const float* hh = h_hi + b * C_H + lane * 32;
      const float* hl = h_lo + b * C_H + lane * 32;
      zj = cs[CS_HB + lane];
#pragma unroll
      for (int c = 0; c < 32; c++) zj = fmaf(cs[CS_HW + lane * 32 + c], hh[c] + hl[c], zj);
    }
    const float sg = 1.f / (1.f + expf(-zg));
    const float sj = 1.f / (1.f + expf(-zj));
    if (b < B) {
      if (scores_out != nullptr) {
        if (lane == 0) scores_out[b * 25] = sg;
        if (lane < NJ) scores_out[b * 25 + 1 + lane] = sj;
      }
      float l = (lane < NJ) ? (sj - target) * (sj - target) : 0.f;
      if (lane == 0) l += (sg - target) * (sg - target);
      for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
      lsum = l;
    }
    if (dz2_hi != nullptr) {
      const float dg = (b < B) ? gscale * (sg - target) * sg * (1.f - sg) : 0.f;
      const float dj = (b < B && lane < NJ) ? gscale * (sj - target) * sj * (1.f - sj) : 0.f;
      if (lane < NJ) dzj[b * NJ + lane] = dj;
      if (dzg != nullptr && lane == 0) dzg[b] = dg;
#pragma unroll
      for (int q = 0; q < 32; q++) {
        const int i = lane + 32 * q;
        const float v = z2[q] > 0.f ? dg * cs[CS_W3 + i] : 0.f;
        const float hi = tf32_hi_c(v);
        dz2_hi[b * C_Z + i] = hi;
        dz2_lo[b * C_Z + i] = tf32_hi_c(v - hi);
      }
    }
  }
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss_part != nullptr) {
    float t = 0.f;
    for (int i = 0; i < HEAD_WARPS; i++) t += red[i];
    loss_part[blockIdx.x] = t;
  }
}

// Refinement-step head when the global head's logit partials come out of the second wide GEMM's
// epilogue (EPI_BIAS_RELU_HEAD) and the joint logits out of critic_pre: warp = pose, lane < 24 the
// joint heads, lane 24 the global head.  Writes dL/dlogit (dzj, dzg) and the loss partial.
__global__ void __launch_bounds__(HEAD_WARPS * 32)
critic_head_light_kernel(const float* __restrict__ cs, const float* __restrict__ zj, const float* __restrict__ zg_part,
                         int n_part, int64_t B, int64_t BP, float gscale, float* __restrict__ dzj,
                         float* __restrict__ dzg, float* __restrict__ loss_part) {
  __shared__ float red[HEAD_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * HEAD_WARPS + warp;
  float l = 0.f;
  if (b < BP) {
    float z = 0.f;
    if (lane < NJ) z = zj[b * NJ + lane];
    else if (lane == NJ) {
      z = cs[CS_B3];
      for (int i = 0; i < n_part; i++) z += zg_part[(int64_t)i * BP + b];
    }
    const float sg = 1.f / (1.f + expf(-z));
    const bool on = b < B && lane <= NJ;
    const float d = on ? gscale * (sg - 1.f) * sg * (1.f - sg) : 0.f;
    if (lane < NJ) dzj[b * NJ + lane] = d;
    else if (lane == NJ) dzg[b] = d;
    l = on ? (sg - 1.f) * (sg - 1.f) : 0.f;
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  }
  if (lane == 0) red[warp] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < HEAD_WARPS; i++) t += red[i];
    loss_part[blockIdx.x] = t;
  }
}

// thread = (pose, joint): back through the joint head and the two 1x1 convs -> dx6c[b][j*6+i].
// The ReLU masks come from critic_pre_kernel (no forward recompute).
__global__ void __launch_bounds__(256)
critic_post_kernel(const float* __restrict__ cs, const uint2* __restrict__ masks,
                   const float* __restrict__ dh, const float* __restrict__ dzj, int64_t B,
                   float* __restrict__ dx6c, const float* __restrict__ zj, const float* __restrict__ zg_part,
                   int n_part, int64_t BP, float gscale, float* __restrict__ loss_part) {
  __shared__ float sw[CS_HW];
  __shared__ float shw[NJ * 33];
  __shared__ float red[8];
  for (int i = threadIdx.x; i < CS_HW; i += blockDim.x) sw[i] = cs[i];
  for (int i = threadIdx.x; i < NJ * 32; i += blockDim.x) shw[(i >> 5) * 33 + (i & 31)] = cs[CS_HW + i];
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < B * NJ;
  const int j = (int)(idx % NJ);
  float dj = 0.f;
  if (zj != nullptr) {
    // head-less chain: the joint heads' sigmoid / loss / dL/dlogit happen here (and the global head's loss term
    // on the joint-0 thread of every frame); the global head's dL/dlogit is applied inside the backward GEMM
    float l = 0.f;
    if (live) {
      const float sj = 1.f / (1.f + expf(-zj[idx]));
      dj = gscale * (sj - 1.f) * sj * (1.f - sj);
      l = (sj - 1.f) * (sj - 1.f);
      if (j == 0) {
        const int64_t b = idx / NJ;
        float z = cs[CS_B3];
        for (int i = 0; i < n_part; i++) z += zg_part[(int64_t)i * BP + b];
        const float sg = 1.f / (1.f + expf(-z));
        l += (sg - 1.f) * (sg - 1.f);
      }
    }
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
      loss_part[blockIdx.x] = t;
    }
  } else if (live) {
    dj = dzj[idx];
  }
  if (!live) return;
  const uint2 mk = masks[idx];
  float d[32];
  const float4* dh4 = reinterpret_cast<const float4*>(dh + idx * 32);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 t = dh4[q];
    d[4 * q + 0] = t.x; d[4 * q + 1] = t.y; d[4 * q + 2] = t.z; d[4 * q + 3] = t.w;
  }
#pragma unroll
  for (int c = 0; c < 32; c++) d[c] = ((mk.y >> c) & 1u) ? fmaf(dj, shw[j * 33 + c], d[c]) : 0.f;
  float dh1[32];
#pragma unroll
  for (int k = 0; k < 32; k++) dh1[k] = 0.f;
#pragma unroll
  for (int c = 0; c < 32; c++) {           // (fully unrolled: a partial unroll indexes d[] dynamically and spills it)
#pragma unroll
    for (int k = 0; k < 32; k++) dh1[k] = fmaf(sw[CS_C2W + c * 32 + k], d[c], dh1[k]);
  }
  float dx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 32; k++) {
    const float t = ((mk.x >> k) & 1u) ? dh1[k] : 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) dx[i] = fmaf(sw[CS_C1W + k * 6 + i], t, dx[i]);
  }
  for (int i = 0; i < 6; i++) dx6c[idx * 6 + i] = dx[i];
}

// Shape critic (scripts/discriminator.py:57-74; optimize.py:244,249-250): 10 -> 10 -> ReLU -> 5 -> ReLU -> 1
// -> sigmoid, loss mean((sigma - 1)^2) over the batch.  171 parameters: one thread per pose does the
// forward and the input gradient.  sc layout: W0[10][10] b0[10] W1[5][10] b1[5] W2[5] b2[1].
__global__ void __launch_bounds__(128)
shape_critic_kernel(const float* __restrict__ sc, const float* __restrict__ betas, int64_t B, float gscale,
                    float* __restrict__ dbeta_s, float* __restrict__ loss_part, float* __restrict__ scores_out) {
  __shared__ float sw[171];
  __shared__ float red[4];
  for (int i = threadIdx.x; i < 171; i += blockDim.x) sw[i] = sc[i];
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (b < B) {
    float x[NB], h0[10], h1[5];
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = betas[b * NB + i];
#pragma unroll
    for (int i = 0; i < 10; i++) {
      float a = sw[100 + i];
#pragma unroll
      for (int k = 0; k < 10; k++) a = fmaf(sw[i * 10 + k], x[k], a);
      h0[i] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
      float a = sw[160 + i];
#pragma unroll
      for (int k = 0; k < 10; k++) a = fmaf(sw[110 + i * 10 + k], h0[k], a);
      h1[i] = fmaxf(a, 0.f);
    }
    float z = sw[170];
#pragma unroll
    for (int k = 0; k < 5; k++) z = fmaf(sw[165 + k], h1[k], z);
    const float sg = 1.f / (1.f + expf(-z));
    l = (sg - 1.f) * (sg - 1.f);
    if (scores_out != nullptr) scores_out[b] = sg;
    if (dbeta_s != nullptr) {   // nullptr: inference only
    const float dz = gscale * (sg - 1.f) * sg * (1.f - sg);
    float d0[10];
#pragma unroll
    for (int k = 0; k < 10; k++) d0[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const float d1 = h1[i] > 0.f ? dz * sw[165 + i] : 0.f;
#pragma unroll
      for (int k = 0; k < 10; k++) d0[k] = fmaf(sw[110 + i * 10 + k], d1, d0[k]);
    }
    float dx[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) dx[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const float d = h0[i] > 0.f ? d0[i] : 0.f;
#pragma unroll
      for (int k = 0; k < 10; k++) dx[k] = fmaf(sw[i * 10 + k], d, dx[k]);
    }
#pragma unroll
    for (int k = 0; k < NB; k++) dbeta_s[b * NB + k] = dx[k];
    }
  }
  if (loss_part == nullptr) return;   // uniform over the grid
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 4; i++) t += red[i];
    loss_part[blockIdx.x] = t;
  }
}

int launch_shape_critic(const JrrModel* m, const Workspace& w, const float* betas, int64_t B_logical, cudaStream_t st) {
  shape_critic_kernel<<<(unsigned)(w.BP / 128), 128, 0, st>>>(
      m->shape_critic, betas, w.B, m->w_shape * 2.f / (float)B_logical, w.dbeta_s, w.shape_part, nullptr);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_shape_critic_scores(const JrrModel* m, int64_t B, const float* betas, float* scores_out, cudaStream_t st) {
  shape_critic_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(m->shape_critic, betas, B, 0.f, nullptr, nullptr,
                                                                  scores_out);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_critic_pre(const JrrModel* m, const Workspace& w, const float* x6, cudaStream_t st, bool want_zj) {
  const int64_t n = w.BP * NJ;
  critic_pre_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->critic_small, x6, w.B, w.BP,
                                                                w.h_hi, (want_zj && m->critic_ts) ? nullptr : w.h_lo,
                                                                want_zj ? w.zj : nullptr, w.cmask);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_critic_head_light(const JrrModel* m, Workspace& w, int64_t B_logical, float w_pose, cudaStream_t st) {
  const unsigned nblk = (unsigned)((w.BP + HEAD_WARPS - 1) / HEAD_WARPS);
  const float gscale = w_pose * 2.f / (25.f * (float)B_logical);
  w.n_pose_part = (int)nblk;
  critic_head_light_kernel<<<nblk, HEAD_WARPS * 32, 0, st>>>(m->critic_small, w.zj, w.zg_part, C_Z / 128, w.B, w.BP,
                                                            gscale, w.dzj, w.dzg, w.loss_part + LOSS_PART_POSE);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_critic_head(const JrrModel* m, Workspace& w, int64_t B_logical, float w_pose,
                       float* scores_out, bool want_grad, cudaStream_t st, float target, float* dzg) {
  const unsigned nblk = (unsigned)((w.BP + HEAD_WARPS - 1) / HEAD_WARPS);
  const float gscale = w_pose * 2.f / (25.f * (float)B_logical);
  w.n_pose_part = (int)nblk;
  critic_head_kernel<<<nblk, HEAD_WARPS * 32, 0, st>>>(
      m->critic_small, w.h_hi, w.h_lo, w.z2_hi, w.z2_lo, w.B, w.BP, gscale, target, scores_out,
      want_grad ? w.dz2_hi : nullptr, w.dz2_lo, w.dzj, dzg, w.loss_part + LOSS_PART_POSE);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int launch_critic_post(const JrrModel* m, Workspace& w, const float* x6, cudaStream_t st, bool with_head,
                       int64_t B_logical, float w_pose) {
  const int64_t n = w.B * NJ;
  (void)x6;
  const unsigned nblk = (unsigned)((n + 255) / 256);
  if (with_head) w.n_pose_part = (int)nblk;
  critic_post_kernel<<<nblk, 256, 0, st>>>(m->critic_small, w.cmask, w.dh, w.dzj, w.B, w.dx6c,
                                           with_head ? w.zj : nullptr, w.zg_part, C_Z / 128, w.BP,
                                           w_pose * 2.f / (25.f * (float)B_logical), w.loss_part + LOSS_PART_POSE);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

// The four wide GEMMs around the small kernels.  head_fused: the second layer's epilogue also does
// the global head (logit partials + its masked gradient row), see EPI_BIAS_RELU_HEAD.
int critic_forward_gemms(const JrrModel* m, const Workspace& w, cudaStream_t st, bool head_fused) {
  const bool ts = head_fused && m->critic_ts;    // plain fp32 activations, staged through tensor memory by the GEMM
  GemmDesc g{};
  g.a_via_tmem = ts;
  g.A_hi = w.h_hi; g.A_lo = w.h_lo; g.lda = C_H;
  g.B_hi = m->W1_hi; g.B_lo = m->W1_lo; g.ldb = C_H;
  g.M = w.BP; g.N = C_Z; g.K = C_H; g.ksplit = 1; g.epi = EPI_BIAS_RELU_SPLIT;
  g.out0 = w.z1_hi; g.out1 = w.z1_lo; g.ldo = C_Z; g.bias = m->critic_small + CS_B1;
  g.mask_bits_out = ts ? w.zmask : nullptr;      // layer-1 ReLU mask for the backward epilogue
  int rc = launch_gemm(m, g, st);
  if (rc) return rc;
  g.mask_bits_out = nullptr;
  g.A_hi = w.z1_hi; g.A_lo = w.z1_lo; g.lda = C_Z;
  g.B_hi = m->W2_hi; g.B_lo = m->W2_lo; g.ldb = C_Z;
  g.K = C_Z; g.out0 = w.z2_hi; g.out1 = w.z2_lo; g.bias = m->critic_small + CS_B2;
  if (head_fused) {
    g.epi = EPI_BIAS_RELU_HEAD;
    g.out0 = w.dz2_hi; g.out1 = w.dz2_lo;          // (z2 > 0) * w3, the head's gradient row up to dL/dlogit
    g.vec = m->critic_small + CS_W3; g.out2 = w.zg_part;
    // ... or only the mask z2 > 0 as bits, when the backward GEMM can take it as a 0/1 operand against diag(w3) W2
    if (ts && gemm_pair_bits_available(m, w.BP)) g.mask_bits_out = w.zmask2;
  }
  return launch_gemm(m, g, st);
}

int critic_backward_gemms(const JrrModel* m, const Workspace& w, cudaStream_t st, const float* rowscale,
                          float headless_gscale) {
  const bool ts = (rowscale != nullptr || headless_gscale != 0.f) && m->critic_ts;
  GemmDesc g{};
  g.a_via_tmem = ts;
  // dz1 = (dz2 . W2) * [z1 > 0]   (head_fused: dz2 rows still lack their scalar dL/dlogit = rowscale; head-less:
  // the epilogue derives that scalar from the logit partials itself)
  g.rowscale = rowscale;
  if (headless_gscale != 0.f) {
    g.logit_part = w.zg_part; g.n_logit_part = C_Z / 128; g.logit_bias = m->critic_small + CS_B3;
    g.logit_gscale = headless_gscale; g.rows_valid = w.B;
  }
  g.A_hi = w.dz2_hi; g.A_lo = w.dz2_lo; g.lda = C_Z;
  g.B_hi = m->W2t_hi; g.B_lo = m->W2t_lo; g.ldb = C_Z;
  if (ts && gemm_pair_bits_available(m, w.BP)) {     // (the forward wrote the mask bits, see critic_forward_gemms)
    g.a_bits = w.zmask2;
    g.B_hi = m->W2tw_hi; g.B_lo = m->W2tw_lo;
  }
  g.M = w.BP; g.N = C_Z; g.K = C_Z; g.ksplit = 1; g.epi = EPI_MASK_SPLIT;
  g.out0 = w.dz1_hi; g.out1 = w.dz1_lo; g.ldo = C_Z; g.mask = w.z1_hi; g.ldmask = C_Z;
  g.mask_bits = ts ? w.zmask : nullptr;
  int rc = launch_gemm(m, g, st);
  if (rc) return rc;
  g.mask_bits = nullptr;
  g.a_bits = nullptr;
  // dh = dz1 . W1
  g.A_hi = w.dz1_hi; g.A_lo = w.dz1_lo; g.lda = C_Z;
  g.B_hi = m->W1t_hi; g.B_lo = m->W1t_lo; g.ldb = C_Z;
  g.N = C_H; g.K = C_Z; g.epi = EPI_STORE_SPLITK; g.out0 = w.dh; g.out1 = nullptr; g.ldo = C_H;
  g.mask = nullptr; g.rowscale = nullptr; g.logit_part = nullptr;
  return launch_gemm(m, g, st);
}

}  // namespace jrr
