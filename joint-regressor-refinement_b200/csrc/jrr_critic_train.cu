// Critic training step (scripts/optimize.py:276-293): after a batch has been refined, both
// discriminators take one Adam step on  MSE(D(refined), 0) + MSE(D(initial), 1).  The two halves
// are independent sums, so the C ABI exposes "accumulate the parameter gradient of
// mean((D(x) - target)^2)" (called once per half, and once per chunk / rank), an all-reduce by the
// caller when frames are sharded, and "apply": Adam over the flat parameter vector followed by the
// refresh of the model's packed copies.
//
// Weight gradients of the two wide layers are transposed-operand GEMMs on the tensor core
// (dW2 = dz2^T z1, dW1 = dz1^T h; K = batch, split-K partials summed in fp32); everything else is
// small reductions with fixed summation order (no atomics: results are run-to-run identical).
#include "jrr_internal.cuh"

namespace jrr {

constexpr int CT_POSES = 10;              // poses per CTA of the conv weight-gradient kernel
constexpr int CT_ITEMS = CT_POSES * NJ;   // 240 (pose, joint) items
constexpr int CT_SMALL = 2072;            // conv1 W,b | conv2 W,b | 24 x (head W[32], b[1]): state_dict order
constexpr int CT_LD = 33;
constexpr int CT_SMEM = (4 * CT_ITEMS * CT_LD + CT_ITEMS * 6 + CT_ITEMS + CS_HB) * (int)sizeof(float);
constexpr int CT_ROWS = 256;              // rows per partial of the column sums
// flat (state_dict-order) offsets of the wide layers
constexpr int64_t FO_W1 = CT_SMALL;
constexpr int64_t FO_B1 = FO_W1 + (int64_t)C_Z * C_H;
constexpr int64_t FO_W2 = FO_B1 + C_Z;
constexpr int64_t FO_B2 = FO_W2 + (int64_t)C_Z * C_Z;
constexpr int64_t FO_W3 = FO_B2 + C_Z;
constexpr int64_t FO_B3 = FO_W3 + C_Z;
static_assert(FO_B3 + 1 == JRR_CRITIC_PARAMS, "flat critic layout");

// dst[c][r] = src[r][c]
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t rows, int cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; i++) tile[ty + 8 * i][tx] = src[(r0 + ty + 8 * i) * cols + c0 + tx];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) dst[(int64_t)(c0 + ty + 8 * i) * rows + r0 + tx] = tile[tx][ty + 8 * i];
}

// partial[rb][c] = sum_{r in row block rb} (hi[r][c] + lo[r][c]) * rowscale[r]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ hi, const float* __restrict__ lo, const float* __restrict__ rowscale,
              int cols, float* __restrict__ partial) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * CT_ROWS;
  float a = 0.f;
  for (int r = 0; r < CT_ROWS; r++) {
    float x = hi[(r0 + r) * cols + c];
    if (lo != nullptr) x += lo[(r0 + r) * cols + c];
    a = fmaf(x, rowscale != nullptr ? rowscale[r0 + r] : 1.f, a);
  }
  partial[(int64_t)blockIdx.y * cols + c] = a;
}

// Weight gradients of the two 1x1 convs and the 24 joint heads for CT_POSES poses per CTA.
__global__ void __launch_bounds__(256)
critic_conv_wgrad_kernel(const float* __restrict__ cs, const float* __restrict__ x6, const float* __restrict__ dh,
                         const float* __restrict__ dzj, int64_t B, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* s_h1 = sm;
  float* s_da2 = s_h1 + CT_ITEMS * CT_LD;
  float* s_da1 = s_da2 + CT_ITEMS * CT_LD;
  float* s_h2 = s_da1 + CT_ITEMS * CT_LD;
  float* s_x = s_h2 + CT_ITEMS * CT_LD;
  float* s_dj = s_x + CT_ITEMS * 6;
  float* sw = s_dj + CT_ITEMS;
  for (int i = threadIdx.x; i < CS_HB; i += blockDim.x) sw[i] = cs[i];
  __syncthreads();
  const int it = threadIdx.x;
  if (it < CT_ITEMS) {
    const int64_t idx = (int64_t)blockIdx.x * CT_ITEMS + it;
    const int j = it % NJ;
    if (idx < B * NJ) {
      float x[6];
      for (int i = 0; i < 6; i++) x[i] = x6[idx * 6 + i];
      float h1[32];
#pragma unroll
      for (int k = 0; k < 32; k++) {
        float a = sw[CS_C1B + k];
#pragma unroll
        for (int i = 0; i < 6; i++) a = fmaf(sw[CS_C1W + k * 6 + i], x[i], a);
        h1[k] = fmaxf(a, 0.f);
      }
      const float dj = dzj[idx];
      float dh1[32];
#pragma unroll
      for (int k = 0; k < 32; k++) dh1[k] = 0.f;
#pragma unroll 4
      for (int c = 0; c < 32; c++) {
        float a = sw[CS_C2B + c];
#pragma unroll
        for (int k = 0; k < 32; k++) a = fmaf(sw[CS_C2W + c * 32 + k], h1[k], a);
        const float d = a > 0.f ? dh[idx * 32 + c] + dj * sw[CS_HW + j * 32 + c] : 0.f;
        s_h2[it * CT_LD + c] = fmaxf(a, 0.f);
        s_da2[it * CT_LD + c] = d;
#pragma unroll
        for (int k = 0; k < 32; k++) dh1[k] = fmaf(sw[CS_C2W + c * 32 + k], d, dh1[k]);
      }
#pragma unroll
      for (int k = 0; k < 32; k++) {
        s_h1[it * CT_LD + k] = h1[k];
        s_da1[it * CT_LD + k] = h1[k] > 0.f ? dh1[k] : 0.f;
      }
      for (int i = 0; i < 6; i++) s_x[it * 6 + i] = x[i];
      s_dj[it] = dj;
    } else {
      for (int k = 0; k < 32; k++) {
        s_h1[it * CT_LD + k] = 0.f; s_da2[it * CT_LD + k] = 0.f;
        s_da1[it * CT_LD + k] = 0.f; s_h2[it * CT_LD + k] = 0.f;
      }
      for (int i = 0; i < 6; i++) s_x[it * 6 + i] = 0.f;
      s_dj[it] = 0.f;
    }
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * CT_SMALL;
  for (int o = threadIdx.x; o < CT_SMALL; o += blockDim.x) {
    float a = 0.f;
    if (o < CS_C1B) {                       // conv_operations.0.weight[k][i]
      const int k = o / 6, i = o % 6;
      for (int t = 0; t < CT_ITEMS; t++) a = fmaf(s_da1[t * CT_LD + k], s_x[t * 6 + i], a);
    } else if (o < CS_C2W) {                // conv_operations.0.bias[k]
      const int k = o - CS_C1B;
      for (int t = 0; t < CT_ITEMS; t++) a += s_da1[t * CT_LD + k];
    } else if (o < CS_C2B) {                // conv_operations.2.weight[c][k]
      const int c = (o - CS_C2W) >> 5, k = (o - CS_C2W) & 31;
      for (int t = 0; t < CT_ITEMS; t++) a = fmaf(s_da2[t * CT_LD + c], s_h1[t * CT_LD + k], a);
    } else if (o < CS_HW) {                 // conv_operations.2.bias[c]
      const int c = o - CS_C2B;
      for (int t = 0; t < CT_ITEMS; t++) a += s_da2[t * CT_LD + c];
    } else {                                // linears.j.weight[32], linears.j.bias[1]
      const int q = o - CS_HW, j = q / 33, c = q % 33;
      for (int p = 0; p < CT_POSES; p++) {
        const int t = p * NJ + j;
        a = c < 32 ? fmaf(s_dj[t], s_h2[t * CT_LD + c], a) : a + s_dj[t];
      }
    }
    out[o] = a;
  }
}

// G[i] += sum of the partials that belong to flat parameter i
__global__ void __launch_bounds__(256)
critic_grad_finish_kernel(const float* __restrict__ conv_part, int n_conv, const float* __restrict__ w1_part,
                          const float* __restrict__ w2_part, int ksplit, const float* __restrict__ b1_part,
                          const float* __restrict__ b2_part, const float* __restrict__ w3_part,
                          const float* __restrict__ b3_part, int n_rb, float* __restrict__ G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= JRR_CRITIC_PARAMS) return;
  float a = 0.f;
  if (i < FO_W1) {
    for (int r = 0; r < n_conv; r++) a += conv_part[(int64_t)r * CT_SMALL + i];
  } else if (i < FO_B1) {
    for (int s = 0; s < ksplit; s++) a += w1_part[(int64_t)s * C_Z * C_H + (i - FO_W1)];
  } else if (i < FO_W2) {
    for (int r = 0; r < n_rb; r++) a += b1_part[(int64_t)r * C_Z + (i - FO_B1)];
  } else if (i < FO_B2) {
    for (int s = 0; s < ksplit; s++) a += w2_part[(int64_t)s * C_Z * C_Z + (i - FO_W2)];
  } else if (i < FO_W3) {
    for (int r = 0; r < n_rb; r++) a += b2_part[(int64_t)r * C_Z + (i - FO_B2)];
  } else if (i < FO_B3) {
    for (int r = 0; r < n_rb; r++) a += w3_part[(int64_t)r * C_Z + (i - FO_W3)];
  } else {
    for (int r = 0; r < n_rb; r++) a += b3_part[r];
  }
  G[i] += a;
}

__global__ void sum_scale_add_kernel(const float* __restrict__ parts, int n, float scale, float* __restrict__ out) {
  const int lane = threadIdx.x;
  float a = 0.f;
  for (int i = lane; i < n; i += 32) a += parts[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[0] += a * scale;
}

// torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no weight decay) over a flat vector
__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ am, float* __restrict__ av,
                 const int32_t* __restrict__ step_count, float lr, int64_t n) {
  __shared__ float s_bc[2];      // double-precision bias corrections, once per block
  if (threadIdx.x == 0) {
    const int t = *step_count + 1;
    s_bc[0] = (float)sqrt(1.0 - pow(0.999, (double)t));
    s_bc[1] = (float)((double)lr / (1.0 - pow(0.9, (double)t)));
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bc2s = s_bc[0], step = s_bc[1];
  const float gi = g[i];
  const float m = 0.9f * am[i] + 0.1f * gi;
  const float v = 0.999f * av[i] + 0.001f * gi * gi;
  am[i] = m;
  av[i] = v;
  p[i] -= step * (m / (sqrtf(v) / bc2s + 1e-8f));
}

__global__ void bump_count_kernel(int32_t* c) { *c += 1; }

// ---- shape critic (171 parameters): 128 poses per CTA ------------------------------------------
constexpr int SC_LD = 41;   // x[10] h0[10] h1[5] dz[1] d1[5] d0[10]
__global__ void __launch_bounds__(128)
shape_wgrad_kernel(const float* __restrict__ sc, const float* __restrict__ betas, int64_t B, float gscale, float target,
                   float* __restrict__ partial, float* __restrict__ loss_part) {
  __shared__ float sw[JRR_SHAPE_CRITIC_PARAMS];
  __shared__ float st[128 * SC_LD];
  __shared__ float red[4];
  for (int i = threadIdx.x; i < JRR_SHAPE_CRITIC_PARAMS; i += blockDim.x) sw[i] = sc[i];
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  float* r = st + threadIdx.x * SC_LD;
  float l = 0.f;
  if (b < B) {
    float x[10], h0[10], h1[5];
#pragma unroll
    for (int i = 0; i < 10; i++) x[i] = betas[b * NB + i];
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
    l = (sg - target) * (sg - target);
    const float dz = gscale * (sg - target) * sg * (1.f - sg);
    float d0[10];
#pragma unroll
    for (int k = 0; k < 10; k++) d0[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const float d1 = h1[i] > 0.f ? dz * sw[165 + i] : 0.f;
      r[26 + i] = d1;
#pragma unroll
      for (int k = 0; k < 10; k++) d0[k] = fmaf(sw[110 + i * 10 + k], d1, d0[k]);
    }
#pragma unroll
    for (int i = 0; i < 10; i++) {
      r[i] = x[i];
      r[10 + i] = h0[i];
      r[31 + i] = h0[i] > 0.f ? d0[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) r[20 + i] = h1[i];
    r[25] = dz;
  } else {
    for (int i = 0; i < SC_LD; i++) r[i] = 0.f;
  }
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) loss_part[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
  for (int o = threadIdx.x; o < JRR_SHAPE_CRITIC_PARAMS; o += blockDim.x) {
    int ia, ib;     // columns of the staged record: gradient = sum_t st[t][ia] * st[t][ib]  (ib < 0: * 1)
    if (o < 100) { ia = 31 + o / 10; ib = o % 10; }
    else if (o < 110) { ia = 31 + (o - 100); ib = -1; }
    else if (o < 160) { ia = 26 + (o - 110) / 10; ib = 10 + (o - 110) % 10; }
    else if (o < 165) { ia = 26 + (o - 160); ib = -1; }
    else if (o < 170) { ia = 25; ib = 20 + (o - 165); }
    else { ia = 25; ib = -1; }
    float a = 0.f;
    for (int t = 0; t < 128; t++) a = ib >= 0 ? fmaf(st[t * SC_LD + ia], st[t * SC_LD + ib], a) : a + st[t * SC_LD + ia];
    partial[(int64_t)blockIdx.x * JRR_SHAPE_CRITIC_PARAMS + o] = a;
  }
}

__global__ void __launch_bounds__(256)
shape_grad_finish_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ G) {
  const int i = threadIdx.x;
  if (i >= JRR_SHAPE_CRITIC_PARAMS) return;
  float a = 0.f;
  for (int r = 0; r < nblk; r++) a += partial[(int64_t)r * JRR_SHAPE_CRITIC_PARAMS + i];
  G[i] += a;
}

// ---- host side ---------------------------------------------------------------------------------
int launch_transpose(const float* src, int64_t rows, int cols, float* dst, cudaStream_t st) {
  dim3 grid((unsigned)(cols / 32), (unsigned)(rows / 32));
  transpose_kernel<<<grid, 256, 0, st>>>(src, rows, cols, dst);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}
static int transpose(const float* src, int64_t rows, int cols, float* dst, cudaStream_t st) {
  return launch_transpose(src, rows, cols, dst, st);
}

static int colsum(const float* hi, const float* lo, const float* rowscale, int64_t rows, int cols, float* partial,
                  cudaStream_t st) {
  dim3 grid((unsigned)((cols + 255) / 256), (unsigned)(rows / CT_ROWS));
  colsum_kernel<<<grid, 256, 0, st>>>(hi, lo, rowscale, cols, partial);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int critic_grad_accumulate(JrrModel* m, Workspace& w, int64_t B_logical, const float* x6, float target, float* G_accum,
                           float* loss_accum, cudaStream_t st) {
  const int64_t BP = w.BP;
  // scratch carved out of the (idle) blend-gradient buffers of the workspace
  float* t = w.dvp_hi;                                   // [BP][20736] floats available
  float* z1T_hi = t;  t += (int64_t)C_Z * BP;
  float* z1T_lo = t;  t += (int64_t)C_Z * BP;
  float* dz2T_hi = t; t += (int64_t)C_Z * BP;
  float* dz2T_lo = t; t += (int64_t)C_Z * BP;
  float* dz1T_hi = t; t += (int64_t)C_Z * BP;
  float* dz1T_lo = t; t += (int64_t)C_Z * BP;
  float* hT_hi = t;   t += (int64_t)C_H * BP;
  float* hT_lo = t;   t += (int64_t)C_H * BP;
  const int n_rb = (int)(BP / CT_ROWS);
  int ksplit = 1;
  for (int d = 8; d >= 1; d--)
    if (n_rb % d == 0) { ksplit = d; break; }
  const int n_conv = (int)((w.B + CT_POSES - 1) / CT_POSES);
  float* u = w.dvp_lo;
  float* w1_part = u;   u += (int64_t)ksplit * C_Z * C_H;
  float* w2_part = u;   u += (int64_t)ksplit * C_Z * C_Z;
  float* b1_part = u;   u += (int64_t)n_rb * C_Z;
  float* b2_part = u;   u += (int64_t)n_rb * C_Z;
  float* w3_part = u;   u += (int64_t)n_rb * C_Z;
  float* b3_part = u;   u += n_rb;
  float* dzg = u;       u += BP;
  float* conv_part = u; u += (int64_t)n_conv * CT_SMALL;
  if ((u - w.dvp_lo) > (int64_t)NP * BP) return fail(JRR_ERR_WORKSPACE, "critic training scratch does not fit the workspace");

  if (int rc = launch_critic_pre(m, w, x6, st)) return rc;
  if (int rc = critic_forward_gemms(m, w, st)) return rc;
  // loss = mean over B_logical x 25 scores of (sigma - target)^2
  if (int rc = launch_critic_head(m, w, B_logical, 1.f, nullptr, true, st, target, dzg)) return rc;
  if (int rc = critic_backward_gemms(m, w, st)) return rc;
  // transposed copies (hi and lo transposed separately: the split stays exact)
  if (int rc = transpose(w.z1_hi, BP, C_Z, z1T_hi, st)) return rc;
  if (int rc = transpose(w.z1_lo, BP, C_Z, z1T_lo, st)) return rc;
  if (int rc = transpose(w.dz2_hi, BP, C_Z, dz2T_hi, st)) return rc;
  if (int rc = transpose(w.dz2_lo, BP, C_Z, dz2T_lo, st)) return rc;
  if (int rc = transpose(w.dz1_hi, BP, C_Z, dz1T_hi, st)) return rc;
  if (int rc = transpose(w.dz1_lo, BP, C_Z, dz1T_lo, st)) return rc;
  if (int rc = transpose(w.h_hi, BP, C_H, hT_hi, st)) return rc;
  if (int rc = transpose(w.h_lo, BP, C_H, hT_lo, st)) return rc;
  GemmDesc g{};
  // dW2[out][in] = sum_b dz2[b][out] z1[b][in]
  g.A_hi = dz2T_hi; g.A_lo = dz2T_lo; g.lda = BP;
  g.B_hi = z1T_hi; g.B_lo = z1T_lo; g.ldb = BP;
  g.M = C_Z; g.N = C_Z; g.K = BP / ksplit; g.ksplit = ksplit; g.epi = EPI_STORE_SPLITK;
  g.out0 = w2_part; g.ldo = C_Z;
  if (int rc = launch_gemm(m, g, st)) return rc;
  // dW1[out][in] = sum_b dz1[b][out] h[b][in]
  g.A_hi = dz1T_hi; g.A_lo = dz1T_lo;
  g.B_hi = hT_hi; g.B_lo = hT_lo;
  g.N = C_H; g.out0 = w1_part; g.ldo = C_H;
  if (int rc = launch_gemm(m, g, st)) return rc;
  if (int rc = colsum(w.dz1_hi, w.dz1_lo, nullptr, BP, C_Z, b1_part, st)) return rc;
  if (int rc = colsum(w.dz2_hi, w.dz2_lo, nullptr, BP, C_Z, b2_part, st)) return rc;
  if (int rc = colsum(w.z2_hi, w.z2_lo, dzg, BP, C_Z, w3_part, st)) return rc;
  if (int rc = colsum(dzg, nullptr, nullptr, BP, 1, b3_part, st)) return rc;
  JRR_CUDA(cudaFuncSetAttribute(critic_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  critic_conv_wgrad_kernel<<<(unsigned)n_conv, 256, CT_SMEM, st>>>(m->critic_small, x6, w.dh, w.dzj, w.B, conv_part);
  JRR_LAUNCH_CHECK();
  critic_grad_finish_kernel<<<(unsigned)((JRR_CRITIC_PARAMS + 255) / 256), 256, 0, st>>>(
      conv_part, n_conv, w1_part, w2_part, ksplit, b1_part, b2_part, w3_part, b3_part, n_rb, G_accum);
  JRR_LAUNCH_CHECK();
  if (loss_accum != nullptr) {
    sum_scale_add_kernel<<<1, 32, 0, st>>>(w.loss_part + LOSS_PART_POSE, w.n_pose_part, 1.f / (25.f * (float)B_logical),
                                           loss_accum);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

int launch_adam_flat(float* p, const float* g, float* am, float* av, int32_t* step_count, float lr, int64_t n,
                     cudaStream_t st) {
  adam_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, am, av, step_count, lr, n);
  JRR_LAUNCH_CHECK();
  bump_count_kernel<<<1, 1, 0, st>>>(step_count);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

int shape_critic_grad_accumulate(JrrModel* m, Workspace& w, int64_t B_logical, const float* betas, float target,
                                 float* G_accum, float* loss_accum, cudaStream_t st) {
  const int nblk = (int)((w.B + 127) / 128);
  float* partial = w.dvp_lo;
  float* lpart = partial + (int64_t)nblk * JRR_SHAPE_CRITIC_PARAMS;
  shape_wgrad_kernel<<<(unsigned)nblk, 128, 0, st>>>(m->shape_critic, betas, w.B, 2.f / (float)B_logical, target, partial,
                                                    lpart);
  JRR_LAUNCH_CHECK();
  shape_grad_finish_kernel<<<1, 256, 0, st>>>(partial, nblk, G_accum);
  JRR_LAUNCH_CHECK();
  if (loss_accum != nullptr) {
    sum_scale_add_kernel<<<1, 32, 0, st>>>(lpart, nblk, 1.f / (float)B_logical, loss_accum);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

}  // namespace jrr
