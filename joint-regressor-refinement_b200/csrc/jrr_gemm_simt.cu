// Plain fp32 SIMT GEMM, C[m][n] = sum_k A[m][k]*B[n][k] with the same operand format
// (tf32 hi/lo pairs, summed on load) and epilogues as the tcgen05 kernel.  It exists to
// validate the tensor-core kernel on the device (JrrModelDesc.gemm_impl = 1); the product
// path is jrr_gemm_tc.cu.
#include "jrr_internal.cuh"

namespace jrr {

constexpr int SBM = 128, SBN = 128, SBK = 8;

__device__ __forceinline__ float tf32_hi_s(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(GemmDesc g) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  const int split = blockIdx.z;
  const int64_t kbase = (int64_t)split * g.K;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const int lr = tid >> 1, lk = (tid & 1) * 4;
  for (int64_t k0 = 0; k0 < g.K; k0 += SBK) {
    {
      int64_t m = m0 + lr;
      float4 h = make_float4(0, 0, 0, 0), l = h;
      if (m < g.M) {
        h = *reinterpret_cast<const float4*>(g.A_hi + m * g.lda + kbase + k0 + lk);
        l = *reinterpret_cast<const float4*>(g.A_lo + m * g.lda + kbase + k0 + lk);
      }
      As[lk + 0][lr] = h.x + l.x; As[lk + 1][lr] = h.y + l.y;
      As[lk + 2][lr] = h.z + l.z; As[lk + 3][lr] = h.w + l.w;
      int64_t n = n0 + lr;
      h = make_float4(0, 0, 0, 0); l = h;
      if (n < g.N) {
        h = *reinterpret_cast<const float4*>(g.B_hi + n * g.ldb + kbase + k0 + lk);
        l = *reinterpret_cast<const float4*>(g.B_lo + n * g.ldb + kbase + k0 + lk);
      }
      Bs[lk + 0][lr] = h.x + l.x; Bs[lk + 1][lr] = h.y + l.y;
      Bs[lk + 2][lr] = h.z + l.z; Bs[lk + 3][lr] = h.w + l.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; k++) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int64_t mt = m0 + ty * 8, nt = n0 + tx * 8;
  if (EPI == EPI_STORE_T) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int64_t n = nt + j;
      if (n >= g.N) continue;
      float* dst = g.out0 + n * g.ldo + mt;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int64_t m = mt + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int64_t n = nt + j;
        if (n >= g.N) continue;
        float v = acc[i][j];
        if (EPI == EPI_BIAS_RELU_SPLIT) {
          v = fmaxf(v + g.bias[n], 0.f);
          float hi = tf32_hi_s(v);
          g.out0[m * g.ldo + n] = hi;
          g.out1[m * g.ldo + n] = tf32_hi_s(v - hi);
        } else if (EPI == EPI_MASK_SPLIT) {
          v = g.mask[m * g.ldmask + n] > 0.f ? v * (g.rowscale != nullptr ? g.rowscale[m] : 1.f) : 0.f;
          float hi = tf32_hi_s(v);
          g.out0[m * g.ldo + n] = hi;
          g.out1[m * g.ldo + n] = tf32_hi_s(v - hi);
        } else {
          g.out0[((int64_t)split * g.M + m) * g.ldo + n] = v;
        }
      }
    }
  }
}

int launch_gemm_simt(const GemmDesc& g, cudaStream_t st) {
  if (g.M % SBM != 0 || g.K % SBK != 0) return fail(JRR_ERR_INVALID, "simt gemm: M%128 or K%8");
  dim3 grid((unsigned)((g.N + SBN - 1) / SBN), (unsigned)(g.M / SBM), (unsigned)g.ksplit), block(256);
  switch (g.epi) {
    case EPI_STORE_T: gemm_simt_kernel<EPI_STORE_T><<<grid, block, 0, st>>>(g); break;
    case EPI_BIAS_RELU_SPLIT: gemm_simt_kernel<EPI_BIAS_RELU_SPLIT><<<grid, block, 0, st>>>(g); break;
    case EPI_MASK_SPLIT: gemm_simt_kernel<EPI_MASK_SPLIT><<<grid, block, 0, st>>>(g); break;
    case EPI_STORE_SPLITK: gemm_simt_kernel<EPI_STORE_SPLITK><<<grid, block, 0, st>>>(g); break;
    default: return fail(JRR_ERR_INVALID, "simt gemm: epilogue");
  }
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

}  // namespace jrr
