// On-device evaluation metrics: MPJPE and Procrustes-aligned MPJPE of the 17 regressed joints.
// One thread per frame; the similarity alignment needs the SVD of a 3x3 cross-covariance, done as a
// cyclic Jacobi eigen-decomposition of K^T K in double precision (it is a metric, not a hot loop).
//
// Replaces: utils.evaluate (scripts/utils.py:117-145) and batch_compute_similarity_transform_torch
// (scripts/eval_utils.py:7-58, torch.svd + bmm), as used by scripts/optimize.py:314-321 and
// scripts/test.py:110-120.
#include "jrr_internal.cuh"

namespace jrr {

__device__ void jacobi_eig3(double A[3][3], double Vm[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Vm[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {   // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {   // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = Vm[k][p], vkq = Vm[k][q];
          Vm[k][p] = c * vkp - s * vkq;
          Vm[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

constexpr int EV_THREADS = 128;
__global__ void __launch_bounds__(EV_THREADS)
evaluate_kernel(const float* __restrict__ pred, const float* __restrict__ target_mm, int64_t B,
                float* __restrict__ per_frame, double* __restrict__ part) {
  __shared__ double red[2][EV_THREADS / 32];
  const int64_t b = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x;
  double mp = 0.0, pa = 0.0;
  if (b < B) {
    double X1[NH][3], X2[NH][3];
    // pelvis-centre both; target arrives in millimetres (utils.py:124-131)
    for (int j = 0; j < NH; j++)
      for (int c = 0; c < 3; c++) {
        X1[j][c] = (double)pred[(b * NH + j) * 3 + c] - (double)pred[b * NH * 3 + c];
        X2[j][c] = ((double)target_mm[(b * NH + j) * 3 + c] - (double)target_mm[b * NH * 3 + c]) / 1000.0;
      }
    for (int j = 0; j < NH; j++) {
      double d = 0.0;
      for (int c = 0; c < 3; c++) d += (X1[j][c] - X2[j][c]) * (X1[j][c] - X2[j][c]);
      mp += sqrt(d);
    }
    mp /= NH;
    // similarity alignment of X1 onto X2 (eval_utils.py:19-52)
    double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
    for (int j = 0; j < NH; j++)
      for (int c = 0; c < 3; c++) { mu1[c] += X1[j][c] / NH; mu2[c] += X2[j][c] / NH; }
    double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, var1 = 0.0;
    for (int j = 0; j < NH; j++)
      for (int r = 0; r < 3; r++) {
        const double a = X1[j][r] - mu1[r];
        var1 += a * a;
        for (int c = 0; c < 3; c++) K[r][c] += a * (X2[j][c] - mu2[c]);
      }
    // K = U S V^T.  V, S from the eigen-decomposition of K^T K; R = V Z U^T = V Z S^-1 V^T K^T
    double A[3][3], Vm[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[r][c] = K[0][r] * K[0][c] + K[1][r] * K[1][c] + K[2][r] * K[2][c];
    jacobi_eig3(A, Vm);
    double sv[3] = {sqrt(fmax(A[0][0], 0.0)), sqrt(fmax(A[1][1], 0.0)), sqrt(fmax(A[2][2], 0.0))};
    int smallest = 0;
    for (int i = 1; i < 3; i++)
      if (sv[i] < sv[smallest]) smallest = i;
    const double detK = K[0][0] * (K[1][1] * K[2][2] - K[1][2] * K[2][1]) - K[0][1] * (K[1][0] * K[2][2] - K[1][2] * K[2][0]) +
                        K[0][2] * (K[1][0] * K[2][1] - K[1][1] * K[2][0]);
    double z[3] = {1.0, 1.0, 1.0};
    if (detK < 0) z[smallest] = -1.0;
    // M = V diag(z/s) V^T ; R = M K^T
    double M[3][3], R[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double a = 0.0;
        for (int i = 0; i < 3; i++) a += Vm[r][i] * (z[i] / fmax(sv[i], 1e-300)) * Vm[c][i];
        M[r][c] = a;
      }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) R[r][c] = M[r][0] * K[c][0] + M[r][1] * K[c][1] + M[r][2] * K[c][2];
    const double scale = (z[0] * sv[0] + z[1] * sv[1] + z[2] * sv[2]) / var1;
    for (int j = 0; j < NH; j++) {
      double d = 0.0;
      for (int r = 0; r < 3; r++) {
        double y = mu2[r];
        for (int c = 0; c < 3; c++) y += scale * R[r][c] * (X1[j][c] - mu1[c]);
        d += (y - X2[j][r]) * (y - X2[j][r]);
      }
      pa += sqrt(d);
    }
    pa /= NH;
    if (per_frame != nullptr) { per_frame[b * 2] = (float)(mp * 1000.0); per_frame[b * 2 + 1] = (float)(pa * 1000.0); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mp += __shfl_xor_sync(0xffffffffu, mp, o);
    pa += __shfl_xor_sync(0xffffffffu, pa, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mp; red[1][threadIdx.x >> 5] = pa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < EV_THREADS / 32; i++) { a += red[0][i]; c += red[1][i]; }
    part[blockIdx.x * 2] = a;
    part[blockIdx.x * 2 + 1] = c;
  }
}

__global__ void evaluate_finish_kernel(const double* __restrict__ part, int n, int64_t B, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double a = 0.0, c = 0.0;
  for (int i = 0; i < n; i++) { a += part[i * 2]; c += part[i * 2 + 1]; }
  out[0] = (float)(a / (double)B * 1000.0);
  out[1] = (float)(c / (double)B * 1000.0);
}

}  // namespace jrr

using namespace jrr;

extern "C" int jrr_evaluate(int64_t B, const float* pred_j3d, const float* target_j3d_mm, float* out_mm,
                            float* per_frame_mm, void* scratch, size_t scratch_bytes, void* stream) {
  if (B <= 0 || !pred_j3d || !target_j3d_mm || !out_mm || !scratch) return fail(JRR_ERR_INVALID, "bad argument");
  const unsigned nblk = (unsigned)((B + EV_THREADS - 1) / EV_THREADS);
  if (scratch_bytes < (size_t)nblk * 2 * sizeof(double) || ((uintptr_t)scratch & 7))
    return fail(JRR_ERR_WORKSPACE, "jrr_evaluate scratch: 16 bytes per 128 frames, 8-byte aligned");
  reset_launch_count();
  cudaStream_t st = (cudaStream_t)stream;
  evaluate_kernel<<<nblk, EV_THREADS, 0, st>>>(pred_j3d, target_j3d_mm, B, per_frame_mm, (double*)scratch);
  JRR_LAUNCH_CHECK();
  evaluate_finish_kernel<<<1, 32, 0, st>>>((const double*)scratch, (int)nblk, B, out_mm);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}
