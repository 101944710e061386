// Diagnostic only (benchmarks/tma_probe.py): raw TMA delivery rate of one SM / of the chip for the box shapes the
// GEMM kernels use -- a producer lane issues `boxes` 2-D tile loads (32 fp32 = 128 B wide, SWIZZLE_128B) per stage into a
// ring of `stages` stages, a consumer lane frees each stage as soon as it has landed.  Not part of the reference interface.
#include "jrr_internal.cuh"
#include "jrr_tc.cuh"

namespace jrr {

__global__ void __launch_bounds__(64, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map, int iters, int boxes, int stages, int box_bytes, int shared_tiles,
                 int rows_total, int box_rows, int k_blocks, int dwell) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)stages * boxes * box_bytes);
  uint64_t* empty_bar = full_bar + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; it++) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_expect_tx(&full_bar[stage], (uint32_t)(boxes * box_bytes));
      for (int b = 0; b < boxes; b++) {
        const int tile = shared_tiles ? b : (int)blockIdx.x * boxes + b;
        const int row = (tile * box_rows) % rows_total;
        tma_load_2d(&map, &full_bar[stage], smem + ((size_t)stage * boxes + b) * box_bytes, (it % k_blocks) * 32, row);
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; it++) {
      mbar_wait(&full_bar[stage], phase);
      if (dwell > 0) __nanosleep(dwell);
      mbar_arrive(&empty_bar[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
}

}  // namespace jrr

using namespace jrr;

extern "C" int jrr_debug_tma_probe(const float* src, int64_t rows, int64_t cols, int box_rows, int boxes, int stages,
                                   int shared_tiles, int iters, int grid, int dwell_ns, void* stream) {
  if (!src || rows <= 0 || cols % 32 != 0 || box_rows <= 0 || box_rows > 256 || boxes <= 0 || stages <= 0)
    return fail(JRR_ERR_INVALID, "tma probe: bad argument");
  const int box_bytes = box_rows * 128;
  const size_t smem = (size_t)stages * boxes * box_bytes + 1024 + 16 * stages + 64;
  if (smem > 227 * 1024) return fail(JRR_ERR_INVALID, "tma probe: ring does not fit shared memory");
  CUtensorMap map;
  if (int rc = make_tensor_map_2d(&map, src, rows, cols, cols, box_rows)) return rc;
  JRR_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  reset_launch_count();
  tma_probe_kernel<<<grid, 64, smem, (cudaStream_t)stream>>>(map, iters, boxes, stages, box_bytes, shared_tiles, (int)rows,
                                                             box_rows, (int)(cols / 32), dwell_ns);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}
