// 3xTF32 tensor-core GEMM for sm_100a: C[m][n] = sum_k A[m][k] * B[n][k] with both operands
// K-major and pre-split into tf32 hi/lo pairs; D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo is
// accumulated in fp32 in TMEM, which holds fp32 accuracy (the dropped lo.lo term is 2^-22).
//
//   * persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA
//     issuer (one elected lane, tcgen05.mma.cta_group::1.kind::tf32), warps 2-5 = epilogue
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, 32 fp32 = 128 B per row)
//     into an mbarrier ring; UMMA smem descriptors walk the swizzle atom in 32-byte steps
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1
//   * fused epilogues: transposed (pose-contiguous) store, bias+ReLU+tf32-split,
//     ReLU-mask+tf32-split, split-K partial store
//
// Used for: the augmented pose/shape blend [B,224]x[224,20736] and its transpose-side
// backward (smplx.lbs blend_shapes + pose_feature @ posedirs, reached through
// scripts/smpl.py:72-74), and the 768->1024->1024 layers of the pose critic and their input
// gradients (scripts/discriminator.py:24-30,41).
#include "jrr_internal.cuh"
#include "jrr_tc.cuh"

namespace jrr {

constexpr int BM = 128;
constexpr int BK = 32;  // fp32 per stage row = 128 bytes = one swizzle atom
constexpr int TC_THREADS = 192;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int TC_THREADS_TS = 320;       // + warps 6-9: A producers of the TS variant (smem fp32 tile -> registers -> TMEM)

struct TcParams {
  int64_t M, N, K;     // K = per-split extent
  int ksplit;
  int m_tiles, n_tiles;
  float* out0; float* out1; int64_t ldo;
  const float* bias;
  const float* mask; int64_t ldmask;
  const uint32_t* mask_bits; uint32_t* mask_bits_out;   // ReLU masks as bits [M][N/32]
  const float* logit_part; int n_logit_part; const float* logit_bias; float logit_gscale; int64_t rows_valid;
  const float* A; int64_t lda;   // TS: plain fp32 A
  const float* rowscale;   // EPI_MASK_SPLIT: multiply row m by rowscale[m] (or nullptr)
  const float* vec;        // EPI_BIAS_RELU_HEAD: w3[N]
  float* out2;             // EPI_BIAS_RELU_HEAD: zg_part[n_tile][M]
  const uint32_t* a_bits;  // CTA-pair kernel, ABITS: the 0/1 A operand as bits [M][K_total / 32]
  long long* prof;         // diagnostic only (jrr_debug_gemm + JRR_GEMM_PROF): per-CTA role timers, 16 counters each
  int probe;               // diagnostic only (JRR_GEMM_PROBE, benchmarks/gemm_probe.py): bit 0 = MMAs do not wait for the
                           // A producers (garbage A; shows the loop's pace without the smem->TMEM staging chain),
                           // bit 1 = skip the A_hi.B_lo MMA, bit 2 = skip the A_lo.B_hi MMA as well (results wrong)
};

template <int BN>
struct TcCfg {
  static constexpr int STAGES = (BN <= 128) ? 3 : 2;
  static constexpr int A_BYTES = BM * BK * 4;       // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;  // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int A_TMEM_COL = 2 * ACC_STRIDE;   // TS: A staging ring, 64 columns (hi 32 | lo 32) per stage
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int STAGES_TS = 4;                        // TS: [A raw][B_hi][B_lo] = 48 KB at BN = 128
  static constexpr int STAGE_BYTES_TS = A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES_TS = STAGES_TS * STAGE_BYTES_TS + 1024 + 256;
};

// 32 consecutive outputs of row m.  Direct form: eight 16-byte stores per thread -- the 32 lanes of a warp hold 32
// different rows, so every store instruction touches 32 lines (measured: 8192 LSU wavefronts per 256x128 tile, ~9400
// exposed cycles at the end of every critic GEMM).  Staged form (`stage` != nullptr): the warp writes its 32x32 block into a
// SWIZZLE_128B shared-memory box (conflict-free per quarter warp) and one lane hands it to the TMA store unit, which writes
// full 128-byte rows and clips rows past M; `row0` = first row of the warp's block (all 32 rows are in or out together).
__device__ __forceinline__ void tc_store_row32(float* out, int64_t ldo, int64_t m, int64_t n0, const float* x,
                                               float* stage, const CUtensorMap* mapO, int64_t row0, int lane) {
  if (stage == nullptr) {
    float4* o = reinterpret_cast<float4*>(out + m * ldo + n0);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    return;
  }
  if (lane == 0) tma_store_wait_read();          // the previous block has left the staging box
  __syncwarp();
  uint8_t* rowp = reinterpret_cast<uint8_t*>(stage) + lane * 128;
#pragma unroll
  for (int i = 0; i < 8; i++)
    *reinterpret_cast<float4*>(rowp + ((i ^ (lane & 7)) << 4)) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(mapO, stage, (int)n0, (int)row0);
    tma_store_commit();
  }
}

// One accumulator row (thread = row m = TMEM lane, BN columns starting at TMEM address `trow`, output columns starting at
// `n_base`) through the fused epilogue.
template <int BN, int EPI, bool TS>
__device__ __forceinline__ void tc_epilogue_row(const TcParams& p, const uint32_t trow, const int64_t m, const int64_t n_base,
                                                const int split, float* stage = nullptr, const CUtensorMap* mapO = nullptr) {
  static_assert(BN % 32 == 0, "epilogue rows are processed 32 columns at a time");
  const int lane = threadIdx.x & 31;
  const int64_t row0 = m - lane;
  float zsum = 0.f;   // EPI_BIAS_RELU_HEAD: this row's share of the global head's logit
  float rs = 1.f;
  if (EPI == EPI_MASK_SPLIT && p.rowscale != nullptr && m < p.M) rs = p.rowscale[m];
  if (EPI == EPI_MASK_SPLIT && p.logit_part != nullptr && m < p.M) {
    // dL/dlogit of the global critic head for this row, from the forward epilogue's logit partials
    float z = __ldg(p.logit_bias);
    for (int i = 0; i < p.n_logit_part; i++) z += p.logit_part[(int64_t)i * p.M + m];
    const float sg = 1.f / (1.f + expf(-z));
    rs = m < p.rows_valid ? p.logit_gscale * (sg - 1.f) * sg * (1.f - sg) : 0.f;
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; c++) {
    float v[32];
    tc_ld32(trow + c * 32, v);
    const int64_t n0 = n_base + c * 32;
    if (EPI == EPI_BIAS_RELU_HEAD) {
      // x = relu(acc + b); logit partial sum x.w3; and the head's masked gradient row
      // (x > 0 ? w3 : 0) as a tf32 hi/lo pair -- the per-row scalar dL/dlogit is applied by the
      // EPILOGUE of the backward GEMM (rowscale), so no kernel ever reads the activations back
      if (TS && p.mask_bits_out != nullptr) {
        // the consumer of this layer's gradient row (x > 0 ? w3 : 0) takes the ReLU MASK as a 0/1 operand against
        // diag(w3) W2 (a_bits): one bit per element leaves instead of a 4-byte value
        if (m < p.M && n0 < p.N) {
          uint32_t bits = 0;
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float x = fmaxf(v[i] + __ldg(p.bias + n0 + i), 0.f);
            zsum = fmaf(x, __ldg(p.vec + n0 + i), zsum);
            bits |= (x > 0.f ? 1u : 0u) << i;
          }
          p.mask_bits_out[m * (p.N >> 5) + (n0 >> 5)] = bits;
        }
      } else if (m < p.M && n0 < p.N) {
        float hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float x = fmaxf(v[i] + __ldg(p.bias + n0 + i), 0.f);
          const float w = __ldg(p.vec + n0 + i);
          zsum = fmaf(x, w, zsum);
          if (TS) {
            hi[i] = x > 0.f ? w : 0.f;
          } else {
            const float wh = tf32_hi_g(w);
            hi[i] = x > 0.f ? wh : 0.f;
            lo[i] = x > 0.f ? tf32_hi_g(w - wh) : 0.f;
          }
        }
        tc_store_row32(p.out0, p.ldo, m, n0, hi, TS ? stage : nullptr, mapO, row0, lane);
        if (!TS) tc_store_row32(p.out1, p.ldo, m, n0, lo, nullptr, nullptr, row0, lane);
      }
    } else if (EPI == EPI_STORE_T) {
      // out[n][m]: lanes hold consecutive m -> one 128-byte line per column
      if (m < p.M) {
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (n0 + i < p.N) p.out0[(n0 + i) * p.ldo + m] = v[i];
      }
    } else if (EPI == EPI_BIAS_RELU_SPLIT || EPI == EPI_MASK_SPLIT) {
      if (m < p.M && n0 < p.N) {
        float hi[32], lo[32];
        // ReLU masks can travel as one bit per element (row m, word n0 / 32) between the layer's forward
        // epilogue and the backward epilogue, instead of re-reading the fp32 activations
        uint32_t bits = 0;
        if (EPI == EPI_MASK_SPLIT && p.mask_bits != nullptr) bits = p.mask_bits[m * (p.N >> 5) + (n0 >> 5)];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          float x = v[i];
          if (EPI == EPI_BIAS_RELU_SPLIT) {
            x = fmaxf(x + __ldg(p.bias + n0 + i), 0.f);
            bits |= (x > 0.f ? 1u : 0u) << i;
          } else if (p.mask_bits != nullptr) {
            x = ((bits >> i) & 1u) ? x * rs : 0.f;
          } else {
            x = (p.mask[m * p.ldmask + n0 + i] > 0.f) ? x * rs : 0.f;
          }
          if (TS) {
            hi[i] = x;
          } else {
            hi[i] = tf32_hi_g(x);
            lo[i] = tf32_hi_g(x - hi[i]);
          }
        }
        if (EPI == EPI_BIAS_RELU_SPLIT && p.mask_bits_out != nullptr) p.mask_bits_out[m * (p.N >> 5) + (n0 >> 5)] = bits;
        tc_store_row32(p.out0, p.ldo, m, n0, hi, TS ? stage : nullptr, mapO, row0, lane);
        if (!TS) tc_store_row32(p.out1, p.ldo, m, n0, lo, nullptr, nullptr, row0, lane);
      }
    } else {
      // split-K partial s lives at rows [s*M, (s+1)*M) of the output (the store map of the staged form spans ksplit*M rows)
      if (m < p.M && n0 < p.N)
        tc_store_row32(p.out0, p.ldo, (int64_t)split * p.M + m, n0, v, stage, mapO, (int64_t)split * p.M + row0, lane);
    }
  }
  if (EPI == EPI_BIAS_RELU_HEAD && m < p.M) p.out2[(n_base / 128) * p.M + m] = zsum;   // one logit partial per 128 columns (callers use BN = 128 here)
}

// TS = false: both operands arrive pre-split through TMA (four tiles per stage) and are read from
// shared memory by the tensor core -- at 128x128 tiles the 3xTF32 scheme then reads 24 KB of operands
// per 8-column K step and is bound by shared-memory bandwidth (~60 % of the tensor peak).
// TS = true ("A through tensor memory"): A is a plain fp32 activation matrix.  TMA brings ONE raw tile
// per stage; four producer warps (thread = row = TMEM lane) read their row, make the tf32 hi/lo pair
// in registers and tcgen05.st it into a TMEM staging ring; the MMAs take A from TMEM and only B (the
// pre-split weights) from shared memory.  Shared-memory traffic per K block drops from 160 KB to
// 112 KB, the activations' L2/HBM traffic halves, and every epilogue writes a single fp32 array.
template <int BN, int EPI, bool TS>
__global__ void __launch_bounds__(TS ? TC_THREADS_TS : TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
               const TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = TS ? Cfg::STAGES_TS : Cfg::STAGES;
  constexpr int STAGE_BYTES = TS ? Cfg::STAGE_BYTES_TS : Cfg::STAGE_BYTES;
  constexpr int B_OFF = TS ? Cfg::A_BYTES : 2 * Cfg::A_BYTES;    // TS: one raw A tile, then B_hi, B_lo
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* ready_bar = bars + 2 * STAGES + 4;   // [STAGES] splitters -> MMA (TS)
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (int)(p.K / BK);
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_mn * p.ksplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 4); }
    for (int s = 0; s < 2; s++) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TS ? 512 : Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int split = t / tiles_mn;
        const int r = t % tiles_mn;
        const int mb = r % p.m_tiles, nb = r / p.m_tiles;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          const bool skip_bl = TS && (p.probe & 8);      // diagnostic: do not load B_lo (the MMAs then read stale smem)
          mbar_expect_tx(&full_bar[stage], skip_bl ? STAGE_BYTES - Cfg::B_BYTES : STAGE_BYTES);
          const int kc = (int)(split * p.K) + kb * BK;
          tma_load_2d(&mapAh, &full_bar[stage], sa, kc, mb * BM);
          if (!TS) tma_load_2d(&mapAl, &full_bar[stage], sa + Cfg::A_BYTES, kc, mb * BM);
          tma_load_2d(&mapBh, &full_bar[stage], sa + B_OFF, kc, nb * BN);
          if (!skip_bl) tma_load_2d(&mapBl, &full_bar[stage], sa + B_OFF + Cfg::B_BYTES, kc, nb * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      if (lane == 0) mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait((TS && !(p.probe & 1)) ? &ready_bar[stage] : &full_bar[stage], phase);   // TS: the producers waited for the TMA
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t dAh = make_sdesc(sa);
          const uint64_t dAl = make_sdesc(sa + Cfg::A_BYTES);
          const uint64_t dBh = make_sdesc(sa + B_OFF);
          const uint64_t dBl = make_sdesc(sa + B_OFF + Cfg::B_BYTES);
          const uint32_t ta = tmem_base + Cfg::A_TMEM_COL + stage * 64;   // TS: hi at +0, lo at +32
#pragma unroll
          for (int k = 0; k < BK / 8; k++) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);  // 8 tf32 = 32 bytes along K
            if (TS && p.probe) {
              if (!(p.probe & 4)) tc_mma_tf32_ts(d_tmem, ta + 32 + k * 8, dBh + ko, idesc, 1);
              if (!(p.probe & 2)) tc_mma_tf32_ts(d_tmem, ta + k * 8, dBl + ko, idesc, 1);
              tc_mma_tf32_ts(d_tmem, ta + k * 8, dBh + ko, idesc, 1);
            } else if (TS) {
              tc_mma_tf32_ts(d_tmem, ta + 32 + k * 8, dBh + ko, idesc, (kb | k) != 0);
              tc_mma_tf32_ts(d_tmem, ta + k * 8, dBl + ko, idesc, 1);
              tc_mma_tf32_ts(d_tmem, ta + k * 8, dBh + ko, idesc, 1);
            } else {
              tc_mma_tf32(d_tmem, dAl + ko, dBh + ko, idesc, (kb | k) != 0);
              tc_mma_tf32(d_tmem, dAh + ko, dBl + ko, idesc, 1);
              tc_mma_tf32(d_tmem, dAh + ko, dBh + ko, idesc, 1);
            }
          }
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (TS && warp >= 6 && !(p.probe & 1)) {
    // ===================== A producers (warps 6..9): global fp32 -> tf32 hi/lo -> TMEM =====================
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int row = q * 32 + lane;               // A-tile row = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A_TMEM_COL;
    // The raw fp32 A tile arrives by TMA in the stage's A slot (SWIZZLE_128B: 16-byte chunk c of row r sits
    // at chunk c ^ (r & 7)); each thread reads its own row, splits it and stores the pair to TMEM.  The
    // staging slot of a stage is free whenever the stage's smem has been refilled (same MMA commit).
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; kb++) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* arow = smem + stage * STAGE_BYTES + row * 128;
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
          split_tf32(x.x, hi[4 * c], lo[4 * c]);
          split_tf32(x.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32(x.z, hi[4 * c + 2], lo[4 * c + 2]);
          split_tf32(x.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tc_fence_after();
        tc_st32(trow + stage * 64, hi);
        tc_st32(trow + stage * 64 + 32, lo);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int split = t / tiles_mn;
      const int r = t % tiles_mn;
      const int mb = r % p.m_tiles, nb = r / p.m_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int64_t m = (int64_t)mb * BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_STRIDE;
      tc_epilogue_row<BN, EPI, TS>(p, trow, m, (int64_t)nb * BN, split);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TS ? 512 : Cfg::TMEM_COLS));
  }
}


// ------------------------------------------------------------------------------ two row blocks per CTA (critic layers)
// Measured on the B200 (benchmarks/gemm_probe.py, profiles/r2_gemm_probe.jsonl): the 128x128 A-through-TMEM kernel above
// runs at ~1600 cycles per 32-column K block whether it issues 3, 2 or 1 MMA per 8-column step -- it is paced by
// operand delivery (48 KB of TMA boxes per K block and SM, ~30 B/clk/SM), not by the tensor pipe (768 cycles) and not by
// the A staging chain.  This variant halves the weight traffic per unit of work: a CTA owns TWO 128-row blocks of the same
// 128-column panel, so one [B_hi | B_lo] stage (32 KB) feeds 24 MMAs instead of 12 -- 64 KB of operands per 2 units
// instead of 96 KB.  TMEM: two accumulators (columns 0 / 128) + a 2-slot staging ring of {A0 hi, A0 lo, A1 hi, A1 lo}
// (columns 256 + 128 * slot).  The accumulators are single-buffered; the exposed epilogue (two warp quads, one per
// accumulator) is ~4 % of a tile.  Warps: 0 TMA, 1 MMA, 2-5 epilogue of block 0, 6-9 A producers, 10-13 epilogue of block 1.
constexpr int TS2_THREADS = 448;
constexpr int TS2_STAGES = 3;
constexpr int TS2_SLOTS = 2;
constexpr int TS2_STORE_BYTES = 8 * 4096;   // one 32x32 fp32 staging box per epilogue warp (TMA store)
template <int BN>
struct Ts2Cfg {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = TS2_STAGES * STAGE_BYTES + TS2_STORE_BYTES + 1024 + 256;
  static constexpr int ACC_STRIDE = 128;
  static constexpr int RING_COL = 256;
  static constexpr int SLOT_COLS = 128;
};

// role timers of the diagnostic build path (p.prof != nullptr): cycles spent inside the named wait
#define JRR_TIMED_WAIT(slot, stmt)                                   \
  do {                                                               \
    if (p.prof) {                                                    \
      const long long _t0 = clock64();                               \
      stmt;                                                          \
      tacc[slot] += clock64() - _t0;                                 \
    } else {                                                         \
      stmt;                                                          \
    }                                                                \
  } while (0)

template <int BN, int EPI>
__global__ void __launch_bounds__(TS2_THREADS, 1)
gemm_ts2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                const __grid_constant__ CUtensorMap mapBl, const __grid_constant__ CUtensorMap mapO, const TcParams p) {
  using Cfg = Ts2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* store_smem = smem + TS2_STAGES * Cfg::STAGE_BYTES;        // [8 epilogue warps][4 KB], 1024-byte aligned
  uint64_t* bars = (uint64_t*)(store_smem + TS2_STORE_BYTES);
  uint64_t* full_bar = bars;                               // [STAGES]  TMA -> producers, MMA
  uint64_t* empty_bar = bars + TS2_STAGES;                 // [STAGES]  MMA commit -> TMA
  uint64_t* ready_bar = bars + 2 * TS2_STAGES;             // [SLOTS]   producers -> MMA
  uint64_t* afree_bar = bars + 2 * TS2_STAGES + TS2_SLOTS; // [SLOTS]   MMA commit -> producers
  uint64_t* tfull_bar = bars + 2 * TS2_STAGES + 2 * TS2_SLOTS;   // [1]  MMA commit -> epilogue
  uint64_t* tempty_bar = tfull_bar + 1;                    // [1]  epilogue (8 warps) -> MMA
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long tacc[6] = {0, 0, 0, 0, 0, 0};
  const long long t_start = clock64();
  const int num_kb = (int)(p.K / BK);
  const int tiles_mn = p.m_tiles * p.n_tiles;              // m_tiles counts 256-row tiles here
  const int num_tiles = tiles_mn * p.ksplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO) : "memory");
    for (int s = 0; s < TS2_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < TS2_SLOTS; s++) { mbar_init(&ready_bar[s], 4); mbar_init(&afree_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int split = t / tiles_mn;
        const int r = t % tiles_mn;
        const int mb = r % p.m_tiles, nb = r / p.m_tiles;
        const int row0 = mb * 2 * BM;
        // an odd number of 128-row blocks: the last tile's second block re-reads the first (its rows are never stored)
        const int row1 = (row0 + BM < (int)p.M) ? row0 + BM : row0;
        for (int kb = 0; kb < num_kb; kb++) {
          JRR_TIMED_WAIT(0, mbar_wait(&empty_bar[stage], phase ^ 1));
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          const bool skip_a = (p.probe & 4) != 0;       // diagnostic: no A tiles (the producers then stage zeros)
          mbar_expect_tx(&full_bar[stage], skip_a ? 2 * Cfg::B_BYTES : Cfg::STAGE_BYTES);
          const int kc = (int)(split * p.K) + kb * BK;
          if (!skip_a) {
            tma_load_2d(&mapA, &full_bar[stage], sa, kc, row0);
            tma_load_2d(&mapA, &full_bar[stage], sa + Cfg::A_BYTES, kc, row1);
          }
          tma_load_2d(&mapBh, &full_bar[stage], sa + 2 * Cfg::A_BYTES, kc, nb * BN);
          tma_load_2d(&mapBl, &full_bar[stage], sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, kc, nb * BN);
          if (++stage == TS2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 16 + 0] = tacc[0]; p.prof[blockIdx.x * 16 + 1] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    // Measured with the role timers (benchmarks/gemm_prof.py): tcgen05.mma issue blocks the lane at the pipe's pace
    // (~64 cycles per 128x128x8 MMA: the queue is shallow), so every cycle this lane spends between two K blocks --
    // barrier probes, fences, descriptor arithmetic -- is a cycle the tensor pipe drains and idles.  Hence: the next K
    // block's "staged" barrier is probed while this block's last MMAs are still queued, the stage's own TMA barrier is not
    // waited for again (the A producers waited for it before they arrived on `ready`), and the other 31 lanes stay out
    // of the loop.
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
    if (lane == 0) {
      int stage = 0, slot = 0;
      uint32_t sphase = 0, acc_phase = 0;
      bool staged = false;              // ready_bar[slot] of the K block about to be issued is known to have completed
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        JRR_TIMED_WAIT(0, mbar_wait(tempty_bar, acc_phase ^ 1));
        for (int kb = 0; kb < num_kb; kb++) {
          if (!staged) JRR_TIMED_WAIT(2, mbar_wait(&ready_bar[slot], sphase));
          tc_fence_after();
          const long long t_issue0 = p.prof ? clock64() : 0;
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t dBh = make_sdesc(sa + 2 * Cfg::A_BYTES);
          const uint64_t dBl = make_sdesc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
          const uint32_t ring = tmem_base + Cfg::RING_COL + slot * Cfg::SLOT_COLS;
          const int nslot = slot + 1 == TS2_SLOTS ? 0 : slot + 1;
          const uint32_t nphase = slot + 1 == TS2_SLOTS ? sphase ^ 1 : sphase;
          const int nrep = (p.probe & 8) ? 2 : ((p.probe & 16) ? 0 : 1);    // diagnostic: every MMA twice / no MMAs at all
          for (int rep = 0; rep < nrep; rep++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const uint32_t d_tmem = tmem_base + i * Cfg::ACC_STRIDE;
              const uint32_t ta = ring + i * 64;           // hi at +0, lo at +32
#pragma unroll
              for (int k = 0; k < BK / 8; k++) {
                const uint64_t ko = (uint64_t)(k * 32 >> 4);
                tc_mma_tf32_ts(d_tmem, ta + 32 + k * 8, dBh + ko, idesc, (kb | k) != 0);
                tc_mma_tf32_ts(d_tmem, ta + k * 8, dBl + ko, idesc, 1);
                tc_mma_tf32_ts(d_tmem, ta + k * 8, dBh + ko, idesc, 1);
              }
            }
          }
          if (p.prof) tacc[4] += clock64() - t_issue0;
          // ONE commit per K block (it releases the smem stage to the TMA lane and, two K blocks later, the staging slot
          // to the A producers), issued before the probe below so the release is not delayed
          if (p.prof && (p.probe & 64)) {
            JRR_TIMED_WAIT(5, tc_commit(&empty_bar[stage]));
          } else {
            tc_commit(&empty_bar[stage]);
          }
          if (kb == num_kb - 1) tc_commit(tfull_bar);
          // probe the next K block's barrier while the MMAs above are still executing
          if (p.prof && (p.probe & 128)) {
            JRR_TIMED_WAIT(3, staged = mbar_try(&ready_bar[nslot], nphase));
          } else {
            staged = mbar_try(&ready_bar[nslot], nphase);
          }
          if (++stage == TS2_STAGES) stage = 0;
          slot = nslot;
          sphase = nphase;
        }
        acc_phase ^= 1;
      }
      if (p.prof) {
        p.prof[blockIdx.x * 16 + 2] = tacc[0]; p.prof[blockIdx.x * 16 + 3] = 0; p.prof[blockIdx.x * 16 + 4] = tacc[2];
        p.prof[blockIdx.x * 16 + 5] = clock64() - t_start;
        p.prof[blockIdx.x * 16 + 11] = tacc[3]; p.prof[blockIdx.x * 16 + 12] = tacc[4]; p.prof[blockIdx.x * 16 + 13] = tacc[5];
      }
    }
  } else if (warp >= 6 && warp < 10) {
    // ===================== A producers: smem fp32 rows -> tf32 hi/lo -> TMEM staging ring =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::RING_COL;
    int stage = 0, slot = 0;
    uint32_t phase = 0;
    int n = 0;                        // running K-block count of this CTA
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; kb++, n++) {
        JRR_TIMED_WAIT(0, mbar_wait(&full_bar[stage], phase));
        // the MMAs that read this staging slot belong to K block n - 2; their completion is completion (n-2)/3 of the
        // empty barrier of stage (n-2) % 3 (it cannot run a phase ahead: its next completion needs K block n + 1 staged)
        if (n >= TS2_SLOTS)
          JRR_TIMED_WAIT(1, mbar_wait(&empty_bar[(n - TS2_SLOTS) % TS2_STAGES], (uint32_t)(((n - TS2_SLOTS) / TS2_STAGES) & 1)));
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const uint8_t* arow = smem + stage * Cfg::STAGE_BYTES + i * Cfg::A_BYTES + row * 128;
          float hi[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float4 x = make_float4(1.f, 2.f, 3.f, 4.f);
            if (!(p.probe & 6)) x = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));   // (diagnostic: no smem read)
            split_tf32(x.x, hi[4 * c], lo[4 * c]);
            split_tf32(x.y, hi[4 * c + 1], lo[4 * c + 1]);
            split_tf32(x.z, hi[4 * c + 2], lo[4 * c + 2]);
            split_tf32(x.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
          tc_st32(trow + slot * Cfg::SLOT_COLS + i * 64, hi);
          tc_st32(trow + slot * Cfg::SLOT_COLS + i * 64 + 32, lo);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[slot]);
        if (++stage == TS2_STAGES) { stage = 0; phase ^= 1; }
        if (++slot == TS2_SLOTS) slot = 0;
      }
    }
    if (p.prof && warp == 6 && lane == 0) {
      p.prof[blockIdx.x * 16 + 6] = tacc[0]; p.prof[blockIdx.x * 16 + 7] = tacc[1]; p.prof[blockIdx.x * 16 + 8] = clock64() - t_start;
    }
  } else if (warp >= 2) {
    // ===================== epilogue: warps 2-5 row block 0, warps 10-13 row block 1 =====================
    const int q = warp & 3;
    const int blk = warp >= 10 ? 1 : 0;
    float* my_stage = p.probe & 32 ? nullptr : reinterpret_cast<float*>(store_smem + (blk * 4 + q) * 4096);   // (probe 32: direct stores)
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int split = t / tiles_mn;
      const int r = t % tiles_mn;
      const int mb = r % p.m_tiles, nb = r / p.m_tiles;
      JRR_TIMED_WAIT(0, mbar_wait(tfull_bar, acc_phase));
      tc_fence_after();
      // rows past M (the duplicated block of an odd last tile, or padding) fall out through the m < p.M guards
      const int64_t m = (int64_t)mb * 2 * BM + blk * BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + blk * Cfg::ACC_STRIDE;
      tc_epilogue_row<BN, EPI, true>(p, trow, m, (int64_t)nb * BN, split, my_stage, &mapO);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();       // every bulk store of this lane has landed before the CTA retires
    if (p.prof && warp == 2 && lane == 0) { p.prof[blockIdx.x * 16 + 9] = tacc[0]; p.prof[blockIdx.x * 16 + 10] = clock64() - t_start; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ------------------------------------------------------------------------------ CTA pairs (critic layers)
// The two-row-block kernel above still reads 96 KB of weights from shared memory per K block next to 64 KB of TMA
// writes and 32 KB of producer reads: 192 KB per 1536 MMA cycles = the whole 128 B/clk of the SM's shared memory, so the
// tensor pipe waits on it (~59 % busy in steady state).  Here two CTAs of a cluster (the two SMs of a TPC) run ONE
// tcgen05.mma.cta_group::2 of 256 rows x BN columns per K step: every CTA stages its own 128 activation rows in its own
// tensor memory and holds only HALF of the weight panel (BN/2 rows) in shared memory -- the tensor cores of the pair
// share the halves -- so per K block and SM the shared memory sees 48 KB of TMA writes + 48 KB of MMA reads + 16 KB of
// producer reads = 112 KB per 1536 cycles, and L2 delivers 48 KB instead of 64 KB.
//   rank 0 = leader (issues every MMA, owns the `ready` / `tempty` barriers both CTAs arrive on), rank 1 = peer.
//   warps (both CTAs): 0 TMA (own A rows, own half of B_hi / B_lo -> own smem, own `full` barrier), 6-9 A producers
//   (wait for the own `full`, split, tcgen05.st into the own staging ring, arrive on the LEADER's `ready`), 2-5 / 10-13
//   epilogue of the own 128 rows (column halves), 1 = MMA lane (leader only).  Completion of the MMAs is multicast to the
//   `empty` (smem stage + staging slot free) and `tfull` (accumulator complete) barriers of both CTAs.
//   TMEM per CTA: accumulator 128 lanes x BN columns at column 0, staging ring {hi 32 | lo 32} x 4 slots at column 256.
constexpr int PAIR_THREADS = 448;
// Footprint: the pair kernel leaves room on its SMs for one CTA of the step's small kernels (pose chain, critic_pre / post:
// 256 threads x <= 96 registers, <= 17 KB of shared memory), which are latency-bound and would otherwise queue for the 20 SMs
// the 64 pairs of a 4096-frame GEMM do not use: 88 registers per thread (448 x 88 = 39 424 of 65 536; the epilogues spill
// 4-12 bytes) and <= 194 KB of shared memory (three operand stages at 256 columns).
constexpr int PAIR_MAXREG = 88;
constexpr int PAIR_SMEM_BUDGET = 194 * 1024;
constexpr int PAIR_STORE_BYTES = 8 * 4096;
template <int BN, bool ABITS = false>
struct PairCfg {
  static constexpr int HB = BN / 2;                 // weight rows (output columns) held by one CTA
  static constexpr int A_BYTES = ABITS ? 0 : BM * BK * 4;     // (ABITS: the producers expand mask bits, no A tile in smem)
  static constexpr int B_BYTES = HB * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (4 * STAGE_BYTES + PAIR_STORE_BYTES + 1024 + 256 <= PAIR_SMEM_BUDGET) ? 4 : 3;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PAIR_STORE_BYTES + 1024 + 256;
  static constexpr int RING_COL = 256;
  static constexpr int E0 = ((HB + 31) / 32) * 32;  // accumulator columns drained by epilogue warps 2-5 ...
  static constexpr int E1 = BN - E0;                // ... and by warps 10-13
  static_assert(BN % 32 == 0 && BN <= 256 && B_BYTES % 1024 == 0, "pair tile width");
};

// ABITS: the A operand is a 0/1 matrix given as bits (TcParams::a_bits; the layer-2 backward of the critic reads the ReLU mask
// against diag(w3) W2): exact in tf32, so every K step is two MMAs (A . B_hi + A . B_lo) instead of three, the A tile does
// not travel through TMA at all -- a producer thread loads ONE word per K block and row -- and the layer that produces the
// mask writes 512 KB of bits instead of 16.8 MB of fp32 values.
template <int BN, int EPI, bool ABITS = false>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(PAIR_MAXREG)
gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                 const __grid_constant__ CUtensorMap mapBl, const __grid_constant__ CUtensorMap mapO, const TcParams p) {
  using Cfg = PairCfg<BN, ABITS>;
  extern __shared__ uint8_t smem_raw[];
  // (the dynamic shared memory window starts at the same offset in both CTAs, so the aligned layouts coincide)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* store_smem = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(store_smem + PAIR_STORE_BYTES);
  uint64_t* full_bar = bars;                          // [STAGES] own TMA -> own A producers
  uint64_t* empty_bar = bars + Cfg::STAGES;           // [STAGES] MMA commit (multicast) -> own TMA lane
  uint64_t* ready_bar = bars + 2 * Cfg::STAGES;       // [STAGES] A producers of both CTAs -> MMA lane (leader's copy)
  uint64_t* tfull_bar = bars + 3 * Cfg::STAGES;       // [1] MMA commit (multicast) -> own epilogue
  uint64_t* tempty_bar = tfull_bar + 1;               // [1] epilogue warps of both CTAs -> MMA lane (leader's copy)
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_entry = clock64();       // (role time stamps of the diagnostic path, p.prof != nullptr: cycles since entry)
#define JRR_STAMP(slot) do { if (p.prof) p.prof[blockIdx.x * 16 + (slot)] = clock64() - t_entry; } while (0)
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = (int)(p.K / BK);
  const int tiles_mn = p.m_tiles * p.n_tiles;         // m_tiles counts 256-row pair tiles
  const int num_tiles = tiles_mn * p.ksplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO) : "memory");
    for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 8); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised and both allocations are done before anything remote
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) JRR_STAMP(1);        // prologue done

  if (warp == 0) {
    // ===================== TMA producer (each CTA feeds its own shared memory) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int split = t / tiles_mn;
        const int r = t % tiles_mn;
        const int mb = r % p.m_tiles, nb = r / p.m_tiles;
        int rowA = mb * 2 * BM + (int)rank * BM;
        if (rowA >= (int)p.M) rowA -= BM;      // odd number of 128-row blocks: the peer re-reads the leader's (never stored)
        const int rowB = nb * BN + (int)rank * Cfg::HB;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int kc = (int)(split * p.K) + kb * BK;
          if (!ABITS) tma_load_2d(&mapA, &full_bar[stage], sa, kc, rowA);
          tma_load_2d(&mapBh, &full_bar[stage], sa + Cfg::A_BYTES, kc, rowB);
          tma_load_2d(&mapBl, &full_bar[stage], sa + Cfg::A_BYTES + Cfg::B_BYTES, kc, rowB);
          if (t == pair && kb == 0) JRR_STAMP(2);      // first stage issued
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // the last commits still arrive on this CTA's `empty` barriers: wait for them before the CTA may retire
      for (int s = 0; s < Cfg::STAGES; s++) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane of the leader CTA) =====================
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)((2 * BM) >> 4) << 24);
    if (rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool staged = false;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(tempty_bar, acc_phase ^ 1);
        for (int kb = 0; kb < num_kb; kb++) {
          if (!staged) mbar_wait(&ready_bar[stage], phase);
          tc_fence_after();
          if (t == pair && kb == 0) JRR_STAMP(4);      // first K block staged in both CTAs
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t dBh = make_sdesc(sa + Cfg::A_BYTES);
          const uint64_t dBl = make_sdesc(sa + Cfg::A_BYTES + Cfg::B_BYTES);
          const uint32_t ta = tmem_base + Cfg::RING_COL + stage * 64;     // hi at +0, lo at +32
#pragma unroll
          for (int k = 0; k < BK / 8; k++) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            if (!ABITS) tc_mma_tf32_ts_pair(tmem_base, ta + 32 + k * 8, dBh + ko, idesc, (kb | k) != 0);
            tc_mma_tf32_ts_pair(tmem_base, ta + k * 8, dBl + ko, idesc, ABITS ? (uint32_t)((kb | k) != 0) : 1u);
            tc_mma_tf32_ts_pair(tmem_base, ta + k * 8, dBh + ko, idesc, 1);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (kb == num_kb - 1) { tc_commit_pair(tfull_bar); if (t == pair) JRR_STAMP(5); }   // all MMAs of the first tile issued
          const int nstage = stage + 1 == Cfg::STAGES ? 0 : stage + 1;
          const uint32_t nphase = stage + 1 == Cfg::STAGES ? phase ^ 1 : phase;
          staged = mbar_try(&ready_bar[nstage], nphase);     // probed while the MMAs above are still queued
          stage = nstage;
          phase = nphase;
        }
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 6 && warp < 10) {
    // ===================== A producers: own smem fp32 rows -> tf32 hi/lo -> own TMEM staging ring =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::RING_COL;
    const uint32_t ready_leader = map_to_cta(ready_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      if (ABITS) {
        // 0/1 operand: one word of mask bits per K block and row, expanded to 32 exact tf32 values (no lo half)
        const int split = t / tiles_mn, mb = (t % tiles_mn) % p.m_tiles;
        const int64_t m = (int64_t)mb * 2 * BM + (int64_t)rank * BM + row;
        const uint32_t* wrow = p.a_bits + (m < p.M ? m : 0) * (int64_t)(p.K * p.ksplit / BK) + (int64_t)split * num_kb;
        uint32_t word = m < p.M ? wrow[0] : 0u;
        for (int kb = 0; kb < num_kb; kb++) {
          const uint32_t next = (m < p.M && kb + 1 < num_kb) ? wrow[kb + 1] : 0u;     // in flight across the wait below
          mbar_wait(&full_bar[stage], phase);      // the weights landed; the staging slot's previous readers are done
          if (t == pair && kb == 0 && warp == 6 && lane == 0) JRR_STAMP(3);
          float hi[32];
#pragma unroll
          for (int i = 0; i < 32; i++) hi[i] = ((word >> i) & 1u) ? 1.f : 0.f;
          tc_fence_after();
          tc_st32(trow + stage * 64, hi);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ready_leader + stage * 8);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          word = next;
        }
        continue;
      }
      for (int kb = 0; kb < num_kb; kb++) {
        // `full` of this stage follows the `empty` commit of the MMAs that read staging slot `stage` four K blocks ago
        mbar_wait(&full_bar[stage], phase);
        if (t == pair && kb == 0 && warp == 6 && lane == 0) JRR_STAMP(3);    // first operands landed
        const uint8_t* arow = smem + stage * Cfg::STAGE_BYTES + row * 128;
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
          split_tf32(x.x, hi[4 * c], lo[4 * c]);
          split_tf32(x.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32(x.z, hi[4 * c + 2], lo[4 * c + 2]);
          split_tf32(x.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tc_fence_after();
        tc_st32(trow + stage * 64, hi);
        tc_st32(trow + stage * 64 + 32, lo);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_leader + stage * 8);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue: warps 2-5 columns [0, E0), warps 10-13 columns [E0, BN) of the own rows ==========
    const int q = warp & 3;
    const int half = warp >= 10 ? 1 : 0;
    float* my_stage = reinterpret_cast<float*>(store_smem + (half * 4 + q) * 4096);
    const uint32_t tempty_leader = map_to_cta(tempty_bar, 0);
    uint32_t acc_phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int split = t / tiles_mn;
      const int r = t % tiles_mn;
      const int mb = r % p.m_tiles, nb = r / p.m_tiles;
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      if (t == pair && warp == 2 && lane == 0) JRR_STAMP(6);                 // accumulator complete
      const int64_t m = (int64_t)mb * 2 * BM + (int64_t)rank * BM + q * 32 + lane;   // rows past M fall out through the guards
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + half * Cfg::E0;
      if (half == 0) tc_epilogue_row<Cfg::E0, EPI, true>(p, trow, m, (int64_t)nb * BN, split, my_stage, &mapO);
      else tc_epilogue_row<Cfg::E1, EPI, true>(p, trow, m, (int64_t)nb * BN + Cfg::E0, split, my_stage, &mapO);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader);
      if (t == pair && warp == 2 && lane == 0) JRR_STAMP(7);                 // epilogue of the first tile issued
      acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
    if (warp == 2 && lane == 0) JRR_STAMP(8);                                // stores landed
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's shared / tensor memory stays alive until the leader's last MMA has used it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
  if (threadIdx.x == 32) JRR_STAMP(9);       // teardown done
#undef JRR_STAMP
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Encoded maps are cached: the operands of a step are the same device buffers call after call (model constants, workspace
// slices), and cuTensorMapEncodeTiled is host time that a small-batch SMPL forward otherwise pays four or five times.
struct MapKey {
  const float* base; int64_t rows, cols, ld; int box_rows;
  bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
static constexpr int MAP_CACHE = 64;
static thread_local MapKey g_map_keys[MAP_CACHE];
static thread_local CUtensorMap g_map_vals[MAP_CACHE];
static thread_local int g_map_count = 0, g_map_next = 0;

int make_tensor_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  const MapKey key{base, rows, cols, ld, box_rows};
  for (int i = 0; i < g_map_count; i++)
    if (g_map_keys[i] == key) { *map = g_map_vals[i]; return JRR_OK; }
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(JRR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(JRR_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  const int slot = g_map_count < MAP_CACHE ? g_map_count++ : (g_map_next++ % MAP_CACHE);
  g_map_keys[slot] = key;
  g_map_vals[slot] = *map;
  return JRR_OK;
}

int device_num_sms(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) device = 0;
  if (!cached[device]) cudaDeviceGetAttribute(&cached[device], cudaDevAttrMultiProcessorCount, device);
  return cached[device];
}

template <int BN, int EPI, bool TS = false>
static int launch_tc(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  CUtensorMap mAh, mAl, mBh, mBl;
  const int64_t Ktot = g.K * g.ksplit;
  if (int rc = make_tensor_map_2d(&mBh, g.B_hi, g.N, Ktot, g.ldb, BN)) return rc;
  if (int rc = make_tensor_map_2d(&mBl, g.B_lo, g.N, Ktot, g.ldb, BN)) return rc;
  const int64_t Ka = g.k_valid > 0 ? g.k_valid : Ktot;      // columns of A that exist; TMA zero-fills beyond
  if (int rc = make_tensor_map_2d(&mAh, g.A_hi, g.M, Ka, g.lda, BM)) return rc;
  if (TS) {
    mAl = mAh;                 // unused by the kernel: one raw fp32 A tile per stage
  } else {
    if (int rc = make_tensor_map_2d(&mAl, g.A_lo, g.M, Ka, g.lda, BM)) return rc;
  }
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.ksplit = g.ksplit;
  p.m_tiles = (int)(g.M / BM);
  p.n_tiles = (int)((g.N + BN - 1) / BN);
  p.out0 = g.out0; p.out1 = g.out1; p.ldo = g.ldo; p.bias = g.bias; p.mask = g.mask; p.ldmask = g.ldmask;
  p.rowscale = g.rowscale; p.vec = g.vec; p.out2 = g.out2;
  p.A = g.A_hi; p.lda = g.lda;
  p.mask_bits = g.mask_bits; p.mask_bits_out = g.mask_bits_out;
  p.logit_part = g.logit_part; p.n_logit_part = g.n_logit_part; p.logit_bias = g.logit_bias;
  p.logit_gscale = g.logit_gscale; p.rows_valid = g.rows_valid;
  auto kern = gemm_tc_kernel<BN, EPI, TS>;
  constexpr int smem_bytes = TS ? Cfg::SMEM_BYTES_TS : Cfg::SMEM_BYTES;
  JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int tiles = p.m_tiles * p.n_tiles * p.ksplit;
  int grid_override = 0;
  int reps = 1;
  if (g.probe_env) {       // jrr_debug_gemm only: diagnostic knobs for benchmarks/gemm_probe.py
    const char* e = getenv("JRR_GEMM_PROBE");
    p.probe = e ? atoi(e) : 0;
    e = getenv("JRR_GEMM_PROBE_REPS");
    reps = e ? std::max(1, atoi(e)) : 1;
    e = getenv("JRR_GEMM_PROBE_GRID");
    if (e && atoi(e) > 0) grid_override = atoi(e);
  }
  const int grid = std::min(tiles, grid_override > 0 ? grid_override : device_num_sms(m->device));
  for (int r = 0; r < reps; r++) {
    kern<<<grid, TS ? TC_THREADS_TS : TC_THREADS, smem_bytes, st>>>(mAh, mAl, mBh, mBl, p);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}


// diagnostic (benchmarks/gemm_prof.py): role timers of every gemm_ts2_kernel launch go to consecutive 148x16 slots
static long long* g_prof_base = nullptr;
static int g_prof_slots = 0, g_prof_next = 0;

template <int BN, int EPI>
static int launch_ts2(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  using Cfg = Ts2Cfg<BN>;
  CUtensorMap mA, mBh, mBl;
  const int64_t Ktot = g.K * g.ksplit;
  if (int rc = make_tensor_map_2d(&mBh, g.B_hi, g.N, Ktot, g.ldb, BN)) return rc;
  if (int rc = make_tensor_map_2d(&mBl, g.B_lo, g.N, Ktot, g.ldb, BN)) return rc;
  const int64_t Ka = g.k_valid > 0 ? g.k_valid : Ktot;
  if (int rc = make_tensor_map_2d(&mA, g.A_hi, g.M, Ka, g.lda, BM)) return rc;
  CUtensorMap mO;                  // epilogue store boxes: 32 rows x 32 columns of out0 (split-K partials stacked along rows)
  if (int rc = make_tensor_map_2d(&mO, g.out0, g.M * g.ksplit, g.N, g.ldo, 32)) return rc;
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.ksplit = g.ksplit;
  p.m_tiles = (int)((g.M + 2 * BM - 1) / (2 * BM));
  p.n_tiles = (int)((g.N + BN - 1) / BN);
  p.out0 = g.out0; p.out1 = g.out1; p.ldo = g.ldo; p.bias = g.bias; p.mask = g.mask; p.ldmask = g.ldmask;
  p.rowscale = g.rowscale; p.vec = g.vec; p.out2 = g.out2;
  p.A = g.A_hi; p.lda = g.lda;
  p.mask_bits = g.mask_bits; p.mask_bits_out = g.mask_bits_out;
  p.logit_part = g.logit_part; p.n_logit_part = g.n_logit_part; p.logit_bias = g.logit_bias;
  p.logit_gscale = g.logit_gscale; p.rows_valid = g.rows_valid;
  if (g.probe_env) {        // jrr_debug_gemm only: JRR_GEMM_PROF = device address (decimal) of 16 int64 counters per CTA
    const char* e = getenv("JRR_GEMM_PROF");
    p.prof = e ? (long long*)strtoull(e, nullptr, 10) : nullptr;
    e = getenv("JRR_GEMM_PROBE");
    p.probe = e ? atoi(e) : 0;
  }
  if (g_prof_base != nullptr && g_prof_slots > 0) {
    p.prof = g_prof_base + (int64_t)(g_prof_next % g_prof_slots) * 148 * 16;
    g_prof_next++;
  }
  auto kern = gemm_ts2_kernel<BN, EPI>;
  JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int tiles = p.m_tiles * p.n_tiles * p.ksplit;
  int reps = 1;
  if (g.probe_env) {
    const char* e = getenv("JRR_GEMM_PROBE_REPS");
    reps = e ? std::max(1, atoi(e)) : 1;
  }
  const int grid = std::min(tiles, device_num_sms(m->device));
  for (int r = 0; r < reps; r++) {
    kern<<<grid, TS2_THREADS, Cfg::SMEM_BYTES, st>>>(mA, mBh, mBl, mO, p);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

// CTA-pair kernel: grid = 2 x min(pair tiles, SM pairs); the static cluster dimension keeps a pair on one TPC
template <int BN, int EPI, bool ABITS = false>
static int launch_pair(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  using Cfg = PairCfg<BN, ABITS>;
  CUtensorMap mA, mBh, mBl, mO;
  const int64_t Ktot = g.K * g.ksplit;
  if (int rc = make_tensor_map_2d(&mBh, g.B_hi, g.N, Ktot, g.ldb, Cfg::HB)) return rc;
  if (int rc = make_tensor_map_2d(&mBl, g.B_lo, g.N, Ktot, g.ldb, Cfg::HB)) return rc;
  const int64_t Ka = g.k_valid > 0 ? g.k_valid : Ktot;
  if (ABITS) mA = mBh;       // (unused: the A operand comes from a_bits)
  else if (int rc = make_tensor_map_2d(&mA, g.A_hi, g.M, Ka, g.lda, BM)) return rc;
  if (int rc = make_tensor_map_2d(&mO, g.out0, g.M * g.ksplit, g.N, g.ldo, 32)) return rc;
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.ksplit = g.ksplit;
  p.m_tiles = (int)((g.M + 2 * BM - 1) / (2 * BM));
  p.n_tiles = (int)((g.N + BN - 1) / BN);      // a last partial tile: TMA zero-fills the missing weight rows, stores are clipped
  p.out0 = g.out0; p.out1 = g.out1; p.ldo = g.ldo; p.bias = g.bias; p.mask = g.mask; p.ldmask = g.ldmask;
  p.rowscale = g.rowscale; p.vec = g.vec; p.out2 = g.out2;
  p.A = g.A_hi; p.lda = g.lda;
  p.mask_bits = g.mask_bits; p.mask_bits_out = g.mask_bits_out;
  p.logit_part = g.logit_part; p.n_logit_part = g.n_logit_part; p.logit_bias = g.logit_bias;
  p.logit_gscale = g.logit_gscale; p.rows_valid = g.rows_valid;
  if (EPI == EPI_BIAS_RELU_HEAD && (Cfg::E0 != 128 || Cfg::E1 != 128))
    return fail(JRR_ERR_INVALID, "tc gemm (CTA pairs): the fused head needs 128-column epilogue parts");
  p.a_bits = g.a_bits;
  auto kern = gemm_pair_kernel<BN, EPI, ABITS>;
  JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int tiles = p.m_tiles * p.n_tiles * p.ksplit;
  int reps = 1;
  if (g.probe_env) {
    const char* e = getenv("JRR_GEMM_PROBE_REPS");
    reps = e ? std::max(1, atoi(e)) : 1;
  }
  if (g_prof_base != nullptr && g_prof_slots > 0) {      // diagnostic (benchmarks/gemm_prof.py): 16 stamps per CTA and launch
    p.prof = g_prof_base + (int64_t)(g_prof_next % g_prof_slots) * 148 * 16;
    g_prof_next++;
  }
  const int pairs = std::min(tiles, device_num_sms(m->device) / 2);
  for (int r = 0; r < reps; r++) {
    kern<<<2 * pairs, PAIR_THREADS, Cfg::SMEM_BYTES, st>>>(mA, mBh, mBl, mO, p);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

template <int BN>
static int launch_pair_epi(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  switch (g.epi) {
    case EPI_BIAS_RELU_SPLIT: return launch_pair<BN, EPI_BIAS_RELU_SPLIT>(m, g, st);
    case EPI_MASK_SPLIT: return launch_pair<BN, EPI_MASK_SPLIT>(m, g, st);
    case EPI_BIAS_RELU_HEAD: return launch_pair<BN, EPI_BIAS_RELU_HEAD>(m, g, st);
    case EPI_STORE_SPLITK: return launch_pair<BN, EPI_STORE_SPLITK>(m, g, st);
    default: return fail(JRR_ERR_INVALID, "tc gemm (CTA pairs): unsupported epilogue");
  }
}

}  // namespace jrr
extern "C" int jrr_debug_set_gemm_prof(long long* base, int slots) {
  jrr::g_prof_base = base; jrr::g_prof_slots = slots; jrr::g_prof_next = 0;
  return JRR_OK;
}
namespace jrr {

// JRR_GEMM_PAIR: 0 = never, 1 (default) = CTA pairs for the A-through-TMEM GEMMs with M >= 256 and N % 256 == 0
// (N = 768 as 192-column pair tiles: 64 tiles of 0.75 units instead of 48 of 1), 2 = 256-column tiles for N = 768 too
static int pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JRR_GEMM_PAIR");
    v = e ? atoi(e) : 1;
  }
  return v;
}

bool gemm_pair_bits_available(const JrrModel* m, int64_t M) {
  static const bool on = [] { const char* e = getenv("JRR_CRITIC_BITS"); return !(e && e[0] == '0'); }();
  return on && m->gemm_impl == 0 && pair_mode() > 0 && M >= 2 * BM;
}

static bool use_ts2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JRR_GEMM_TS2");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_gemm_tc(const JrrModel* m, const GemmDesc& g, cudaStream_t st) {
  if (g.M % BM != 0 || g.K % BK != 0) return fail(JRR_ERR_INVALID, "tc gemm: M%128 or K%32");
  if (((uintptr_t)g.A_hi | (uintptr_t)(g.a_via_tmem ? nullptr : g.A_lo) | (uintptr_t)g.B_hi | (uintptr_t)g.B_lo) & 15)
    return fail(JRR_ERR_INVALID, "tc gemm: operands must be 16-byte aligned");
  if (g.a_via_tmem) {      // plain fp32 A (A_hi) through tensor memory, pre-split B
    if (g.lda % 4 != 0) return fail(JRR_ERR_INVALID, "tc gemm (A through TMEM): lda % 4");
    // the two small GEMMs of the folded loss path (CTA pairs only)
    if (g.epi == EPI_STORE_T && g.M >= 2 * BM) {
      static const int bn = [] { const char* e = getenv("JRR_FOLD_BN"); return e ? atoi(e) : 256; }();
      if (bn == 128) return launch_pair<128, EPI_STORE_T>(m, g, st);
      if (bn == 192) return launch_pair<192, EPI_STORE_T>(m, g, st);
      return launch_pair<256, EPI_STORE_T>(m, g, st);
    }
    if (g.epi == EPI_STORE_SPLITK && g.N == 224 && g.M >= 2 * BM) return launch_pair<224, EPI_STORE_SPLITK>(m, g, st);
    if (g.a_bits != nullptr) {
      if (!gemm_pair_bits_available(m, g.M) || g.epi != EPI_MASK_SPLIT || g.N % 256 != 0 || g.ksplit != 1)
        return fail(JRR_ERR_INVALID, "tc gemm: a 0/1 A operand needs the CTA-pair kernel (mask epilogue, N % 256 == 0)");
      return launch_pair<256, EPI_MASK_SPLIT, true>(m, g, st);
    }
    if (g.N % 128 != 0) return fail(JRR_ERR_INVALID, "tc gemm (A through TMEM): N % 128");
    if (g.M >= 2 * BM && pair_mode() > 0 && !(g.probe_env && getenv("JRR_GEMM_PROBE_TS1"))) {
      if (g.N == 768 && pair_mode() == 1) return launch_pair_epi<192>(m, g, st);
      if (g.N % 256 == 0) return launch_pair_epi<256>(m, g, st);
    }
    if (g.M >= 2 * BM && use_ts2() && !(g.probe_env && getenv("JRR_GEMM_PROBE_TS1"))) {
      // two 128-row blocks per CTA: half the weight traffic per MMA (see gemm_ts2_kernel)
      switch (g.epi) {
        case EPI_BIAS_RELU_SPLIT: return launch_ts2<128, EPI_BIAS_RELU_SPLIT>(m, g, st);
        case EPI_MASK_SPLIT: return launch_ts2<128, EPI_MASK_SPLIT>(m, g, st);
        case EPI_BIAS_RELU_HEAD: return launch_ts2<128, EPI_BIAS_RELU_HEAD>(m, g, st);
        case EPI_STORE_SPLITK:
          // (a 96-column panel buys nothing here: measured ~58 cycles per tcgen05.mma at N = 96 against ~60 at N = 128 --
          //  below 128 columns the instruction time does not shrink with N -- so N = 768 runs as 6 x 16 = 96 CTAs of
          //  256 x 128 and leaves the other SMs to the concurrent branch of the step)
          return launch_ts2<128, EPI_STORE_SPLITK>(m, g, st);
        default: break;
      }
    }
    switch (g.epi) {
      case EPI_BIAS_RELU_SPLIT: return launch_tc<128, EPI_BIAS_RELU_SPLIT, true>(m, g, st);
      case EPI_MASK_SPLIT: return launch_tc<128, EPI_MASK_SPLIT, true>(m, g, st);
      case EPI_BIAS_RELU_HEAD: return launch_tc<128, EPI_BIAS_RELU_HEAD, true>(m, g, st);
      case EPI_STORE_SPLITK:
        // N = 768 (the critic's dh): 96-column tiles give 256 tiles = 1.73 waves of 0.75 units instead of
        // 192 tiles = 1.3 waves of 1 unit on 148 SMs
        if (g.N == 768 && g.ksplit == 1) return launch_tc<96, EPI_STORE_SPLITK, true>(m, g, st);
        return launch_tc<128, EPI_STORE_SPLITK, true>(m, g, st);
      default: return fail(JRR_ERR_INVALID, "tc gemm (A through TMEM): unsupported epilogue");
    }
  }
  switch (g.epi) {
    case EPI_STORE_T:
      if (g.N % 128 == 0) return launch_tc<128, EPI_STORE_T>(m, g, st);
      break;
    case EPI_BIAS_RELU_SPLIT:
      if (g.N % 128 == 0) return launch_tc<128, EPI_BIAS_RELU_SPLIT>(m, g, st);
      break;
    case EPI_MASK_SPLIT:
      if (g.N % 128 == 0) return launch_tc<128, EPI_MASK_SPLIT>(m, g, st);
      break;
    case EPI_BIAS_RELU_HEAD:
      if (g.N % 128 == 0) return launch_tc<128, EPI_BIAS_RELU_HEAD>(m, g, st);
      break;
    case EPI_STORE_SPLITK:
      if (g.N == 224) return launch_tc<224, EPI_STORE_SPLITK>(m, g, st);
      if (g.N % 128 == 0) return launch_tc<128, EPI_STORE_SPLITK>(m, g, st);
      break;
  }
  return fail(JRR_ERR_INVALID, "tc gemm: unsupported shape/epilogue");
}

}  // namespace jrr
