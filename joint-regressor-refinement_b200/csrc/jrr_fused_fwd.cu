// Fused forward of the loss path: the augmented pose/shape blend GEMM (3xTF32, tcgen05 + TMA)
// whose EPILOGUE is linear blend skinning and the 17x6890 joint-regressor reduction.
//
//   feat[128 poses, 224] x Pt[64 vertices * 3, 224]^T  --tcgen05-->  TMEM accumulator
//   epilogue (8 warps, thread = pose = TMEM lane, two warps per lane quarter split the tile's
//   64 vertices): tcgen05.ld -> v = sum_k w_k (A_k [vp;1]) with the joint transforms cached in
//   registers -> 51 regressor accumulators in registers, carried across the consecutive vertex
//   tiles a CTA owns and written once per (pose block, CTA segment) as a per-CTA partial sum.
//
// Skinned vertices never exist in memory on this path.  The blended vertices vp are written
// once (pose-contiguous) because the backward pass needs them; the refit path stores the
// skinned vertices instead.  Tensor-pipe work of tile i+1 overlaps the SIMT epilogue of tile i
// through two TMEM accumulator stages.
//
// Replaces: smplx.lbs.lbs (blend_shapes, pose blend, W.A, T.v) reached through
// scripts/smpl.py:72-74 and the regressor contraction of utils.find_joints
// (scripts/utils.py:96-98) as used at scripts/optimize.py:228-229,306-307.
#include <algorithm>
#include <cstdlib>

#include "jrr_internal.cuh"
#include "jrr_tc.cuh"
#include "jrr_f32x2.cuh"

namespace jrr {

constexpr int FV = 64;                 // vertices per tile
constexpr int FBN = 3 * FV;            // 192 accumulator columns
constexpr int FBM = 128;
constexpr int FBK = 32;
constexpr int F_STAGES = 2;
constexpr int F_EPI_WARPS = 8;
constexpr int F_CTRL_WARPS = 4;                   // warp group 0: warp 0 TMA, warp 1 MMA, warps 2-3 idle
constexpr int F_THREADS = 32 * (F_CTRL_WARPS + F_EPI_WARPS);   // 384 threads launched at 168 registers; setmaxnreg then
constexpr int F_CTRL_REGS = 24;                   // shrinks the control group ...
constexpr int F_EPI_REGS = 240;                   // ... and grows the two epilogue warp groups
constexpr int F_A_BYTES = FBM * FBK * 4;          // 16 KB
constexpr int F_B_BYTES = FBN * FBK * 4;          // 24 KB
constexpr int F_STAGE_BYTES = 2 * F_A_BYTES + 2 * F_B_BYTES;   // 80 KB
constexpr int F_ACC_STRIDE = 256;
constexpr int F_TMEM_COLS = 512;
constexpr int F_REC_F4 = FV * REC_WORDS / 4;      // 448 float4 of vertex records per tile
constexpr int F_SMEM_BYTES = F_STAGES * F_STAGE_BYTES + 2 * F_REC_F4 * 16 + 1024 + 256;
// CTA pairs (PAIR, batches with an even number of 128-pose blocks): the single-CTA mainloop reads 120 KB of operands from
// shared memory per K block next to 80 KB of TMA writes -- 200 KB per ~1 680 MMA cycles is about all the shared memory can
// move, and the kernel ran at 0.6 of the tensor peak even with an empty epilogue (measured: 221 us for 110.8 GFLOP).  Two
// CTAs of a cluster run ONE tcgen05.mma.cta_group::2 of 256 poses x 64 vertices per K step: each CTA brings its own 128
// feature rows and only HALF of the vertex tile's blend-matrix rows (the pair's tensor cores share the halves), 56 KB of
// TMA writes + 84 KB of MMA reads per K block and SM, three stages instead of two.  The epilogue is unchanged: every CTA
// skins its own 128 poses out of its own tensor memory.
constexpr int FP_STAGES = 3;
constexpr int FP_B_BYTES = F_B_BYTES / 2;                          // 12 KB: 96 of the tile's 192 blend-matrix rows
constexpr int FP_STAGE_BYTES = 2 * F_A_BYTES + 2 * FP_B_BYTES;     // 56 KB
constexpr int FP_SMEM_BYTES = FP_STAGES * FP_STAGE_BYTES + 2 * F_REC_F4 * 16 + 1024 + 256;

// tile t = mb * n_tiles + nb; CTA c owns [ceil(c*T/G), ceil((c+1)*T/G))
__host__ __device__ inline int fused_tile_begin(int c, int T, int G) { return (int)(((int64_t)c * T + G - 1) / G); }
__host__ __device__ inline int fused_cta_of_tile(int t, int T, int G) { return (int)(((int64_t)t * G) / T); }

enum { FSTORE_NONE = 0, FSTORE_VP = 1, FSTORE_V = 2, FSTORE_V_ADD = 3 /* later skinning passes: added to the stored vertices */ };

// MODE (module path, which walks EVERY packed vertex and has no use for the regressor sums): F_FULL = skinning + 17x6890
// reduction (loss path); F_SKIN = skinning only (SMPL.forward: the skinned vertices are the output); F_BLEND = neither
// (the recomputation inside SMPL.backward needs the blended vertices only: the epilogue is a plain store).
enum { F_FULL = 0, F_SKIN = 1, F_BLEND = 2 };

template <int STORE, int MODE, bool PAIR>
__global__ void __launch_bounds__(F_THREADS, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                 const VtxRec* __restrict__ vrec, const float* __restrict__ AT, int64_t BP, int m_tiles,
                 int n_tiles, float* __restrict__ vT_out, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still provably shared
  constexpr int STAGES = PAIR ? FP_STAGES : F_STAGES;
  constexpr int STAGE_BYTES = PAIR ? FP_STAGE_BYTES : F_STAGE_BYTES;
  constexpr int B_BYTES = PAIR ? FP_B_BYTES : F_B_BYTES;
  float4* srec = (float4*)(smem + STAGES * STAGE_BYTES);                // [2][448]
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES + 2 * F_REC_F4 * 16);
  uint64_t* full_bar = bars;                      // [STAGES] own TMA -> MMA lane (PAIR: -> own relay lane)
  uint64_t* empty_bar = bars + 3;                 // [STAGES] MMA commit (PAIR: multicast) -> own TMA lane
  uint64_t* ready_bar = bars + 6;                 // [STAGES] PAIR: relay lanes of both CTAs -> MMA lane (leader's copy)
  uint64_t* tfull_bar = bars + 9;                 // [2] MMA commit (PAIR: multicast) -> own epilogue
  uint64_t* tempty_bar = bars + 11;               // [2] epilogue warps (PAIR: of both CTAs, leader's copy) -> MMA lane
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int num_kb = KA / FBK;  // 7
  // a tile = (row of pose blocks, vertex tile nb): one 128-pose block per row, or the pair's two (rank picks its own)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int grp = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int G = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int T = (PAIR ? m_tiles / 2 : m_tiles) * n_tiles;
  const int t_begin = fused_tile_begin(grp, T, G);
  const int t_end = fused_tile_begin(grp + 1, T, G);
  auto pose_block = [&](int t) { return PAIR ? 2 * (t / n_tiles) + (int)rank : t / n_tiles; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&ready_bar[s], 2); }
    for (int s = 0; s < 2; s++) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], PAIR ? 2 * F_EPI_WARPS : F_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(F_TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(F_TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // both CTAs' barriers are initialised and both allocations are done before anything remote
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < F_CTRL_WARPS) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F_CTRL_REGS));
   if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; t++) {
        const int mb = pose_block(t), nb = t % n_tiles;
        const int rowB = nb * FBN + (PAIR ? (int)rank * (FBN / 2) : 0);     // PAIR: this CTA's half of the tile's rows
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&mapAh, &full_bar[stage], sa, kb * FBK, mb * FBM);
          tma_load_2d(&mapAl, &full_bar[stage], sa + F_A_BYTES, kb * FBK, mb * FBM);
          tma_load_2d(&mapBh, &full_bar[stage], sa + 2 * F_A_BYTES, kb * FBK, rowB);
          tma_load_2d(&mapBl, &full_bar[stage], sa + 2 * F_A_BYTES + B_BYTES, kb * FBK, rowB);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (PAIR) {
        // the pair's last commits still arrive on this CTA's `empty` barriers: wait for them before the CTA may retire
        for (int s = 0; s < STAGES; s++) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // (PAIR: one lane of the leader CTA issues for both; M = 256 = the pair's two pose blocks)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FBN >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * FBM : FBM) >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (!PAIR || rank == 0) {
    for (int t = t_begin; t < t_end; t++) {
      if (lane == 0) mbar_wait_backoff(&tempty_bar[acc], acc_phase ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * F_ACC_STRIDE;
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait_backoff(PAIR ? &ready_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t dAh = make_sdesc(sa);
          const uint64_t dAl = make_sdesc(sa + F_A_BYTES);
          const uint64_t dBh = make_sdesc(sa + 2 * F_A_BYTES);
          const uint64_t dBl = make_sdesc(sa + 2 * F_A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < FBK / 8; k++) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            if (PAIR) {
              tc_mma_tf32_ss_pair(d_tmem, dAl + ko, dBh + ko, idesc, (kb | k) != 0);
              tc_mma_tf32_ss_pair(d_tmem, dAh + ko, dBl + ko, idesc, 1);
              tc_mma_tf32_ss_pair(d_tmem, dAh + ko, dBh + ko, idesc, 1);
            } else {
              tc_mma_tf32(d_tmem, dAl + ko, dBh + ko, idesc, (kb | k) != 0);
              tc_mma_tf32(d_tmem, dAh + ko, dBl + ko, idesc, 1);
              tc_mma_tf32(d_tmem, dAh + ko, dBh + ko, idesc, 1);
            }
          }
          if (PAIR) {
            tc_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) tc_commit_pair(&tfull_bar[acc]);
          } else {
            tc_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
   } else if (PAIR && warp == 2) {
    // ===================== relay: this CTA's operands landed -> the leader's `ready` barrier =====================
    // (same visibility chain as the CTA-pair GEMM of jrr_gemm_tc.cu: TMA completion observed on the own `full` barrier,
    // then a remote arrive; the MMA lane reads both CTAs' shared memory through the async proxy)
    if (lane == 0) {
      const uint32_t ready_leader = map_to_cta(ready_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; t++)
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&full_bar[stage], phase);
          mbar_arrive_cluster(ready_leader + stage * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
   }
  } else {
    // ===================== epilogue: skinning + regressor partial sums =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F_EPI_REGS));
    const int ew = warp - F_CTRL_WARPS; // 0..7
    const int q = warp & 3;             // TMEM lane quarter of this warp
    const int h = ew >> 2;              // which half of the tile's 64 vertices
    const int etid = threadIdx.x - 32 * F_CTRL_WARPS;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    // joint transforms of the 4 cached slots: rows 0/1 packed per column, row 2 scalar
    f32x2 A01[4][4];
    float A2[4][4];
    // regressor accumulators: sum2[c][p] = (joint 2p, joint 2p+1) of coordinate c
    f32x2 sum2[3][9];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int p = 0; p < 9; p++) sum2[c][p] = pk2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int c = 0; c < 4; c++) { A01[k][c] = pk2(0.f, 0.f); A2[k][c] = 0.f; }

    int cached_joint[4] = {-1, -1, -1, -1};   // joints whose transforms A01/A2 hold (for the current pose block)
    const uint32_t tempty_leader = PAIR ? map_to_cta(tempty_bar, 0) : 0u;
    if (t_begin < t_end) {
      const float4* g = reinterpret_cast<const float4*>(vrec + (t_begin % n_tiles) * FV);
      for (int e = etid; e < F_REC_F4; e += 32 * F_EPI_WARPS) srec[e] = __ldg(g + e);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int t = t_begin; t < t_end; t++) {
      const int mb = pose_block(t), nb = t % n_tiles;
      const int row_t = t / n_tiles;            // tile row: the pose block, or the pair's two
      const int buf = (t - t_begin) & 1;
      const float4* rec = srec + buf * F_REC_F4;
      const bool has_next = t + 1 < t_end;
      float4 pf0 = make_float4(0, 0, 0, 0), pf1 = pf0;
      if (has_next) {
        const float4* g = reinterpret_cast<const float4*>(vrec + ((t + 1) % n_tiles) * FV);
        pf0 = __ldg(g + etid);
        if (etid + 256 < F_REC_F4) pf1 = __ldg(g + etid + 256);
      }
      const int64_t b = (int64_t)mb * FBM + q * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * F_ACC_STRIDE + h * (FBN / 2);
      float nxt[12];
      tc_ld12_issue(trow, nxt);
      tc_wait_ld12(nxt);
      // rows of the pose-contiguous vertex store this warp walks (3 rows per vertex)
      float* vout = (STORE != FSTORE_NONE) ? vT_out + (int64_t)(3 * (nb * FV + h * 32)) * BP + b : nullptr;
      const float4* rh = rec + (h * 32) * 7;     // records of my 32 vertices
#pragma unroll 1
      for (int g = 0; g < 8; g++, rh += 4 * 7) {
        float vp[12];
#pragma unroll
        for (int e = 0; e < 12; e++) vp[e] = nxt[e];
        if (g + 1 < 8) tc_ld12_issue(trow + (g + 1) * 12, nxt);   // lands while this group is skinned
        float4 r0[4];
        float w3[4];
        uint32_t meta[4], many = 0;
#pragma unroll
        for (int ii = 0; ii < 4; ii++) {
          r0[ii] = rh[ii * 7];
          w3[ii] = rh[ii * 7 + 1].x;
          meta[ii] = __float_as_uint(r0[ii].x);
          many |= meta[ii];
        }
        if (g == 0) {
          // my first vertex of this tile: the record's reload bits refer to a vertex another warp
          // handled, so compare the slot joints with what this thread still caches from its last
          // vertex (32 packed vertices earlier, same pose block) and fetch only what changed
          uint32_t need = 0;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (cached_joint[k] != (int)((meta[0] >> (5 * k)) & 31u)) need |= 1u << (20 + k);
          meta[0] = (meta[0] & ~(0xFu << 20)) | need;
          many = meta[0] | meta[1] | meta[2] | meta[3];
        }
        const bool any_reload = MODE != F_BLEND && ((many >> 20) & 0xFu);
        const bool any_col = MODE == F_FULL && ((many >> 24) & 1u);
        float v[4][3];
        // ---- skinning.  Fast path: no slot changes inside the group -> branch-free, the four
        // vertices interleave in the instruction stream.
#define JRR_SKIN_VERTEX(ii)                                                                         \
        {                                                                                           \
          const float x = vp[ii * 3 + 0], y = vp[ii * 3 + 1], z = vp[ii * 3 + 2];                   \
          const f32x2 xx = pk2(x, x), yy = pk2(y, y), zz = pk2(z, z);                               \
          const float wk[4] = {r0[ii].y, r0[ii].z, r0[ii].w, w3[ii]};                               \
          f32x2 v01 = pk2(0.f, 0.f);                                                                \
          float v2 = 0.f;                                                                           \
          _Pragma("unroll") for (int k = 0; k < 4; k++) {                                           \
            const f32x2 y01 = fma2(A01[k][2], zz, fma2(A01[k][1], yy, fma2(A01[k][0], xx, A01[k][3]))); \
            const float y2 = fmaf(A2[k][2], z, fmaf(A2[k][1], y, fmaf(A2[k][0], x, A2[k][3])));     \
            const f32x2 ww = pk2(wk[k], wk[k]);                                                     \
            v01 = (k == 0) ? mul2(ww, y01) : fma2(ww, y01, v01);                                    \
            v2 = (k == 0) ? wk[k] * y2 : fmaf(wk[k], y2, v2);                                       \
          }                                                                                         \
          upk2(v01, v[ii][0], v[ii][1]);                                                            \
          v[ii][2] = v2;                                                                            \
        }
        if (MODE == F_BLEND) {
          // (nothing to skin)
        } else if (!any_reload) {
#pragma unroll
          for (int ii = 0; ii < 4; ii++) JRR_SKIN_VERTEX(ii)
        } else {
#pragma unroll
          for (int ii = 0; ii < 4; ii++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if ((meta[ii] >> (20 + k)) & 1u) {
                const int j = (meta[ii] >> (5 * k)) & 31u;
                const float* src = AT + (int64_t)(j * 12) * BP + b;
                float a[12];
#pragma unroll
                for (int e = 0; e < 12; e++) { a[e] = *src; src += BP; }
#pragma unroll
                for (int c = 0; c < 4; c++) { A01[k][c] = pk2(a[c], a[4 + c]); A2[k][c] = a[8 + c]; }
              }
            }
            JRR_SKIN_VERTEX(ii)
          }
        }
#undef JRR_SKIN_VERTEX
        if (STORE == FSTORE_VP) {
#pragma unroll
          for (int e = 0; e < 12; e++) { *vout = vp[e]; vout += BP; }
        } else if (STORE == FSTORE_V) {
#pragma unroll
          for (int ii = 0; ii < 4; ii++)
#pragma unroll
            for (int r = 0; r < 3; r++) { *vout = v[ii][r]; vout += BP; }
        } else if (STORE == FSTORE_V_ADD) {
          float old[12];
#pragma unroll
          for (int e = 0; e < 12; e++) old[e] = vout[(int64_t)e * BP];
#pragma unroll
          for (int ii = 0; ii < 4; ii++)
#pragma unroll
            for (int r = 0; r < 3; r++) { *vout = old[ii * 3 + r] + v[ii][r]; vout += BP; }
        }
        // ---- 17x6890 regressor reduction (zero columns contribute zeros; skipped per group)
        if (any_col) {
#pragma unroll
          for (int ii = 0; ii < 4; ii++) {
            const f32x2 vb[3] = {pk2(v[ii][0], v[ii][0]), pk2(v[ii][1], v[ii][1]), pk2(v[ii][2], v[ii][2])};
#pragma unroll
            for (int qq = 0; qq < JH_STRIDE / 4; qq++) {
              const float4 tt = rh[ii * 7 + 2 + qq];
              const f32x2 ja = pk2(tt.x, tt.y), jb = pk2(tt.z, tt.w);
#pragma unroll
              for (int c = 0; c < 3; c++) {
                if (2 * qq < 9) sum2[c][2 * qq] = fma2(ja, vb[c], sum2[c][2 * qq]);
                if (2 * qq + 1 < 9) sum2[c][2 * qq + 1] = fma2(jb, vb[c], sum2[c][2 * qq + 1]);
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cached_joint[k] = (int)((meta[3] >> (5 * k)) & 31u);
        if (g + 1 < 8) tc_wait_ld12(nxt);
      }
      // TMEM stage drained -> the MMA warp may start tile t+2 in it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tempty_leader + acc * 8);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // per-CTA partial sums leave when the pose block changes (or the CTA runs out of tiles)
      if (MODE != F_FULL) {
        if (!has_next || (t + 1) / n_tiles != row_t) {
#pragma unroll
          for (int k = 0; k < 4; k++) cached_joint[k] = -1;
        }
      } else if (!has_next || (t + 1) / n_tiles != row_t) {
#pragma unroll
        for (int k = 0; k < 4; k++) cached_joint[k] = -1;   // next tile belongs to other poses
        const int seg = grp - fused_cta_of_tile(row_t * n_tiles, T, G);
        float* dst = part + ((int64_t)(seg * 2 + h) * NACC) * BP + b;
#pragma unroll
        for (int p = 0; p < 9; p++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float lo, hi;
            upk2(sum2[c][p], lo, hi);
            dst[(int64_t)((2 * p) * 3 + c) * BP] = lo;
            if (2 * p + 1 < NH) dst[(int64_t)((2 * p + 1) * 3 + c) * BP] = hi;
            sum2[c][p] = pk2(0.f, 0.f);
          }
      }
      if (has_next) {
        float4* d = srec + (buf ^ 1) * F_REC_F4;
        d[etid] = pf0;
        if (etid + 256 < F_REC_F4) d[etid + 256] = pf1;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer's shared / tensor memory stays alive until the leader's last MMA has used it
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(F_TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(F_TMEM_COLS));
  }
}

// How the tiles of a launch are dealt out (host; the loss-seed kernel mirrors it to find a pose block's partial sums):
// rows of `mdiv` pose blocks (2 = CTA pairs) x n_tiles vertex tiles, contiguous ranges over G CTAs / CTA pairs.
static bool fused_pair_enabled() {
  static const bool on = [] { const char* e = getenv("JRR_FUSED_PAIR"); return !(e && e[0] == '0'); }();
  return on;
}
FusedSched fused_fwd_sched(const JrrModel* m, int64_t BP, int nv) {
  FusedSched s;
  const int m_tiles = (int)(BP / FBM);
  s.mdiv = (fused_pair_enabled() && m_tiles % 2 == 0) ? 2 : 1;
  s.n_tiles = nv / FV;
  s.T = (m_tiles / s.mdiv) * s.n_tiles;
  s.G = std::min(s.T, m->num_sms / s.mdiv);
  return s;
}

// number of partial-sum slots a pose block can receive for this batch (host; sizes `part`)
int fused_fwd_slots(const JrrModel* m, int64_t BP) {
  // worst case over every possible active-vertex prefix (multiples of 192 vertices)
  int worst = 1;
  for (int ns = 1; ns <= VP / 192; ns++) {
    const FusedSched s = fused_fwd_sched(m, BP, ns * 192);
    const int rows = s.T / s.n_tiles;
    for (int r = 0; r < rows; r++) {
      const int c0 = fused_cta_of_tile(r * s.n_tiles, s.T, s.G), c1 = fused_cta_of_tile((r + 1) * s.n_tiles - 1, s.T, s.G);
      worst = std::max(worst, c1 - c0 + 1);
    }
  }
  return 2 * worst;
}

int launch_fused_fwd(const JrrModel* m, const Workspace& w, int store, float* vT_out, cudaStream_t st, bool all_vertices) {
  const FusedSched sc = fused_fwd_sched(m, w.BP, all_vertices ? VP : m->nv_act);
  const bool pair = sc.mdiv == 2;
  CUtensorMap mAh, mAl, mBh, mBl;
  if (int rc = make_tensor_map_2d(&mAh, w.feat_hi, w.BP, KA, KA, FBM)) return rc;
  if (int rc = make_tensor_map_2d(&mAl, w.feat_lo, w.BP, KA, KA, FBM)) return rc;
  if (int rc = make_tensor_map_2d(&mBh, m->Pt_hi, NP, KA, KA, pair ? FBN / 2 : FBN)) return rc;
  if (int rc = make_tensor_map_2d(&mBl, m->Pt_lo, NP, KA, KA, pair ? FBN / 2 : FBN)) return rc;
  // loss path: only the active vertex prefix; module path (all_vertices): every packed vertex
  const int m_tiles = (int)(w.BP / FBM), n_tiles = sc.n_tiles;
  float* part = w.part + (int64_t)m->cur_pass * w.part_stride;
  const VtxRec* vrec = m->vrec;
  const float* AT = w.AT;
  int64_t BP = w.BP;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sc.G * sc.mdiv));
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = pair ? FP_SMEM_BYTES : F_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pair ? 1 : 0;
#define JRR_FFP(S, MD, PR)                                                                          \
  do {                                                                                              \
    auto kern = fused_fwd_kernel<S, MD, PR>;                                                        \
    JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes)); \
    JRR_CUDA(cudaLaunchKernelEx(&cfg, kern, mAh, mAl, mBh, mBl, vrec, AT, BP, m_tiles, n_tiles, vT_out, part)); \
  } while (0)
#define JRR_FFM(S, MD) do { if (pair) JRR_FFP(S, MD, true); else JRR_FFP(S, MD, false); } while (0)
  if (all_vertices) {
    // module path: no regressor sums; the recomputation of SMPL.backward (blended vertices out) does not skin either
    if (store == FSTORE_VP) JRR_FFM(FSTORE_VP, F_BLEND);
    else if (store == FSTORE_V) JRR_FFM(FSTORE_V, F_SKIN);
    else if (store == FSTORE_V_ADD) JRR_FFM(FSTORE_V_ADD, F_SKIN);
    else return fail(JRR_ERR_INVALID, "fused forward over every vertex needs an output");
  } else if (store == FSTORE_NONE) JRR_FFM(FSTORE_NONE, F_FULL);
  else if (store == FSTORE_VP) JRR_FFM(FSTORE_VP, F_FULL);
  else if (store == FSTORE_V) JRR_FFM(FSTORE_V, F_FULL);
  else JRR_FFM(FSTORE_V_ADD, F_FULL);
#undef JRR_FFM
#undef JRR_FFP
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

}  // namespace jrr
