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

// tile t = mb * n_tiles + nb; CTA c owns [ceil(c*T/G), ceil((c+1)*T/G))
__host__ __device__ inline int fused_tile_begin(int c, int T, int G) { return (int)(((int64_t)c * T + G - 1) / G); }
__host__ __device__ inline int fused_cta_of_tile(int t, int T, int G) { return (int)(((int64_t)t * G) / T); }

enum { FSTORE_NONE = 0, FSTORE_VP = 1, FSTORE_V = 2, FSTORE_V_ADD = 3 /* later skinning passes: added to the stored vertices */ };

// MODE (module path, which walks EVERY packed vertex and has no use for the regressor sums): F_FULL = skinning + 17x6890
// reduction (loss path); F_SKIN = skinning only (SMPL.forward: the skinned vertices are the output); F_BLEND = neither
// (the recomputation inside SMPL.backward needs the blended vertices only: the epilogue is a plain store).
enum { F_FULL = 0, F_SKIN = 1, F_BLEND = 2 };

template <int STORE, int MODE>
__global__ void __launch_bounds__(F_THREADS, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                 const VtxRec* __restrict__ vrec, const float* __restrict__ AT, int64_t BP, int m_tiles,
                 int n_tiles, float* __restrict__ vT_out, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still provably shared
  float4* srec = (float4*)(smem + F_STAGES * F_STAGE_BYTES);            // [2][448]
  uint64_t* bars = (uint64_t*)(smem + F_STAGES * F_STAGE_BYTES + 2 * F_REC_F4 * 16);
  uint64_t* full_bar = bars;                      // [F_STAGES]
  uint64_t* empty_bar = bars + F_STAGES;          // [F_STAGES]
  uint64_t* tfull_bar = bars + 2 * F_STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * F_STAGES + 2; // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * F_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int num_kb = KA / FBK;  // 7
  const int T = m_tiles * n_tiles, G = gridDim.x;
  const int t_begin = fused_tile_begin(blockIdx.x, T, G);
  const int t_end = fused_tile_begin(blockIdx.x + 1, T, G);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    for (int s = 0; s < F_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], F_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(F_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
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
        const int mb = t / n_tiles, nb = t % n_tiles;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * F_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], F_STAGE_BYTES);
          tma_load_2d(&mapAh, &full_bar[stage], sa, kb * FBK, mb * FBM);
          tma_load_2d(&mapAl, &full_bar[stage], sa + F_A_BYTES, kb * FBK, mb * FBM);
          tma_load_2d(&mapBh, &full_bar[stage], sa + 2 * F_A_BYTES, kb * FBK, nb * FBN);
          tma_load_2d(&mapBl, &full_bar[stage], sa + 2 * F_A_BYTES + F_B_BYTES, kb * FBK, nb * FBN);
          if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FBN >> 3) << 17) |
                               ((uint32_t)(FBM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = t_begin; t < t_end; t++) {
      if (lane == 0) mbar_wait_backoff(&tempty_bar[acc], acc_phase ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * F_ACC_STRIDE;
      for (int kb = 0; kb < num_kb; kb++) {
        if (lane == 0) {
          mbar_wait_backoff(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * F_STAGE_BYTES);
          const uint64_t dAh = make_sdesc(sa);
          const uint64_t dAl = make_sdesc(sa + F_A_BYTES);
          const uint64_t dBh = make_sdesc(sa + 2 * F_A_BYTES);
          const uint64_t dBl = make_sdesc(sa + 2 * F_A_BYTES + F_B_BYTES);
#pragma unroll
          for (int k = 0; k < FBK / 8; k++) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            tc_mma_tf32(d_tmem, dAl + ko, dBh + ko, idesc, (kb | k) != 0);
            tc_mma_tf32(d_tmem, dAh + ko, dBl + ko, idesc, 1);
            tc_mma_tf32(d_tmem, dAh + ko, dBh + ko, idesc, 1);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
    if (t_begin < t_end) {
      const float4* g = reinterpret_cast<const float4*>(vrec + (t_begin % n_tiles) * FV);
      for (int e = etid; e < F_REC_F4; e += 32 * F_EPI_WARPS) srec[e] = __ldg(g + e);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int t = t_begin; t < t_end; t++) {
      const int mb = t / n_tiles, nb = t % n_tiles;
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
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // per-CTA partial sums leave when the pose block changes (or the CTA runs out of tiles)
      if (MODE != F_FULL) {
        if (!has_next || (t + 1) / n_tiles != mb) {
#pragma unroll
          for (int k = 0; k < 4; k++) cached_joint[k] = -1;
        }
      } else if (!has_next || (t + 1) / n_tiles != mb) {
#pragma unroll
        for (int k = 0; k < 4; k++) cached_joint[k] = -1;   // next tile belongs to other poses
        const int seg = blockIdx.x - fused_cta_of_tile(mb * n_tiles, T, G);
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
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(F_TMEM_COLS));
  }
}

// number of partial-sum slots a pose block can receive for this batch (host; sizes `part`)
int fused_fwd_slots(int64_t BP, int num_sms) {
  // worst case over every possible active-vertex prefix (1..NSPLIT_B ranges of VS_B vertices)
  const int m_tiles = (int)(BP / FBM);
  int worst = 1;
  for (int ns = 1; ns <= VP / 192; ns++) {          // every possible active prefix (multiples of 192 vertices)
    const int n_tiles = ns * 192 / FV;
    const int T = m_tiles * n_tiles, G = std::min(T, num_sms);
    for (int mb = 0; mb < m_tiles; mb++) {
      const int c0 = fused_cta_of_tile(mb * n_tiles, T, G), c1 = fused_cta_of_tile((mb + 1) * n_tiles - 1, T, G);
      worst = std::max(worst, c1 - c0 + 1);
    }
  }
  return 2 * worst;
}

int launch_fused_fwd(const JrrModel* m, const Workspace& w, int store, float* vT_out, cudaStream_t st, bool all_vertices) {
  CUtensorMap mAh, mAl, mBh, mBl;
  if (int rc = make_tensor_map_2d(&mAh, w.feat_hi, w.BP, KA, KA, FBM)) return rc;
  if (int rc = make_tensor_map_2d(&mAl, w.feat_lo, w.BP, KA, KA, FBM)) return rc;
  if (int rc = make_tensor_map_2d(&mBh, m->Pt_hi, NP, KA, KA, FBN)) return rc;
  if (int rc = make_tensor_map_2d(&mBl, m->Pt_lo, NP, KA, KA, FBN)) return rc;
  // loss path: only the active vertex prefix; module path (all_vertices): every packed vertex
  const int m_tiles = (int)(w.BP / FBM), n_tiles = (all_vertices ? VP : m->nv_act) / FV;
  const int T = m_tiles * n_tiles, G = std::min(T, m->num_sms);
#define JRR_FFM(S, MD)                                                                              \
  do {                                                                                              \
    auto kern = fused_fwd_kernel<S, MD>;                                                            \
    JRR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES)); \
    kern<<<G, F_THREADS, F_SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, m->vrec, w.AT, w.BP, m_tiles, n_tiles, \
                                             vT_out, w.part + (int64_t)m->cur_pass * w.part_stride);  \
  } while (0)
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
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

}  // namespace jrr
