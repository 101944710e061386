// Fused backward of the loss path: skinning backward (SIMT "generator" warps) feeding the
// transpose-side blend GEMM (tcgen05) through shared memory -- the pose-blend gradient
// dvp[B,20736] never exists in HBM.
//
//   generator (8 warps, thread = pose, 256 poses per CTA): for every vertex of the CTA's K range
//     dv = Jhat^T g, dvp = (sum_k w_k AR_k)^T dv, dA_k += w_k dv (x) [vp;1] (flushed per run)
//     and the 4-vertex group's 12 dvp values are tf32-split and stored as three 16-byte chunks
//     into the K-major SWIZZLE_128B A-operand tile of the current 32-column K block
//   TMA warp: P_hi / P_lo tiles [224 x 32] of the augmented blend matrix (B operand)
//   MMA warp: dfeat[2 x 128 poses, 224] += dvp . P^T, 3xTF32, accumulators in TMEM
//   epilogue (the generator warps): TMEM -> split-K partial dfeat[ks][b][224]
//
// Work item = (256-pose block, K split of 768 packed vertices = the backward skinning range, so
// the dA flush lists are the ones the unfused kernel uses).  A tiles are double-buffered, the P
// tile single-buffered (its reload hides behind the generator, which paces the pipeline).
//
// Replaces: autograd backward of smplx.lbs.lbs (skinning, pose/shape blend) and of the regressor
// contraction in utils.find_joints, as reached from scripts/optimize.py:264.
#include <algorithm>

#include "jrr_internal.cuh"
#include "jrr_tc.cuh"
#include "jrr_f32x2.cuh"

namespace jrr {

constexpr int GB_POSES = 256;                    // poses per CTA (two M=128 accumulators)
constexpr int GB_BK = 32;                        // K columns per block (one swizzle atom)
constexpr int GB_GEN_WARPS = 8;
constexpr int GB_CTRL_WARPS = 4;                 // warp group 0: warp 0 = TMA producer, warp 1 = MMA issuer, 2-3 idle
constexpr int GB_THREADS = 32 * (GB_CTRL_WARPS + GB_GEN_WARPS);   // 384 threads, launched at 168 registers each ...
constexpr int GB_CTRL_REGS = 24;                 // ... then the control group shrinks to 24
constexpr int GB_GEN_REGS = 240;                 // ... and the two generator warp groups grow to 240 (setmaxnreg)
constexpr int GB_A_TILE = 128 * GB_BK * 4;       // 16 KB: one [128 x 32] fp32 tile
constexpr int GB_A_STAGE = 4 * GB_A_TILE;        // {pose half 0,1} x {hi,lo} = 64 KB
constexpr int GB_P_TILE = KA * GB_BK * 4;        // 28 KB
constexpr int GB_P_BYTES = 2 * GB_P_TILE;        // hi + lo = 56 KB
constexpr int GB_REC_F4 = 32 * REC_WORDS / 4;    // vertex records per 32-vertex tile (224 float4)
constexpr int GB_SMEM = 2 * GB_A_STAGE + GB_P_BYTES + 2 * GB_REC_F4 * 16 + 1024 + 256;
constexpr int GB_TMEM_COLS = 512;                // accumulators at columns 0 and 256

// USE_DV = false: loss path, the vertex gradient is generated as dv = Jhat^T g (gT: the loss seed).
// USE_DV = true: module path (autograd backward of SMPL.forward), dv is READ from dvT [3*VP][BP] -- the caller's
// d loss / d vertices re-packed pose-contiguous, with the joints49 contributions already added (pack_dvertices_kernel).
template <bool USE_DV>
__global__ void __launch_bounds__(GB_THREADS, 1)
fused_bwd_kernel(const __grid_constant__ CUtensorMap mapPh, const __grid_constant__ CUtensorMap mapPl,
                 const VtxRec* __restrict__ vrec, const int* __restrict__ range_flush_base,
                 const float* __restrict__ AT, const float* __restrict__ vpT, const float* __restrict__ gT,
                 int64_t BP, int n_items, int nsplit, int vs /* vertices per K range: 768, 384 or 192 */,
                 float* __restrict__ dfeat, float* __restrict__ dAflush) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                      // [2 stages][64 KB]
  uint8_t* sP = smem + 2 * GB_A_STAGE;                     // [56 KB]
  float4* srec = (float4*)(sP + GB_P_BYTES);               // [2][224]
  uint64_t* bars = (uint64_t*)((uint8_t*)srec + 2 * GB_REC_F4 * 16);
  uint64_t* afull = bars;        // [2] generators -> MMA
  uint64_t* aempty = bars + 2;   // [2] MMA -> generators
  uint64_t* pfull = bars + 4;    // TMA -> MMA
  uint64_t* pempty = bars + 5;   // MMA -> TMA
  uint64_t* tfull = bars + 6;    // MMA -> epilogue
  uint64_t* tempty = bars + 7;   // epilogue -> MMA
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int GB_KB_PER_ITEM = vs * 3 / GB_BK;     // K blocks per item (72 / 36 / 18)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapPh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapPl) : "memory");
    for (int s = 0; s < 2; s++) { mbar_init(&afull[s], GB_GEN_WARPS); mbar_init(&aempty[s], 1); }
    mbar_init(pfull, 1);
    mbar_init(pempty, 1);
    mbar_init(tfull, 1);
    mbar_init(tempty, GB_GEN_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(GB_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < GB_CTRL_WARPS) {
    // ===================== control warp group: hands its registers to the generators =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GB_CTRL_REGS));
    if (warp == 0 && lane == 0) {
      // ---- TMA producer: P tiles (single buffer: refilled as soon as the MMAs that read it completed)
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ks = item % nsplit;
        for (int kb = 0; kb < GB_KB_PER_ITEM; kb++, n++) {
          mbar_wait_backoff(pempty, (n & 1) ^ 1);
          mbar_expect_tx(pfull, GB_P_BYTES);
          const int col = ks * (vs * 3) + kb * GB_BK;
          tma_load_2d(&mapPh, pfull, sP, col, 0);
          tma_load_2d(&mapPl, pfull, sP + GB_P_TILE, col, 0);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ---- MMA issuer
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KA >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
      uint32_t n = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        mbar_wait_backoff(tempty, (it & 1) ^ 1);           // accumulators drained by the previous item's epilogue
        tc_fence_after();
        for (int kb = 0; kb < GB_KB_PER_ITEM; kb++, n++) {
          const int s = n & 1;
          mbar_wait_backoff(&afull[s], (n >> 1) & 1);
          mbar_wait_backoff(pfull, n & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + s * GB_A_STAGE);
          const uint32_t pb = smem_u32(sP);
          const uint64_t dPh = make_sdesc(pb), dPl = make_sdesc(pb + GB_P_TILE);
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const uint64_t dAh = make_sdesc(a0 + hf * 2 * GB_A_TILE);
            const uint64_t dAl = make_sdesc(a0 + hf * 2 * GB_A_TILE + GB_A_TILE);
            const uint32_t d_tmem = tmem_base + hf * 256;
#pragma unroll
            for (int k = 0; k < GB_BK / 8; k++) {
              const uint64_t ko = (uint64_t)(k * 32 >> 4);
              tc_mma_tf32(d_tmem, dAl + ko, dPh + ko, idesc, (kb | k) != 0);
              tc_mma_tf32(d_tmem, dAh + ko, dPl + ko, idesc, 1);
              tc_mma_tf32(d_tmem, dAh + ko, dPh + ko, idesc, 1);
            }
          }
          tc_commit(&aempty[s]);
          tc_commit(pempty);
          if (kb == GB_KB_PER_ITEM - 1) tc_commit(tfull);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== generators (skinning backward) + epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GB_GEN_REGS));
    const int half = (warp - GB_CTRL_WARPS) >> 2;  // which 128-pose accumulator
    const int row = (warp & 3) * 32 + lane;        // TMEM lane == A-tile row of this thread
    const int gtid = threadIdx.x - 32 * GB_CTRL_WARPS;   // 0..255
    uint32_t n = 0, it = 0;                        // K-block / item counters

    for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
      const int ks = item % nsplit, mb2 = item / nsplit;
      const int64_t b = (int64_t)mb2 * GB_POSES + half * 128 + row;
      const int i0 = ks * vs;
      float* flush_dst = dAflush + (int64_t)range_flush_base[ks] * 12 * BP + b;

      f32x2 gp[USE_DV ? 1 : 3][USE_DV ? 1 : 9];
      if (!USE_DV) {
#pragma unroll
        for (int p = 0; p < 9; p++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const float lo = gT[(int64_t)((2 * p) * 3 + c) * BP + b];
            const float hi = (2 * p + 1 < NH) ? gT[(int64_t)((2 * p + 1) * 3 + c) * BP + b] : 0.f;
            gp[USE_DV ? 0 : c][USE_DV ? 0 : p] = pk2(lo, hi);
          }
      }
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
        if (gtid < GB_REC_F4) srec[gtid] = __ldg(gsrc + gtid);
      }
      const float* vsrc = vpT + (int64_t)(3 * i0) * BP + b;
      float nx[12];
#pragma unroll
      for (int q = 0; q < 12; q++) { nx[q] = *vsrc; vsrc += BP; }
      const float* dsrc = gT + (int64_t)(3 * i0) * BP + b;     // USE_DV: gT is dvT, walked like vpT
      float dnx[USE_DV ? 12 : 1];
      if (USE_DV) {
#pragma unroll
        for (int q = 0; q < 12; q++) { dnx[USE_DV ? q : 0] = *dsrc; dsrc += BP; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // (nothing of this item is queued yet)

      const int NT = vs / 32;         // record tiles per item (24 / 12 / 6)
#pragma unroll 1
      for (int t = 0; t < NT; t++) {
        const float4* rt = srec + (t & 1) * GB_REC_F4;
        const bool has_next = t + 1 < NT;
        float4 pf = make_float4(0, 0, 0, 0);
        if (has_next && gtid < GB_REC_F4)
          pf = __ldg(reinterpret_cast<const float4*>(vrec + i0 + (t + 1) * 32) + gtid);
#pragma unroll 1
        for (int sub = 0; sub < 8; sub++) {
          const int g = t * 8 + sub;               // 4-vertex group index inside the item (0..191)
          float cur[12];
#pragma unroll
          for (int q = 0; q < 12; q++) cur[q] = nx[q];
          float dcur[USE_DV ? 12 : 1];
          if (USE_DV) {
#pragma unroll
            for (int q = 0; q < 12; q++) dcur[USE_DV ? q : 0] = dnx[USE_DV ? q : 0];
          }
          if (g + 1 < NT * 8) {
#pragma unroll
            for (int q = 0; q < 12; q++) { nx[q] = *vsrc; vsrc += BP; }
            if (USE_DV) {
#pragma unroll
              for (int q = 0; q < 12; q++) { dnx[USE_DV ? q : 0] = *dsrc; dsrc += BP; }
            }
          }
          const float4* rh = rt + (sub * 4) * 7;
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
          const bool any_reload = (many >> 20) & 0xFu;
          float dv[4][3];
#pragma unroll
          for (int ii = 0; ii < 4; ii++) { dv[ii][0] = 0.f; dv[ii][1] = 0.f; dv[ii][2] = 0.f; }
          if (USE_DV) {
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
#pragma unroll
              for (int c = 0; c < 3; c++) dv[ii][c] = dcur[USE_DV ? ii * 3 + c : 0];
          } else if ((many >> 24) & 1u) {
#pragma unroll
            for (int ii = 0; ii < 4; ii++) {
              f32x2 a[3] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
              for (int qq = 0; qq < JH_STRIDE / 4; qq++) {
                const float4 tt = rh[ii * 7 + 2 + qq];
                const f32x2 ja = pk2(tt.x, tt.y), jb = pk2(tt.z, tt.w);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                  if (2 * qq < 9) a[c] = fma2(ja, gp[USE_DV ? 0 : c][USE_DV ? 0 : 2 * qq], a[c]);
                  if (2 * qq + 1 < 9) a[c] = fma2(jb, gp[USE_DV ? 0 : c][USE_DV ? 0 : 2 * qq + 1], a[c]);
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
          float o[12];
#define JRR_GB_VERTEX(ii)                                                                             \
          {                                                                                           \
            const float wk[4] = {r0[ii].y, r0[ii].z, r0[ii].w, w3[ii]};                               \
            const f32x2 d0 = pk2(dv[ii][0], dv[ii][0]), d1 = pk2(dv[ii][1], dv[ii][1]),               \
                        d2 = pk2(dv[ii][2], dv[ii][2]);                                               \
            const f32x2 vp01 = pk2(cur[ii * 3 + 0], cur[ii * 3 + 1]), vp21 = pk2(cur[ii * 3 + 2], 1.f); \
            f32x2 o01 = pk2(0.f, 0.f);                                                                \
            float o2 = 0.f;                                                                           \
            _Pragma("unroll") for (int k = 0; k < 4; k++) {                                           \
              const f32x2 u01 = fma2(AR01[k][2], d2, fma2(AR01[k][1], d1, mul2(AR01[k][0], d0)));     \
              const float u2 = fmaf(AR2[k][2], dv[ii][2], fmaf(AR2[k][1], dv[ii][1], AR2[k][0] * dv[ii][0])); \
              const f32x2 ww = pk2(wk[k], wk[k]);                                                     \
              o01 = fma2(ww, u01, o01);                                                               \
              o2 = fmaf(wk[k], u2, o2);                                                               \
              const f32x2 wv01 = mul2(ww, vp01), wv21 = mul2(ww, vp21);                               \
              dA01[k][0] = fma2(d0, wv01, dA01[k][0]); dA23[k][0] = fma2(d0, wv21, dA23[k][0]);       \
              dA01[k][1] = fma2(d1, wv01, dA01[k][1]); dA23[k][1] = fma2(d1, wv21, dA23[k][1]);       \
              dA01[k][2] = fma2(d2, wv01, dA01[k][2]); dA23[k][2] = fma2(d2, wv21, dA23[k][2]);       \
            }                                                                                         \
            upk2(o01, o[ii * 3 + 0], o[ii * 3 + 1]);                                                  \
            o[ii * 3 + 2] = o2;                                                                       \
          }
          if (!any_reload) {
#pragma unroll
            for (int ii = 0; ii < 4; ii++) JRR_GB_VERTEX(ii)
          } else {
#pragma unroll
            for (int ii = 0; ii < 4; ii++) {
              const bool first = (meta[ii] >> 25) & 1u;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                if ((meta[ii] >> (20 + k)) & 1u) {
                  if (!first) {
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
              JRR_GB_VERTEX(ii)
            }
          }
#undef JRR_GB_VERTEX
          // ---- the group's 12 dvp columns -> A-operand tiles (tf32 hi / lo), 3 swizzled 16-byte chunks
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const int col = g * 12 + j * 4;          // column inside the item's K range
            const int kbl = col >> 5;                // K block inside the item
            const uint32_t nk = n + kbl;             // global K-block number (ring position)
            const int s = nk & 1;
            if ((col & 31) == 0) {
              // first chunk this thread writes into K block kbl: its A stage must have been consumed
              mbar_wait(&aempty[s], ((nk >> 1) & 1) ^ 1);
            }
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; e++) split_tf32(o[j * 4 + e], hi[e], lo[e]);
            const int chunk = ((col & 31) >> 2) ^ (row & 7);
            uint8_t* dsth = sA + s * GB_A_STAGE + half * 2 * GB_A_TILE + row * 128 + chunk * 16;
            *reinterpret_cast<float4*>(dsth) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(dsth + GB_A_TILE) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            if ((col & 31) == 28) {
              // K block complete for this thread: publish to the async proxy, one arrive per warp
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(&afull[s]);
            }
          }
        }
        if (has_next && gtid < GB_REC_F4) srec[((t + 1) & 1) * GB_REC_F4 + gtid] = pf;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      n += GB_KB_PER_ITEM;
      // remaining dA slot contents leave as the range's last 4 flush events
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
      // ---- epilogue: split-K partial of the blend-feature gradient
      mbar_wait(tfull, it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + half * 256;
      float* out = dfeat + ((int64_t)ks * BP + b) * KA;
#pragma unroll 1
      for (int c = 0; c < KA / 32; c++) {
        float v[32];
        tc_ld32(trow + c * 32, v);
        float4* o4 = reinterpret_cast<float4*>(out + c * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(GB_TMEM_COLS));
  }
}

int launch_fused_bwd(const JrrModel* m, const Workspace& w, cudaStream_t st, const float* dvT) {
  if (w.BP % GB_POSES != 0) return fail(JRR_ERR_INVALID, "fused backward needs BP % 256 == 0");
  CUtensorMap mPh, mPl;
  if (int rc = make_tensor_map_2d(&mPh, m->P_hi, KA, NP, NP, KA)) return rc;
  if (int rc = make_tensor_map_2d(&mPl, m->P_lo, KA, NP, NP, KA)) return rc;
  // skinning pass p: its split-K partials follow those of the earlier passes, its dA flush events have their own region
  const bool small = dvT != nullptr && module_small_ranges(m, w.BP);
  const int ns_p = dvT != nullptr ? (small ? NSPLIT_S : NSPLIT_B) : m->nsplit_act;
  float* dfeat_p = w.dfeat + (int64_t)m->cur_pass * ns_p * w.BP * KA;
  float* flush_p = w.dAflush + (int64_t)m->flush_off[m->cur_pass] * 12 * w.BP;
  if (dvT != nullptr) {
    // module path: every packed vertex in 768-vertex ranges (the records / flush lists of the module backward)
    // (a batch of <= 1024 poses has at most four 256-pose blocks: 192-vertex ranges give it 36 items per block instead of 9)
    const int nsplit = small ? NSPLIT_S : NSPLIT_B;
    const int n_items = (int)(w.BP / GB_POSES) * nsplit;
    const int grid = std::min(n_items, m->num_sms);
    JRR_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
    fused_bwd_kernel<true><<<grid, GB_THREADS, GB_SMEM, st>>>(mPh, mPl, small ? m->vrec_s : m->vrec_b,
                                                              small ? m->range_flush_base_s : m->range_flush_base, w.AT, w.vpT, dvT,
                                                              w.BP, n_items, nsplit, small ? VS_S : VS_B, dfeat_p, flush_p);
    JRR_LAUNCH_CHECK();
    return JRR_OK;
  }
  const int nsplit = m->nsplit_act;      // K ranges of the active vertex prefix
  const int n_items = (int)(w.BP / GB_POSES) * nsplit;
  const int grid = std::min(n_items, m->num_sms);
  JRR_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
  fused_bwd_kernel<false><<<grid, GB_THREADS, GB_SMEM, st>>>(mPh, mPl, m->vrec_l, m->range_flush_base_l, w.AT, w.vpT, w.gT,
                                                             w.BP, n_items, nsplit, m->vs_l, dfeat_p, flush_p);
  JRR_LAUNCH_CHECK();
  return JRR_OK;
}

}  // namespace jrr
