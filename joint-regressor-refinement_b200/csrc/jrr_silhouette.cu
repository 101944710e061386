// Soft-silhouette term of the refinement loss (widening row 8f-4): the differentiable rasteriser the reference gets from
// pytorch3d==0.3.0 -- MeshRasterizer(blur_radius=0, faces_per_pixel=1) + SoftSilhouetteShader(sigma=1e-4) behind
// PerspectiveCameras(T=cam, focal_length=5000/image_size, principal_point=0) -- scripts/mesh_renderer.py:23-79, called
// through render_mesh (scripts/optimize.py:77-85: x and y flipped, vertices scaled by 2) and compared with the Mask R-CNN
// silhouette by MSELoss (optimize.py:234-236).  pytorch3d is not available offline: the rasteriser's rules are restated
// from its published CUDA sources (PARITY UNPINNED, like the 2-D projection of 8f-2) and pinned against
// oracle/silhouette_oracle.py.
//
//   project     thread per (frame, vertex): view = (-2x, -2y, 2z) + cam, ndc = (f X/Z, f Y/Z, Z)
//   raster      thread per (frame, face): the pixel centres of the face's bounding box; barycentric inside test, interpolated
//               depth >= 0, 64-bit atomicMin of (depth bits, face id) per pixel -- the z-buffer of faces_per_pixel = 1
//   shade       thread per (frame, pixel): winner face -> squared distance to its closest edge -> alpha = sigmoid(d2 / sigma)
//               (SoftSilhouetteShader: 1 - prod(1 - sigmoid(-dists / sigma)), dists = -d2 inside a face); optional MSE partials
//   backward    thread per (frame, face) re-walks its bounding box and sums the gradient of ITS pixels (no float atomics:
//               fixed order), thread per (frame, vertex) gathers its faces' corners (CSR) and chains through the projection,
//               a block per frame reduces the camera gradient in a fixed order.
#include "jrr_internal.cuh"

namespace jrr {

constexpr float SIL_EPS = 1e-8f;      // pytorch3d kEpsilon

struct SilVec2 { float x, y; };

__device__ __forceinline__ float sil_edge(float px, float py, float ax, float ay, float bx, float by) {
  // EdgeFunctionForward(p, v0, v1)
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

// squared distance of p to the segment (a, b) and the segment parameter (PointLineDistanceForward)
__device__ __forceinline__ float sil_seg_dist(float px, float py, float ax, float ay, float bx, float by, float* t_out) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= SIL_EPS) {
    *t_out = 1.f;
    return (px - bx) * (px - bx) + (py - by) * (py - by);
  }
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = fminf(fmaxf(t, 0.f), 1.f);
  *t_out = t;
  const float qx = ax + t * bax, qy = ay + t * bay;
  return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

// pixel index (along one axis, pytorch3d's flipped convention already applied by the caller) -> NDC
__device__ __forceinline__ float sil_pix_to_ndc(int i, int S) { return -1.f + (2.f * (float)i + 1.f) / (float)S; }

struct SilFace {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
  int lo_x, hi_x, lo_y, hi_y;      // pixel-index range (NDC-ordered indices xi, yi) that can contain covered centres
  bool live;
};

__device__ __forceinline__ SilFace sil_load_face(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int64_t f, int S) {
  SilFace t;
  const int i0 = faces[f * 3 + 0], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
  t.x0 = ndc[i0 * 3 + 0]; t.y0 = ndc[i0 * 3 + 1]; t.z0 = ndc[i0 * 3 + 2];
  t.x1 = ndc[i1 * 3 + 0]; t.y1 = ndc[i1 * 3 + 1]; t.z1 = ndc[i1 * 3 + 2];
  t.x2 = ndc[i2 * 3 + 0]; t.y2 = ndc[i2 * 3 + 1]; t.z2 = ndc[i2 * 3 + 2];
  const float zmax = fmaxf(t.z0, fmaxf(t.z1, t.z2));
  const float area = sil_edge(t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);     // EdgeFunctionForward(v0, v1, v2)
  const float xmin = fminf(t.x0, fminf(t.x1, t.x2)), xmax = fmaxf(t.x0, fmaxf(t.x1, t.x2));
  const float ymin = fminf(t.y0, fminf(t.y1, t.y2)), ymax = fmaxf(t.y0, fmaxf(t.y1, t.y2));
  const bool finite = isfinite(xmin) && isfinite(xmax) && isfinite(ymin) && isfinite(ymax) && isfinite(zmax);
  t.live = finite && !(zmax < 0.f) && !(area <= SIL_EPS && area >= -SIL_EPS);
  if (t.live) {
    // centre i lies at c_i = -1 + (2 i + 1) / S, i.e. i = (c + 1) S / 2 - 1/2: keep exactly the indices whose centre can fall
    // into [min, max] (1e-3 of a pixel of slack against round-off; the inside test decides).  SMPL faces are one to three
    // pixels across at 224 x 224: a one-pixel margin on every side made the box 16-36 centres for 1-4 covered ones, and the
    // two face kernels 9 / 4 ms per 1024 frames
    const float h = 0.5f * (float)S;
    t.lo_x = max(0, (int)ceilf((fmaxf(xmin, -2.f) + 1.f) * h - 0.5f - 1e-3f));
    t.hi_x = min(S - 1, (int)floorf((fminf(xmax, 2.f) + 1.f) * h - 0.5f + 1e-3f));
    t.lo_y = max(0, (int)ceilf((fmaxf(ymin, -2.f) + 1.f) * h - 0.5f - 1e-3f));
    t.hi_y = min(S - 1, (int)floorf((fminf(ymax, 2.f) + 1.f) * h - 0.5f + 1e-3f));
    if (t.lo_x > t.hi_x || t.lo_y > t.hi_y) t.live = false;
  }
  return t;
}

// barycentric inside test + interpolated depth of pixel centre (px, py) (CheckPixelInsideFace, blur_radius = 0,
// perspective_correct = false, clip_barycentric_coords = false)
__device__ __forceinline__ bool sil_covers(const SilFace& t, float px, float py, float* pz) {
  const float area = sil_edge(t.x2, t.y2, t.x0, t.y0, t.x1, t.y1) + SIL_EPS;     // BarycentricCoordsForward
  const float e0 = sil_edge(px, py, t.x1, t.y1, t.x2, t.y2);
  const float e1 = sil_edge(px, py, t.x2, t.y2, t.x0, t.y0);
  const float e2 = sil_edge(px, py, t.x0, t.y0, t.x1, t.y1);
  // w_i = e_i / area > 0 for all i: decided on the signs, the three divisions only for the pixels that pass
  const bool pos = area > 0.f;
  if (!(pos ? (e0 > 0.f && e1 > 0.f && e2 > 0.f) : (e0 < 0.f && e1 < 0.f && e2 < 0.f))) return false;
  const float w0 = e0 / area, w1 = e1 / area, w2 = e2 / area;
  *pz = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
  return w0 > 0.f && w1 > 0.f && w2 > 0.f && !(*pz < 0.f);
}

struct SilScale { float x, y, z; };     // world = scale * vertex: (-2, -2, 2) for the body model's output (render_mesh), 1 for a ready mesh

__global__ void sil_project_kernel(const float* __restrict__ verts, const float* __restrict__ cam, int64_t n, int64_t V,
                                   float focal, SilScale sc, float* __restrict__ ndc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int64_t b = idx / V;
  const float X = sc.x * verts[idx * 3 + 0] + cam[b * 3 + 0];
  const float Y = sc.y * verts[idx * 3 + 1] + cam[b * 3 + 1];
  const float Z = sc.z * verts[idx * 3 + 2] + cam[b * 3 + 2];
  ndc[idx * 3 + 0] = focal * X / Z;
  ndc[idx * 3 + 1] = focal * Y / Z;
  ndc[idx * 3 + 2] = Z;
}

// A face whose box holds more than SIL_BIG pixel centres is not walked by its own thread (one lane looping over thousands of
// centres while 31 wait) but by the whole warp: its data is broadcast by shuffles and the lanes stride over the box.  SMPL faces
// at 224 x 224 never get there; a close-up camera, a coarse mesh or a larger image do.
constexpr int SIL_BIG = 64;
constexpr bool SIL_FLATTEN = true;     // raster kernel: the warp's small boxes as one pixel list (false: a loop per lane)
constexpr unsigned SIL_FULL = 0xffffffffu;

__device__ __forceinline__ SilFace sil_shfl_face(const SilFace& t, int src) {
  SilFace s;
  s.x0 = __shfl_sync(SIL_FULL, t.x0, src); s.y0 = __shfl_sync(SIL_FULL, t.y0, src); s.z0 = __shfl_sync(SIL_FULL, t.z0, src);
  s.x1 = __shfl_sync(SIL_FULL, t.x1, src); s.y1 = __shfl_sync(SIL_FULL, t.y1, src); s.z1 = __shfl_sync(SIL_FULL, t.z1, src);
  s.x2 = __shfl_sync(SIL_FULL, t.x2, src); s.y2 = __shfl_sync(SIL_FULL, t.y2, src); s.z2 = __shfl_sync(SIL_FULL, t.z2, src);
  s.lo_x = __shfl_sync(SIL_FULL, t.lo_x, src); s.hi_x = __shfl_sync(SIL_FULL, t.hi_x, src);
  s.lo_y = __shfl_sync(SIL_FULL, t.lo_y, src); s.hi_y = __shfl_sync(SIL_FULL, t.hi_y, src);
  s.live = true;
  return s;
}
__device__ __forceinline__ int sil_box(const SilFace& t) { return t.live ? (t.hi_x - t.lo_x + 1) * (t.hi_y - t.lo_y + 1) : 0; }

__device__ __forceinline__ void sil_raster_pixel(const SilFace& t, int xi, int yi, int S, uint32_t f, unsigned long long* zb) {
  float pz;
  if (!sil_covers(t, sil_pix_to_ndc(xi, S), sil_pix_to_ndc(yi, S), &pz)) return;
  // pz >= 0: its bit pattern orders like the value; ties go to the lower face id
  const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned long long)f;
  atomicMin(zb + (int64_t)(S - 1 - yi) * S + (S - 1 - xi), key);
}

__global__ void sil_raster_kernel(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int64_t B, int64_t V,
                                  int64_t F, int S, unsigned long long* __restrict__ zbuf) {
  // grid (faces, frames); whole warps stay for the cooperative parts.  (Measured, 1024 frames at 224 x 224: with a loop per
  // lane 0.72 ms, issue slots 77 % busy, ~1 460 instructions per warp of 32 faces -- the warp waits for its largest box (~30
  // centres) while the mean is 9; neither a lower SIL_BIG nor dropping the 64-bit index division changed that.  Flattened
  // into one pixel list per warp: 0.45 ms.)
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  const bool in = f < F;
  SilFace t = sil_load_face(ndc + b * V * 3, faces, in ? f : 0, S);
  t.live = t.live && in;
  const int box = sil_box(t);
  unsigned long long* zb = zbuf + b * (int64_t)S * S;
  const int lane = threadIdx.x & 31;
  static_assert(SIL_BIG <= 64, "the flattened list indexes a box with a float quotient");
  if (SIL_FLATTEN) {
    // the warp's small boxes as ONE list of pixel centres: entry q belongs to the last lane whose exclusive prefix is <= q
    // (binary search over the lanes' registers), whose face arrives by shuffles -- every lane tests a centre in every round
    // instead of waiting for the warp's largest box
    const int n = box <= SIL_BIG ? box : 0;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(SIL_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - n;
    const int total = __shfl_sync(SIL_FULL, incl, 31);
    const int nx_own = t.hi_x - t.lo_x + 1;
    for (int q0 = 0; q0 < total; q0 += 32) {
      const int q = q0 + lane;
      const bool act = q < total;
      const int qq = act ? q : 0;
      int owner = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int e = __shfl_sync(SIL_FULL, excl, (owner + step) & 31);
        if (owner + step < 32 && e <= qq) owner += step;
      }
      const SilFace s = sil_shfl_face(t, owner);
      const int local = qq - __shfl_sync(SIL_FULL, excl, owner);
      const int nx = __shfl_sync(SIL_FULL, nx_own, owner);
      const int sf = __shfl_sync(SIL_FULL, f, owner);
      const int ly = (int)(((float)local + 0.5f) / (float)nx);      // exact for local < 64
      if (act) sil_raster_pixel(s, s.lo_x + (local - ly * nx), s.lo_y + ly, S, (uint32_t)sf, zb);
    }
  } else if (box > 0 && box <= SIL_BIG) {
    for (int yi = t.lo_y; yi <= t.hi_y; yi++)
      for (int xi = t.lo_x; xi <= t.hi_x; xi++) sil_raster_pixel(t, xi, yi, S, (uint32_t)f, zb);
  }
  unsigned todo = __ballot_sync(SIL_FULL, box > SIL_BIG);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const SilFace s = sil_shfl_face(t, src);
    const int sf = __shfl_sync(SIL_FULL, f, src);
    const int nx = s.hi_x - s.lo_x + 1, n = nx * (s.hi_y - s.lo_y + 1);
    for (int p = lane; p < n; p += 32) sil_raster_pixel(s, s.lo_x + p % nx, s.lo_y + p / nx, S, (uint32_t)sf, zb);
  }
}

// squared distance of the pixel centre to the closest edge of face t, which edge, and its segment parameter
__device__ __forceinline__ float sil_tri_dist(const SilFace& t, float px, float py, int* edge, float* tt) {
  float t01, t02, t12;
  const float e01 = sil_seg_dist(px, py, t.x0, t.y0, t.x1, t.y1, &t01);
  const float e02 = sil_seg_dist(px, py, t.x0, t.y0, t.x2, t.y2, &t02);
  const float e12 = sil_seg_dist(px, py, t.x1, t.y1, t.x2, t.y2, &t12);
  // PointTriangleDistanceForward: min(e01, e02, e12); the backward picks the first edge that attains it in this order
  if (e01 <= e02 && e01 <= e12) { *edge = 0; *tt = t01; return e01; }
  if (e02 <= e01 && e02 <= e12) { *edge = 1; *tt = t02; return e02; }
  *edge = 2; *tt = t12;
  return e12;
}

__global__ void sil_shade_kernel(const float* __restrict__ ndc, const int32_t* __restrict__ faces,
                                 const unsigned long long* __restrict__ zbuf, int64_t B, int64_t V, int S, float inv_sigma,
                                 float* __restrict__ alpha, int32_t* __restrict__ pix_to_face) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;      // grid (pixels, frames)
  const int npix = S * S;
  if (pix >= npix) return;
  const int64_t b = blockIdx.y;
  const int64_t idx = b * npix + pix;
  const unsigned long long key = zbuf[idx];
  if (key == ~0ull) {
    alpha[idx] = 0.f;
    pix_to_face[idx] = -1;
    return;
  }
  const int r = pix / S, c = pix % S;
  const int32_t f = (int32_t)(key & 0xffffffffull);
  const SilFace t = sil_load_face(ndc + b * V * 3, faces, f, S);
  int edge;
  float tt;
  const float d2 = sil_tri_dist(t, sil_pix_to_ndc(S - 1 - c, S), sil_pix_to_ndc(S - 1 - r, S), &edge, &tt);
  alpha[idx] = 1.f / (1.f + __expf(-d2 * inv_sigma));
  pix_to_face[idx] = f;
}

// sum over the frame's pixels of (alpha - target)^2, fixed order: thread-strided partials, shared-memory tree
__global__ void __launch_bounds__(256)
sil_loss_kernel(const float* __restrict__ alpha, const float* __restrict__ target, int64_t npix, float* __restrict__ frame_loss) {
  __shared__ float red[256];
  const int64_t b = blockIdx.x;
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < npix; i += 256) {
    const float d = alpha[b * npix + i] - target[b * npix + i];
    a = fmaf(d, d, a);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) frame_loss[b] = red[0];
}

__global__ void __launch_bounds__(256)
sil_loss_finish_kernel(const float* __restrict__ frame_loss, int64_t B, float scale, float* __restrict__ loss_out) {
  __shared__ float red[256];
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += 256) a += frame_loss[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = red[0] * scale;
}

// d loss / d (the face's three projected corners, x and y), summed over the pixels the face won
__device__ __forceinline__ void sil_grad_pixel(const SilFace& t, int xi, int yi, int S, int32_t f, int64_t base,
                                               const int32_t* __restrict__ pix_to_face, const float* __restrict__ alpha,
                                               const float* __restrict__ dalpha, const float* __restrict__ target,
                                               float mse_scale, float inv_sigma, float g[6]) {
  const int64_t pix = base + (int64_t)(S - 1 - yi) * S + (S - 1 - xi);
  if (pix_to_face[pix] != f) return;
  const float px = sil_pix_to_ndc(xi, S), py = sil_pix_to_ndc(yi, S);
  const float a = alpha[pix];
  const float up = dalpha != nullptr ? dalpha[pix] : mse_scale * 2.f * (a - target[pix]);
  // alpha = sigmoid(d2 / sigma); rasterize_meshes backward: PointLineDistanceBackward on the closest edge with the
  // segment parameter held fixed: grad_v0 = g (1 - t) 2 (q - p), grad_v1 = g t 2 (q - p), q = the closest point
  const float gd = up * a * (1.f - a) * inv_sigma;
  int edge;
  float tt;
  sil_tri_dist(t, px, py, &edge, &tt);
  const int ia = edge == 2 ? 1 : 0, ib = edge == 0 ? 1 : 2;
  const float ax = ia == 0 ? t.x0 : t.x1, ay = ia == 0 ? t.y0 : t.y1;
  const float bx = ib == 1 ? t.x1 : t.x2, by = ib == 1 ? t.y1 : t.y2;
  const float bax = bx - ax, bay = by - ay;
  float ga[2] = {0.f, 0.f}, gb[2];
  if (bax * bax + bay * bay <= SIL_EPS) {
    // degenerate edge: distance to b; grad_v1 = -2 (p - b) g
    gb[0] = -2.f * (px - bx) * gd;
    gb[1] = -2.f * (py - by) * gd;
  } else {
    const float qx = ax + tt * bax - px, qy = ay + tt * bay - py;       // q - p
    ga[0] = gd * (1.f - tt) * 2.f * qx;
    ga[1] = gd * (1.f - tt) * 2.f * qy;
    gb[0] = gd * tt * 2.f * qx;
    gb[1] = gd * tt * 2.f * qy;
  }
  // (no dynamic register indexing: the corner of each end point is one of two)
  if (ia == 0) { g[0] += ga[0]; g[1] += ga[1]; } else { g[2] += ga[0]; g[3] += ga[1]; }
  if (ib == 1) { g[2] += gb[0]; g[3] += gb[1]; } else { g[4] += gb[0]; g[5] += gb[1]; }
}

__global__ void sil_face_grad_kernel(const float* __restrict__ ndc, const int32_t* __restrict__ faces,
                                     const int32_t* __restrict__ pix_to_face, const float* __restrict__ alpha,
                                     const float* __restrict__ dalpha, const float* __restrict__ target, float mse_scale,
                                     int64_t B, int64_t V, int64_t F, int S, float inv_sigma, float* __restrict__ gface) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;        // grid (faces, frames)
  const int64_t b = blockIdx.y;
  const bool in = f < F;
  const int64_t idx = b * F + f;
  const int64_t base = b * (int64_t)S * S;
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  SilFace t = sil_load_face(ndc + b * V * 3, faces, in ? f : 0, S);
  t.live = t.live && in;
  const int box = sil_box(t);
  const int lane = threadIdx.x & 31;
  if (SIL_FLATTEN) {
    // phase 1, the warp's small boxes as one list of pixel centres (as in the raster kernel): which centres of its box did
    // each face WIN?  One bit per centre, OR-ed into the owner's 64-bit word (an order-free atomic on shared memory).
    // phase 2: every lane walks the set bits of its own word in ascending order -- the few pixels it has a gradient for.
    __shared__ unsigned long long won[4][32];
    const int wp = threadIdx.x >> 5;
    won[wp][lane] = 0ull;
    __syncwarp();
    const int n = box <= SIL_BIG ? box : 0;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(SIL_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - n;
    const int total = __shfl_sync(SIL_FULL, incl, 31);
    const int nx_own = t.hi_x - t.lo_x + 1;
    for (int q0 = 0; q0 < total; q0 += 32) {
      const int q = q0 + lane;
      const bool act = q < total;
      const int qq = act ? q : 0;
      int owner = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int e = __shfl_sync(SIL_FULL, excl, (owner + step) & 31);
        if (owner + step < 32 && e <= qq) owner += step;
      }
      const int local = qq - __shfl_sync(SIL_FULL, excl, owner);
      const int nx = __shfl_sync(SIL_FULL, nx_own, owner);
      const int ox = __shfl_sync(SIL_FULL, t.lo_x, owner), oy = __shfl_sync(SIL_FULL, t.lo_y, owner);
      const int sf = __shfl_sync(SIL_FULL, f, owner);
      const int ly = (int)(((float)local + 0.5f) / (float)nx);
      const int xi = ox + (local - ly * nx), yi = oy + ly;
      if (act && pix_to_face[base + (int64_t)(S - 1 - yi) * S + (S - 1 - xi)] == sf) atomicOr(&won[wp][owner], 1ull << local);
    }
    __syncwarp();
    unsigned long long mine = won[wp][lane];
    while (mine) {
      const int local = __ffsll((long long)mine) - 1;
      mine &= mine - 1;
      const int ly = (int)(((float)local + 0.5f) / (float)nx_own);
      sil_grad_pixel(t, t.lo_x + (local - ly * nx_own), t.lo_y + ly, S, (int32_t)f, base, pix_to_face, alpha, dalpha, target,
                     mse_scale, inv_sigma, g);
    }
  } else if (box > 0 && box <= SIL_BIG) {
    for (int yi = t.lo_y; yi <= t.hi_y; yi++)
      for (int xi = t.lo_x; xi <= t.hi_x; xi++)
        sil_grad_pixel(t, xi, yi, S, (int32_t)f, base, pix_to_face, alpha, dalpha, target, mse_scale, inv_sigma, g);
  }
  // large boxes: the warp strides over the box, then a butterfly (a fixed order) sums the lanes' shares for the owner
  unsigned todo = __ballot_sync(SIL_FULL, box > SIL_BIG);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const SilFace s = sil_shfl_face(t, src);
    const int sf = __shfl_sync(SIL_FULL, f, src);
    const int nx = s.hi_x - s.lo_x + 1, n = nx * (s.hi_y - s.lo_y + 1);
    float gs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = lane; p < n; p += 32)
      sil_grad_pixel(s, s.lo_x + p % nx, s.lo_y + p / nx, S, (int32_t)sf, base, pix_to_face, alpha,
                     dalpha, target, mse_scale, inv_sigma, gs);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int e = 0; e < 6; e++) gs[e] += __shfl_xor_sync(SIL_FULL, gs[e], o);
    if (lane == src) {
#pragma unroll
      for (int e = 0; e < 6; e++) g[e] = gs[e];
    }
  }
  if (in) {
#pragma unroll
    for (int e = 0; e < 6; e++) gface[idx * 6 + e] = g[e];
  }
}

// per vertex: its faces' corner gradients (CSR, fixed order) chained through ndc = f (X, Y) / Z, view = (-2x, -2y, 2z) + cam
__global__ void sil_vertex_grad_kernel(const float* __restrict__ verts, const float* __restrict__ cam,
                                       const float* __restrict__ gface, const int32_t* __restrict__ vf_ptr,
                                       const int32_t* __restrict__ vf_idx, int64_t B, int64_t V, int64_t F, float focal,
                                       SilScale sc, float* __restrict__ dverts, float* __restrict__ dview) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * V) return;
  const int64_t b = idx / V, v = idx % V;
  float gx = 0.f, gy = 0.f;
  for (int p = vf_ptr[v]; p < vf_ptr[v + 1]; p++) {
    const int fc = vf_idx[p];        // face * 3 + corner
    gx += gface[(b * F + fc / 3) * 6 + (fc % 3) * 2 + 0];
    gy += gface[(b * F + fc / 3) * 6 + (fc % 3) * 2 + 1];
  }
  const float X = sc.x * verts[idx * 3 + 0] + cam[b * 3 + 0];
  const float Y = sc.y * verts[idx * 3 + 1] + cam[b * 3 + 1];
  const float Z = sc.z * verts[idx * 3 + 2] + cam[b * 3 + 2];
  const float dX = focal / Z * gx, dY = focal / Z * gy;
  const float dZ = -focal * (X * gx + Y * gy) / (Z * Z);
  dverts[idx * 3 + 0] = sc.x * dX;
  dverts[idx * 3 + 1] = sc.y * dY;
  dverts[idx * 3 + 2] = sc.z * dZ;
  if (dview != nullptr) { dview[idx * 3 + 0] = dX; dview[idx * 3 + 1] = dY; dview[idx * 3 + 2] = dZ; }
}

// d loss / d cam = the sum of the view-space gradients of the frame's vertices (fixed order)
__global__ void __launch_bounds__(256)
sil_cam_grad_kernel(const float* __restrict__ dview, int64_t V, float* __restrict__ dcam) {
  __shared__ float red[3][256];
  const int64_t b = blockIdx.x;
  float a[3] = {0.f, 0.f, 0.f};
  for (int64_t v = threadIdx.x; v < V; v += 256) {
    a[0] += dview[(b * V + v) * 3 + 0];
    a[1] += dview[(b * V + v) * 3 + 1];
    a[2] += dview[(b * V + v) * 3 + 2];
  }
  for (int c = 0; c < 3; c++) red[c][threadIdx.x] = a[c];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int c = 0; c < 3; c++) red[c][threadIdx.x] += red[c][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) dcam[b * 3 + threadIdx.x] = red[threadIdx.x][0];
}

static int sil_check(int64_t B, int64_t V, int64_t F, int S, float focal, float sigma) {
  if (B <= 0 || V <= 0 || F <= 0) return fail(JRR_ERR_INVALID, "silhouette: empty batch / mesh");
  if (S < 1 || S > 4096) return fail(JRR_ERR_INVALID, "silhouette: image size out of range");
  if (!(focal > 0.f) || !(sigma > 0.f)) return fail(JRR_ERR_INVALID, "silhouette: focal length and sigma must be positive");
  if (B > 65535) return fail(JRR_ERR_INVALID, "silhouette: at most 65535 frames per call (chunk the batch)");
  if (F >= (1ll << 31) || V >= (1ll << 31)) return fail(JRR_ERR_INVALID, "silhouette: mesh too large");
  return JRR_OK;
}

}  // namespace jrr

using namespace jrr;

extern "C" size_t jrr_silhouette_workspace_bytes(int64_t B, int64_t V, int64_t F, int S) {
  // ndc [B,V,3] f32 | z-buffer [B,S,S] u64 | face gradients [B,F,6] f32 | frame losses [B] f32
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  return up((size_t)B * V * 3 * 4) + up((size_t)B * S * S * 8) + up((size_t)B * F * 6 * 4) + up((size_t)B * 4) + 256;
}

namespace {
struct SilWs { float* ndc; unsigned long long* zbuf; float* gface; float* frame_loss; };
SilWs sil_carve(void* ws, int64_t B, int64_t V, int64_t F, int S) {
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* p = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  SilWs w;
  w.ndc = (float*)p; p += up((size_t)B * V * 3 * 4);
  w.zbuf = (unsigned long long*)p; p += up((size_t)B * S * S * 8);
  w.gface = (float*)p; p += up((size_t)B * F * 6 * 4);
  w.frame_loss = (float*)p;
  return w;
}
}  // namespace

extern "C" int jrr_silhouette_forward(int64_t B, const float* vertices, int64_t V, const float* cam, const int32_t* faces,
                                      int64_t F, int image_size, float focal, float sigma, int flip_scale, const float* target,
                                      int64_t B_logical, float* alpha_out, int32_t* pix_to_face_out, float* loss_out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = sil_check(B, V, F, image_size, focal, sigma)) return rc;
  if (!vertices || !cam || !faces || !alpha_out || !pix_to_face_out || !workspace) return fail(JRR_ERR_INVALID, "silhouette: null argument");
  if (workspace_bytes < jrr_silhouette_workspace_bytes(B, V, F, image_size)) return fail(JRR_ERR_INVALID, "silhouette: workspace too small");
  if (loss_out && (!target || B_logical <= 0)) return fail(JRR_ERR_INVALID, "silhouette: the loss needs a target and the logical batch size");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = image_size;
  const SilWs w = sil_carve(workspace, B, V, F, S);
  const int64_t npix = (int64_t)S * S;
  const SilScale sc = flip_scale ? SilScale{-2.f, -2.f, 2.f} : SilScale{1.f, 1.f, 1.f};
  sil_project_kernel<<<(unsigned)((B * V + 255) / 256), 256, 0, st>>>(vertices, cam, B * V, V, focal, sc, w.ndc);
  JRR_LAUNCH_CHECK();
  JRR_CUDA(cudaMemsetAsync(w.zbuf, 0xff, (size_t)B * npix * 8, st));
  sil_raster_kernel<<<dim3((unsigned)((F + 127) / 128), (unsigned)B), 128, 0, st>>>(w.ndc, faces, B, V, F, S, w.zbuf);
  JRR_LAUNCH_CHECK();
  sil_shade_kernel<<<dim3((unsigned)((npix + 255) / 256), (unsigned)B), 256, 0, st>>>(w.ndc, faces, w.zbuf, B, V, S, 1.f / sigma, alpha_out,
                                                                      pix_to_face_out);
  JRR_LAUNCH_CHECK();
  if (loss_out) {
    // nn.MSELoss over [B_logical, 1, S, S] (optimize.py:128,236)
    sil_loss_kernel<<<(unsigned)B, 256, 0, st>>>(alpha_out, target, npix, w.frame_loss);
    JRR_LAUNCH_CHECK();
    sil_loss_finish_kernel<<<1, 256, 0, st>>>(w.frame_loss, B, 1.f / ((float)B_logical * (float)npix), loss_out);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}

extern "C" int jrr_silhouette_backward(int64_t B, const float* vertices, int64_t V, const float* cam, const int32_t* faces,
                                       int64_t F, const int32_t* vert_face_ptr, const int32_t* vert_face_idx, int image_size,
                                       float focal, float sigma, int flip_scale, const float* alpha, const int32_t* pix_to_face,
                                       const float* dalpha, const float* target, int64_t B_logical, float loss_weight,
                                       float* dvertices_out, float* dcam_out, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (int rc = sil_check(B, V, F, image_size, focal, sigma)) return rc;
  if (!vertices || !cam || !faces || !vert_face_ptr || !vert_face_idx || !alpha || !pix_to_face || !dvertices_out || !workspace)
    return fail(JRR_ERR_INVALID, "silhouette: null argument");
  if (!dalpha && (!target || B_logical <= 0)) return fail(JRR_ERR_INVALID, "silhouette backward: give d loss / d alpha, or the MSE target");
  if (workspace_bytes < jrr_silhouette_workspace_bytes(B, V, F, image_size)) return fail(JRR_ERR_INVALID, "silhouette: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = image_size;
  const SilWs w = sil_carve(workspace, B, V, F, S);     // ndc is the forward's
  const float mse_scale = dalpha ? 0.f : loss_weight / ((float)B_logical * (float)S * (float)S);
  sil_face_grad_kernel<<<dim3((unsigned)((F + 127) / 128), (unsigned)B), 128, 0, st>>>(w.ndc, faces, pix_to_face, alpha, dalpha, target, mse_scale,
                                                                       B, V, F, S, 1.f / sigma, w.gface);
  JRR_LAUNCH_CHECK();
  // (the projected vertices are not needed after the face pass: their buffer takes the view-space gradients)
  const SilScale sc = flip_scale ? SilScale{-2.f, -2.f, 2.f} : SilScale{1.f, 1.f, 1.f};
  sil_vertex_grad_kernel<<<(unsigned)((B * V + 255) / 256), 256, 0, st>>>(vertices, cam, w.gface, vert_face_ptr, vert_face_idx, B, V,
                                                                         F, focal, sc, dvertices_out, dcam_out ? w.ndc : nullptr);
  JRR_LAUNCH_CHECK();
  if (dcam_out) {
    sil_cam_grad_kernel<<<(unsigned)B, 256, 0, st>>>(w.ndc, V, dcam_out);
    JRR_LAUNCH_CHECK();
  }
  return JRR_OK;
}
