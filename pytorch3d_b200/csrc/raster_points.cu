// Point-cloud rasterizer for sm_100a: setup/bin pass, per-tile fine pass (top-K per pixel), backward.
//
// Replaces, behind the same operator signature, the reference's
//   PointBoundingBoxKernel + RasterizeCoarseCudaKernel   (rasterize_coarse.cu:53-74, 76-219)
//   RasterizePointsFineCudaKernel / NaiveCudaKernel       (rasterize_points.cu:223-298, 87-149)
//   RasterizePointsBackwardCudaKernel                     (rasterize_points.cu:366-411)
// Same skeleton as raster_meshes.cu: exact tile binning, one CTA per 16x16 tile, points staged in
// shared memory as 16-byte (x, y, z, r^2) records, warp-footprint culling by ballot, register top-K.
#include <cfloat>
#include <climits>

#include "binning.cuh"
#include "bulk_copy.cuh"
#include "common.cuh"
#include "raster_math.cuh"

namespace b200r {

constexpr int SETUP_POINTS = 256;
constexpr int PCHUNK = 512;  // points staged per round (2 per thread)

// Pass 1: per-point box (x +- r, y +- r), skip z < 0 (rasterize_coarse.cu:53-74), count per tile.
// The CTA's 256 points (3072 contiguous bytes of the packed (P,3) array) arrive by one TMA bulk copy.
__global__ void __launch_bounds__(SETUP_POINTS)
    points_setup_count_kernel(const float* __restrict__ points, const float* __restrict__ radius, int64_t P,
                              const int64_t* __restrict__ first, const int64_t* __restrict__ num, int N, int H,
                              int W, int TY, int TX, float rx, float ry, uint4* __restrict__ rect,
                              int* __restrict__ tile_count) {
  __shared__ __align__(16) float s_pts[SETUP_POINTS * 3];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * SETUP_POINTS;
  const int np = (int)min((int64_t)SETUP_POINTS, P - p0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  cta_load_words(s_pts, points + p0 * 3, np * 3, &bar, 0);
  uint2 rc = make_uint2(RECT_EMPTY_X, 0u);
  int n = -1;
  const int64_t pi = p0 + tid;
  if (tid < np) {
    const float x = s_pts[tid * 3 + 0], y = s_pts[tid * 3 + 1], z = s_pts[tid * 3 + 2];  // stride 3: conflict-free
    const float r = __ldg(radius + pi);
    n = find_owner(first, num, N, pi);
    if (n >= 0 && !(z < 0.0f)) rc = bbox_to_tile_rect(fsub(x, r), fadd(x, r), fsub(y, r), fadd(y, r), H, W, rx, ry);
    rect[pi] = make_uint4(rc.x, rc.y, (uint32_t)max(n, 0), 0u);
  }
  warp_count_rect(rc, n, TY, TX, tile_count, tid & 31);  // all lanes participate
}

// The K nearest points of one pixel: the reference's queue (rasterize_points.cu:61-79) restated for
// registers -- an UNSORTED array of K slots plus the tracked maximum z; a hit fills the next free slot or,
// when full and pz < q_max_z, overwrites the tracked maximum, which is then searched again.  Points arrive
// in ascending index order (sorted tile lists), so tie behaviour equals the reference's naive kernel.
template <int KMAX>
struct PTopK {
  float z[KMAX];
  int id[KMAX];
  float d[KMAX];
  int size;
  float max_z;
  int max_idx;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      z[i] = -1.0f;
      id[i] = -1;
      d[i] = -1.0f;
    }
    size = 0;
    max_z = -1000.0f;
    max_idx = -1;
  }
  __device__ __forceinline__ void put(int slot, float pz, int p, float d2) {
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const bool w = i == slot;
      z[i] = w ? pz : z[i];
      id[i] = w ? p : id[i];
      d[i] = w ? d2 : d[i];
    }
  }
  __device__ __forceinline__ void offer(float pz, int p, float d2, int K) {
    if (size < K) {
      put(size, pz, p, d2);
      if (pz > max_z) {
        max_z = pz;
        max_idx = size;
      }
      ++size;
    } else if (pz < max_z) {
      put(max_idx, pz, p, d2);
      max_z = pz;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        if (i < K && z[i] > max_z) {
          max_z = z[i];
          max_idx = i;
        }
      }
    }
  }
  // BubbleSort on z only (rasterize_points.cu:26-28): a STABLE sort, so z-ties keep slot order.  An
  // odd-even transposition network with a strict compare is stable too and gives the same permutation.
  __device__ __forceinline__ void sort() {
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i >= size) z[i] = FLT_MAX;
#pragma unroll
    for (int r = 0; r < KMAX; ++r) {
#pragma unroll
      for (int i = r & 1; i + 1 < KMAX; i += 2) {
        if (z[i + 1] < z[i]) {
          float t;
          int ti;
          t = z[i]; z[i] = z[i + 1]; z[i + 1] = t;
          ti = id[i]; id[i] = id[i + 1]; id[i + 1] = ti;
          t = d[i]; d[i] = d[i + 1]; d[i + 1] = t;
        }
      }
    }
  }
};

struct __align__(16) PointChunk {
  float4 box[PCHUNK];  // xmin, xmax, ymin, ymax (empty = never hit)
  float4 rec[PCHUNK];  // x, y, z, r^2
  int id[PCHUNK];
};

struct PointFineParams {
  const float* points;
  const float* radius;
  const int64_t* first;
  const int64_t* num;
  const int* tile_offset;
  const int* pairs;
  int64_t capacity;
  int N, H, W, K, TY, TX;
  float rx, ry;
  int32_t* idx;
  float* zbuf;
  float* dists;
};

__device__ __forceinline__ void stage_point(PointChunk& s, int slot, const float* __restrict__ points,
                                            const float* __restrict__ radius, int pi) {
  const float x = __ldg(points + (int64_t)pi * 3 + 0), y = __ldg(points + (int64_t)pi * 3 + 1),
              z = __ldg(points + (int64_t)pi * 3 + 2);
  const float r = __ldg(radius + pi);
  float xmin = FLT_MAX, xmax = -FLT_MAX, ymin = FLT_MAX, ymax = -FLT_MAX;
  if (!(z < 0.0f)) {  // points behind the camera are not rendered (rasterize_points.cu:55-56)
    xmin = fsub(x, r);
    xmax = fadd(x, r);
    ymin = fsub(y, r);
    ymax = fadd(y, r);
  }
  s.box[slot] = make_float4(xmin, xmax, ymin, ymax);
  s.rec[slot] = make_float4(x, y, z, fmul(r, r));
  s.id[slot] = pi;
}

__device__ __forceinline__ void pthread_pixel(int tile_x, int tile_y, int& xo, int& yo) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  xo = tile_x * TILE + (w & 1) * 8 + (lane & 7);
  yo = tile_y * TILE + (w >> 1) * 4 + (lane >> 3);
}

// KMAX > 0: register top-K; KMAX == 0: thread-local arrays for K up to 150.
template <int KMAX>
__global__ void __launch_bounds__(TILE_THREADS) points_fine_kernel(const PointFineParams p) {
  __shared__ PointChunk s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int t = blockIdx.x;
  const int n = t / (p.TY * p.TX);
  const int tile_y = (t / p.TX) % p.TY, tile_x = t % p.TX;
  int xo, yo;
  pthread_pixel(tile_x, tile_y, xo, yo);
  const bool valid = xo < p.W && yo < p.H;
  const float px = pix_to_ndc(p.W - 1 - xo, p.W, p.rx);
  const float py = pix_to_ndc(p.H - 1 - yo, p.H, p.ry);
  float col[8], row[4];  // the footprint's 8 column and 4 row coordinates (lane = row * 8 + column)
#pragma unroll
  for (int c = 0; c < 8; ++c) col[c] = __shfl_sync(0xffffffffu, px, c);
#pragma unroll
  for (int r = 0; r < 4; ++r) row[r] = __shfl_sync(0xffffffffu, py, 8 * r);

  const int seg_begin = p.tile_offset[t], seg_end = p.tile_offset[t + 1];
  const bool overflow = (int64_t)seg_end > p.capacity || seg_end == INT_MAX;
  const int64_t cloud_first = p.first[n];
  const int count = overflow ? (int)p.num[n] : seg_end - seg_begin;
  const int K = p.K;

  constexpr int QN = KMAX > 0 ? KMAX : 1;
  PTopK<QN> q;
  float lz[KMAX > 0 ? 1 : B200R_MAX_K];
  int li[KMAX > 0 ? 1 : B200R_MAX_K];
  float ld[KMAX > 0 ? 1 : B200R_MAX_K];
  int ln = 0, l_max_idx = -1;
  float l_max_z = -1000.0f;
  if (KMAX > 0) q.init();

  for (int base = 0; base < count; base += PCHUNK) {
    const int nc = min(PCHUNK, count - base);
    __syncthreads();
    for (int j = tid; j < nc; j += TILE_THREADS) {
      const int pi = overflow ? (int)(cloud_first + base + j) : p.pairs[seg_begin + base + j];
      stage_point(s, j, p.points, p.radius, pi);
    }
    __syncthreads();
    for (int sub = 0; sub < nc; sub += 64) {
      // pass A: 64-bit mask of the points of this round whose box contains my pixel (see raster_meshes.cu)
      unsigned m0 = 0, m1 = 0;
      if (sub + lane < nc) m0 = box_pixel_mask(s.box[sub + lane], col, row);
      if (sub + 32 + lane < nc) m1 = box_pixel_mask(s.box[sub + 32 + lane], col, row);
      m0 = warp_transpose_bits(m0, lane);
      if (sub + 32 < nc) m1 = warp_transpose_bits(m1, lane);
      unsigned long long mine = valid ? (((unsigned long long)m1 << 32) | m0) : 0ull;
      // pass B: every lane tests and queues its own candidates, in ascending point order
      while (__any_sync(0xffffffffu, mine != 0ull)) {
        if (mine == 0ull) continue;
        const int j = sub + __ffsll((long long)mine) - 1;
        mine &= mine - 1ull;
        const float4 r = s.rec[j];
        // CheckPixelInsidePoint (rasterize_points.cu:49-60): dist2 = fma(dy, dy, rn(dx*dx)) < rn(r*r)
        const float dx = fsub(px, r.x), dy = fsub(py, r.y);
        const float d2 = sqnorm2(dx, dy);
        if (r.z < 0.0f || !(d2 < r.w)) continue;
        const int pi = s.id[j];
        if (KMAX > 0) {
          q.offer(r.z, pi, d2, K);
        } else if (ln < K) {
          lz[ln] = r.z;
          li[ln] = pi;
          ld[ln] = d2;
          if (r.z > l_max_z) {
            l_max_z = r.z;
            l_max_idx = ln;
          }
          ++ln;
        } else if (r.z < l_max_z) {
          lz[l_max_idx] = r.z;
          li[l_max_idx] = pi;
          ld[l_max_idx] = d2;
          l_max_z = r.z;
          for (int i = 0; i < K; ++i)
            if (lz[i] > l_max_z) {
              l_max_z = lz[i];
              l_max_idx = i;
            }
        }
      }
    }
  }
  if (!valid) return;
  const int64_t o = (((int64_t)n * p.H + yo) * p.W + xo) * K;
  if (KMAX > 0) {
    q.sort();
#pragma unroll
    for (int k = 0; k < QN; ++k) {
      if (k < K) {
        const bool e = k >= q.size;
        p.idx[o + k] = e ? -1 : q.id[k];
        p.zbuf[o + k] = e ? -1.0f : q.z[k];
        p.dists[o + k] = q.d[k];
      }
    }
  } else {
    for (int i = 1; i < ln; ++i) {  // stable insertion sort on z only
      const float tz = lz[i], td = ld[i];
      const int ti = li[i];
      int j = i - 1;
      while (j >= 0 && tz < lz[j]) {
        lz[j + 1] = lz[j];
        li[j + 1] = li[j];
        ld[j + 1] = ld[j];
        --j;
      }
      lz[j + 1] = tz;
      li[j + 1] = ti;
      ld[j + 1] = td;
    }
    for (int k = 0; k < K; ++k) {
      p.idx[o + k] = k < ln ? li[k] : -1;
      p.zbuf[o + k] = k < ln ? lz[k] : -1.0f;
      p.dists[o + k] = k < ln ? ld[k] : -1.0f;
    }
  }
}

// Medium K (5..32): the queue lives in dynamic shared memory as three [K][256] arrays (slot-major, one column
// per thread), where the dynamic slot index of the reference's queue costs nothing: appending a hit is three
// stores instead of 3*KMAX predicated register moves, the kernel needs ~50 registers instead of 117 (K = 10),
// and the final stable sort on z is an insertion sort over the thread's own column.
__global__ void __launch_bounds__(TILE_THREADS) points_fine_smem_kernel(const PointFineParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PointChunk& s = *reinterpret_cast<PointChunk*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = p.K;
  float* qz = reinterpret_cast<float*>(smem_raw + sizeof(PointChunk)) + tid;  // qz[k * 256]
  int* qi = reinterpret_cast<int*>(qz - tid + K * TILE_THREADS) + tid;
  float* qd = reinterpret_cast<float*>(qi - tid + K * TILE_THREADS) + tid;
  const int t = blockIdx.x;
  const int n = t / (p.TY * p.TX);
  const int tile_y = (t / p.TX) % p.TY, tile_x = t % p.TX;
  int xo, yo;
  pthread_pixel(tile_x, tile_y, xo, yo);
  const bool valid = xo < p.W && yo < p.H;
  const float px = pix_to_ndc(p.W - 1 - xo, p.W, p.rx);
  const float py = pix_to_ndc(p.H - 1 - yo, p.H, p.ry);

  const int seg_begin = p.tile_offset[t], seg_end = p.tile_offset[t + 1];
  const bool overflow = (int64_t)seg_end > p.capacity || seg_end == INT_MAX;
  const int64_t cloud_first = p.first[n];
  const int count = overflow ? (int)p.num[n] : seg_end - seg_begin;

  int size = 0, max_idx = -1;
  float max_z = -1000.0f;

  for (int base = 0; base < count; base += PCHUNK) {
    const int nc = min(PCHUNK, count - base);
    __syncthreads();
    for (int j = tid; j < nc; j += TILE_THREADS) {
      const int pi = overflow ? (int)(cloud_first + base + j) : p.pairs[seg_begin + base + j];
      stage_point(s, j, p.points, p.radius, pi);
    }
    __syncthreads();
    for (int sub = 0; sub < nc; sub += 64) {
      unsigned m0 = 0, m1 = 0;
      {
        float col[8], row[4];  // the footprint's 8 column and 4 row coordinates (lane = row * 8 + column)
#pragma unroll
        for (int c = 0; c < 8; ++c) col[c] = __shfl_sync(0xffffffffu, px, c);
#pragma unroll
        for (int r = 0; r < 4; ++r) row[r] = __shfl_sync(0xffffffffu, py, 8 * r);
        if (sub + lane < nc) m0 = box_pixel_mask(s.box[sub + lane], col, row);
        if (sub + 32 + lane < nc) m1 = box_pixel_mask(s.box[sub + 32 + lane], col, row);
      }
      m0 = warp_transpose_bits(m0, lane);
      if (sub + 32 < nc) m1 = warp_transpose_bits(m1, lane);
      unsigned long long mine = valid ? (((unsigned long long)m1 << 32) | m0) : 0ull;
      while (__any_sync(0xffffffffu, mine != 0ull)) {
        if (mine == 0ull) continue;
        const int j = sub + __ffsll((long long)mine) - 1;
        mine &= mine - 1ull;
        const float4 r = s.rec[j];
        // CheckPixelInsidePoint (rasterize_points.cu:49-60): dist2 = fma(dy, dy, rn(dx*dx)) < rn(r*r)
        const float dx = fsub(px, r.x), dy = fsub(py, r.y);
        const float d2 = sqnorm2(dx, dy);
        if (r.z < 0.0f || !(d2 < r.w)) continue;
        const int pi = s.id[j];
        if (size < K) {  // (:61-67)
          qz[size * TILE_THREADS] = r.z;
          qi[size * TILE_THREADS] = pi;
          qd[size * TILE_THREADS] = d2;
          if (r.z > max_z) {
            max_z = r.z;
            max_idx = size;
          }
          ++size;
        } else if (r.z < max_z) {  // (:68-78)
          qz[max_idx * TILE_THREADS] = r.z;
          qi[max_idx * TILE_THREADS] = pi;
          qd[max_idx * TILE_THREADS] = d2;
          max_z = r.z;
          for (int i = 0; i < K; ++i) {
            const float v = qz[i * TILE_THREADS];
            if (v > max_z) {
              max_z = v;
              max_idx = i;
            }
          }
        }
      }
    }
  }
  if (!valid) return;
  // BubbleSort on z only (rasterize_points.cu:26-28): stable -> insertion sort over the thread's own column
  for (int i = 1; i < size; ++i) {
    const float tz = qz[i * TILE_THREADS], td = qd[i * TILE_THREADS];
    const int ti = qi[i * TILE_THREADS];
    int j = i - 1;
    while (j >= 0 && tz < qz[j * TILE_THREADS]) {
      qz[(j + 1) * TILE_THREADS] = qz[j * TILE_THREADS];
      qi[(j + 1) * TILE_THREADS] = qi[j * TILE_THREADS];
      qd[(j + 1) * TILE_THREADS] = qd[j * TILE_THREADS];
      --j;
    }
    qz[(j + 1) * TILE_THREADS] = tz;
    qi[(j + 1) * TILE_THREADS] = ti;
    qd[(j + 1) * TILE_THREADS] = td;
  }
  const int64_t o = (((int64_t)n * p.H + yo) * p.W + xo) * K;
  for (int k = 0; k < K; ++k) {
    const bool e = k >= size;
    p.idx[o + k] = e ? -1 : qi[k * TILE_THREADS];
    p.zbuf[o + k] = e ? -1.0f : qz[k * TILE_THREADS];
    p.dists[o + k] = e ? -1.0f : qd[k * TILE_THREADS];
  }
}

// Backward (rasterize_points.cu:366-411): grad_xy = 2 * grad_dist * (p_xy - pix_xy), grad_z = grad_zbuf.
// One thread per pixel on the forward pass's tiles and 8x4 footprints, looping over the K slots.  A point covers
// many neighbouring pixels, so at every slot the warp first merges ALL lanes that hold the same point
// (__match_any_sync + pointer jumping, as in the mesh backward) and only one lane per distinct point issues the
// three atomics.
__global__ void __launch_bounds__(TILE_THREADS)
    points_backward_kernel(const float* __restrict__ points, const int32_t* __restrict__ idxs,
                           const float* __restrict__ grad_zbuf, const float* __restrict__ grad_dists, int n0, int H,
                           int W, int K, float rx, float ry, float* __restrict__ grad_points) {
  const int lane = threadIdx.x & 31;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y, n = n0 + blockIdx.z;
  int xo, yo;
  pthread_pixel(tile_x, tile_y, xo, yo);
  const bool in_image = xo < W && yo < H;
  const float xf = pix_to_ndc(W - 1 - xo, W, rx);
  const float yf = pix_to_ndc(H - 1 - yo, H, ry);
  const int64_t o = in_image ? (((int64_t)n * H + yo) * W + xo) * K : 0;
  for (int k = 0; k < K; ++k) {
    const int pi = in_image ? idxs[o + k] : -1;
    if (!__any_sync(0xffffffffu, pi >= 0)) continue;
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    if (pi >= 0) {
      const float gd = grad_dists[o + k];
      const float g2 = gd + gd;
      gx = g2 * (__ldg(points + (int64_t)pi * 3 + 0) - xf);
      gy = g2 * (__ldg(points + (int64_t)pi * 3 + 1) - yf);
      gz = grad_zbuf[o + k];
    }
    const unsigned grp = __match_any_sync(0xffffffffu, pi);
    const unsigned above = lane == 31 ? 0u : grp & (0xffffffffu << (lane + 1));
    int next = (pi >= 0 && above != 0u) ? __ffs((int)above) - 1 : -1;
    while (__any_sync(0xffffffffu, next >= 0)) {
      const int src = next >= 0 ? next : lane;
      const float vx = __shfl_sync(0xffffffffu, gx, src), vy = __shfl_sync(0xffffffffu, gy, src),
                  vz = __shfl_sync(0xffffffffu, gz, src);
      const int nn = __shfl_sync(0xffffffffu, next, src);
      if (next >= 0) {
        gx += vx;
        gy += vy;
        gz += vz;
      }
      next = next >= 0 ? nn : -1;
    }
    if (pi >= 0 && lane == __ffs((int)grp) - 1) {
      atomicAdd(grad_points + (int64_t)pi * 3 + 0, gx);
      atomicAdd(grad_points + (int64_t)pi * 3 + 1, gy);
      atomicAdd(grad_points + (int64_t)pi * 3 + 2, gz);
    }
  }
}

}  // namespace b200r

using namespace b200r;

extern "C" size_t b200r_rasterize_points_workspace_bytes(int64_t P, int32_t N, int32_t H, int32_t W,
                                                         int64_t pair_capacity) {
  if (P < 0 || N < 0 || H < 0 || W < 0) return 0;
  return carve_workspace(nullptr, P, N, H, W, pair_capacity).bytes;
}

extern "C" int b200r_rasterize_points_forward(const float* points, int64_t P, const int64_t* first,
                                              const int64_t* num, const float* radius, int32_t N, int32_t H,
                                              int32_t W, int32_t K, int32_t bin_size, int32_t max_points_per_bin,
                                              int32_t* idx, float* zbuf, float* dists, void* workspace,
                                              size_t workspace_bytes, int64_t pair_capacity, void* stream_) {
  (void)bin_size;
  (void)max_points_per_bin;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (K > B200R_MAX_K) return fail(B200R_ERR_INVALID_ARGUMENT, "Must have num_closest <= 150");
  if (P < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (P > INT_MAX) return fail(B200R_ERR_INVALID_ARGUMENT, "more than 2^31-1 packed points are not supported");
  if ((int64_t)N * H * W * K == 0) return B200R_OK;
  const int TY = div_up(H, TILE), TX = div_up(W, TILE);
  if (TY > 0xFFFE || TX > 0xFFFE) return fail(B200R_ERR_INVALID_ARGUMENT, "image too large");
  const int64_t ntiles = (int64_t)N * TY * TX;
  if (ntiles > INT_MAX) return fail(B200R_ERR_INVALID_ARGUMENT, "too many tiles");
  BinWorkspace ws = carve_workspace(workspace, P, N, H, W, pair_capacity);
  if (workspace == nullptr || workspace_bytes < ws.bytes)
    return fail(B200R_ERR_WORKSPACE, "workspace too small for rasterize_points_forward");
  const float rx = ndc_range(W, H), ry = ndc_range(H, W);

  const bool prof = profiling_enabled();
  if (prof) phase_timer().record(0, stream);
  B200R_CUDA_OK(cudaMemsetAsync(ws.tile_count, 0, sizeof(int) * (size_t)ntiles, stream));
  if (P > 0) {
    points_setup_count_kernel<<<(unsigned)((P + SETUP_POINTS - 1) / SETUP_POINTS), SETUP_POINTS, 0, stream>>>(
        points, radius, P, first, num, N, H, W, TY, TX, rx, ry, ws.rect, ws.tile_count);
    B200R_LAUNCHED("points_setup_count_kernel");
  }
  tile_scan_kernel<<<1, 1024, 0, stream>>>(ws.tile_count, ws.tile_offset, (int)ntiles);
  B200R_LAUNCHED("tile_scan_kernel");
  if (P > 0) {
    tile_fill_kernel<<<(unsigned)((P + 255) / 256), 256, 0, stream>>>(ws.rect, P, TY, TX, ws.tile_count, ws.pairs,
                                                                    ws.capacity);
    B200R_LAUNCHED("tile_fill_kernel");
  }
  tile_sort_kernel<<<(unsigned)((ntiles + SORT_TILES_PER_CTA - 1) / SORT_TILES_PER_CTA), SORT_THREADS, 0, stream>>>(
        ws.tile_offset, ws.pairs, ws.capacity, (int)ntiles, sort_multiplier(ntiles));
  B200R_LAUNCHED("tile_sort_kernel");
  if (prof) phase_timer().record(1, stream);
  PointFineParams p;
  p.points = points; p.radius = radius; p.first = first; p.num = num;
  p.tile_offset = ws.tile_offset; p.pairs = ws.pairs; p.capacity = ws.capacity;
  p.N = N; p.H = H; p.W = W; p.K = K; p.TY = TY; p.TX = TX; p.rx = rx; p.ry = ry;
  p.idx = idx; p.zbuf = zbuf; p.dists = dists;
  const unsigned grid = (unsigned)ntiles;
  if (K <= 1)
    points_fine_kernel<1><<<grid, TILE_THREADS, 0, stream>>>(p);
  else if (K <= 2)
    points_fine_kernel<2><<<grid, TILE_THREADS, 0, stream>>>(p);
  else if (K <= 4)
    points_fine_kernel<4><<<grid, TILE_THREADS, 0, stream>>>(p);
  else if (K <= 32) {
    const size_t smem = sizeof(PointChunk) + (size_t)K * TILE_THREADS * 12;
    static bool configured[64] = {}; /* > 48 KB of dynamic shared memory: opt-in per kernel and device */
    int dev_ = 0;
    B200R_CUDA_OK(cudaGetDevice(&dev_));
    if (dev_ < 0 || dev_ >= 64 || !configured[dev_]) {
      B200R_CUDA_OK(cudaFuncSetAttribute(points_fine_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(PointChunk) + 32 * TILE_THREADS * 12)));
      if (dev_ >= 0 && dev_ < 64) configured[dev_] = true;
    }
    points_fine_smem_kernel<<<grid, TILE_THREADS, smem, stream>>>(p);
  }
  else
    points_fine_kernel<0><<<grid, TILE_THREADS, 0, stream>>>(p);
  B200R_LAUNCHED("points_fine_kernel");
  if (prof) {
    phase_timer().record(2, stream);
    phase_timer().have_fwd = true;
  }
  return B200R_OK;
}

extern "C" int b200r_rasterize_points_backward(const float* points, int64_t P, const int32_t* idxs,
                                               const float* grad_zbuf, const float* grad_dists, int32_t N,
                                               int32_t H, int32_t W, int32_t K, float* grad_points, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (P == 0) return B200R_OK;
  B200R_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * 3 * (size_t)P, stream));
  const int64_t total = (int64_t)N * H * W * K;
  if (total == 0) return B200R_OK;
  const int TY = div_up(H, TILE), TX = div_up(W, TILE);
  if (TY > 65535) return fail(B200R_ERR_INVALID_ARGUMENT, "image too large");
  const bool prof = profiling_enabled();
  if (prof) phase_timer().record(3, stream);
  for (int n0 = 0; n0 < N; n0 += 65535) {  // grid.z is limited to 65535 images per launch
    const dim3 grid((unsigned)TX, (unsigned)TY, (unsigned)min(N - n0, 65535));
    points_backward_kernel<<<grid, TILE_THREADS, 0, stream>>>(points, idxs, grad_zbuf, grad_dists, n0, H, W, K,
                                                             ndc_range(W, H), ndc_range(H, W), grad_points);
  }
  B200R_LAUNCHED("points_backward_kernel");
  if (prof) {
    phase_timer().record(4, stream);
    phase_timer().have_bwd = true;
  }
  return B200R_OK;
}
