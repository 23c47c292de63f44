// Test hooks of the reference's coarse stage (pytorch3d/csrc/ext.cpp:69-73: `_rasterize_meshes_coarse`,
// `_rasterize_points_coarse`): the dense (N, BH, BW, M) table of element indices per bin, -1 padded.
//
// The product path never builds this table (binning.cuh: exact compact tile lists); these entry points exist so that
// the reference's own tests of its coarse stage (tests/test_rasterize_meshes.py:1096-1163, test_rasterize_points.py:
// 425-460) can run against this build.  Semantics restated from rasterize_coarse.cu:
//   faces : box = [min - sqrt(blur), max + sqrt(blur)] per axis, skipped if zmin < 1e-8              (:20-51)
//   points: box = [x - r, x + r] x [y - r, y + r], skipped if z < 0                                 (:53-74)
//   an element is in bin (by, bx) of its image iff  ymin <= bin_y_max && bin_y_min < ymax  and the same in x, where
//   the bin spans the pixel centres of its first and last pixel widened by half a pixel              (:137-162)
// One thread per element: the candidate bins come from the inverse pixel-centre map (a superset) and are settled with
// the reference's comparisons; a slot in the bin is taken with an atomic, so the order inside a bin is arbitrary -- the
// Python wrapper sorts every bin (the reference's own order depends on its chunking: its tests sort too).
#include "binning.cuh"
#include "common.cuh"
#include "raster_math.cuh"

namespace b200r {

struct CoarseParams {
  int N, H, W, bin_size, BH, BW, M;
  float rx, ry;
  int* bins;      // (N, BH, BW, M), pre-filled with -1
  int* counts;    // (N, BH, BW), zeroed
  int* overflow;  // set to 1 if some bin received more than M elements
};

__device__ __forceinline__ void coarse_insert(const CoarseParams& p, int n, int64_t e, float xmin, float xmax, float ymin,
                                              float ymax) {
  // half a pixel in NDC (rasterize_coarse.cu:105-108): (range / 2) / S
  const float half_pix_x = fdiv(fmul(p.rx, 0.5f), (float)p.W), half_pix_y = fdiv(fmul(p.ry, 0.5f), (float)p.H);
  // candidate pixel range of the box widened by one pixel, then bins
  int ix_lo, ix_hi, iy_lo, iy_hi;
  pixel_range(xmin - 2.0f * half_pix_x, xmax + 2.0f * half_pix_x, p.W, p.rx, ix_lo, ix_hi);
  pixel_range(ymin - 2.0f * half_pix_y, ymax + 2.0f * half_pix_y, p.H, p.ry, iy_lo, iy_hi);
  if (ix_lo > ix_hi || iy_lo > iy_hi) return;
  const int bx_lo = ix_lo / p.bin_size, bx_hi = min(ix_hi / p.bin_size, p.BW - 1);
  const int by_lo = iy_lo / p.bin_size, by_hi = min(iy_hi / p.bin_size, p.BH - 1);
  for (int by = by_lo; by <= by_hi; ++by) {
    const float bin_y_min = fsub(pix_to_ndc(by * p.bin_size, p.H, p.ry), half_pix_y);
    const float bin_y_max = fadd(pix_to_ndc((by + 1) * p.bin_size - 1, p.H, p.ry), half_pix_y);
    if (!((ymin <= bin_y_max) && (bin_y_min < ymax))) continue;
    for (int bx = bx_lo; bx <= bx_hi; ++bx) {
      const float bin_x_max = fadd(pix_to_ndc((bx + 1) * p.bin_size - 1, p.W, p.rx), half_pix_x);
      const float bin_x_min = fsub(pix_to_ndc(bx * p.bin_size, p.W, p.rx), half_pix_x);
      if (!((xmin <= bin_x_max) && (bin_x_min < xmax))) continue;
      const int b = (n * p.BH + by) * p.BW + bx;
      const int slot = atomicAdd(p.counts + b, 1);
      if (slot < p.M)
        p.bins[(int64_t)b * p.M + slot] = (int)e;
      else
        *p.overflow = 1;
    }
  }
}

__global__ void __launch_bounds__(256) coarse_faces_kernel(const float* __restrict__ face_verts, int64_t F,
                                                           const int64_t* __restrict__ first,
                                                           const int64_t* __restrict__ num, float sqrt_blur,
                                                           const CoarseParams p) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int n = find_owner(first, num, p.N, f);
  if (n < 0) return;
  const float* v = face_verts + f * 9;
  const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
  if ((double)fminf(fminf(z0, z1), z2) < kEps) return;  // (:44)
  coarse_insert(p, n, f, fsub(fminf(fminf(x0, x1), x2), sqrt_blur), fadd(fmaxf(fmaxf(x0, x1), x2), sqrt_blur),
                fsub(fminf(fminf(y0, y1), y2), sqrt_blur), fadd(fmaxf(fmaxf(y0, y1), y2), sqrt_blur));
}

__global__ void __launch_bounds__(256) coarse_points_kernel(const float* __restrict__ points,
                                                            const float* __restrict__ radius, int64_t P,
                                                            const int64_t* __restrict__ first,
                                                            const int64_t* __restrict__ num, const CoarseParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int n = find_owner(first, num, p.N, i);
  if (n < 0) return;
  const float x = points[i * 3 + 0], y = points[i * 3 + 1], z = points[i * 3 + 2], r = radius[i];
  if (z < 0.0f) return;  // (:67)
  coarse_insert(p, n, i, fsub(x, r), fadd(x, r), fsub(y, r), fadd(y, r));
}

}  // namespace b200r

using namespace b200r;

static int coarse_prepare(int32_t N, int32_t H, int32_t W, int32_t bin_size, int32_t M, int32_t* bins, int32_t* counts,
                          int32_t* overflow, cudaStream_t stream, CoarseParams& p) {
  if (N < 0 || H < 0 || W < 0 || M < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (bin_size <= 0) return fail(B200R_ERR_INVALID_ARGUMENT, "bin_size must be positive for the coarse stage");
  p.N = N; p.H = H; p.W = W; p.bin_size = bin_size; p.M = M;
  p.BH = 1 + (H - 1) / bin_size; p.BW = 1 + (W - 1) / bin_size;  // (rasterize_coarse.cu:234-236)
  p.rx = ndc_range(W, H); p.ry = ndc_range(H, W);
  p.bins = bins; p.counts = counts; p.overflow = overflow;
  const size_t nb = (size_t)N * p.BH * p.BW;
  if (nb * (size_t)M > 0) B200R_CUDA_OK(cudaMemsetAsync(bins, 0xFF, sizeof(int32_t) * nb * (size_t)M, stream));
  if (nb > 0) B200R_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * nb, stream));
  B200R_CUDA_OK(cudaMemsetAsync(overflow, 0, sizeof(int32_t), stream));
  return B200R_OK;
}

extern "C" int b200r_rasterize_meshes_coarse(const float* face_verts, int64_t F, const int64_t* first,
                                             const int64_t* num, int32_t N, int32_t H, int32_t W, float blur_radius,
                                             int32_t bin_size, int32_t max_faces_per_bin, int32_t* bin_faces,
                                             int32_t* bin_counts, int32_t* overflow, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CoarseParams p;
  const int rc = coarse_prepare(N, H, W, bin_size, max_faces_per_bin, bin_faces, bin_counts, overflow, stream, p);
  if (rc != B200R_OK) return rc;
  if (F <= 0 || (int64_t)N * H * W == 0) return B200R_OK;
  coarse_faces_kernel<<<(unsigned)((F + 255) / 256), 256, 0, stream>>>(face_verts, F, first, num, sqrtf(blur_radius), p);
  B200R_LAUNCHED("coarse_faces_kernel");
  return B200R_OK;
}

extern "C" int b200r_rasterize_points_coarse(const float* points, int64_t P, const int64_t* first, const int64_t* num,
                                             const float* radius, int32_t N, int32_t H, int32_t W, int32_t bin_size,
                                             int32_t max_points_per_bin, int32_t* bin_points, int32_t* bin_counts,
                                             int32_t* overflow, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CoarseParams p;
  const int rc = coarse_prepare(N, H, W, bin_size, max_points_per_bin, bin_points, bin_counts, overflow, stream, p);
  if (rc != B200R_OK) return rc;
  if (P <= 0 || (int64_t)N * H * W == 0) return B200R_OK;
  coarse_points_kernel<<<(unsigned)((P + 255) / 256), 256, 0, stream>>>(points, radius, P, first, num, p);
  B200R_LAUNCHED("coarse_points_kernel");
  return B200R_OK;
}
