// Exact-rounding arithmetic of the rasterizer hot path.
//
// Every operation here is written with explicit round-to-nearest intrinsics so that the
// compiler can neither fuse nor reassociate anything: the FMA placement below is the one
// nvcc 12.9 (-arch=sm_100a, default -fmad=true) produces for the reference's
// pytorch3d/csrc/utils/geometry_utils.cuh, which is what makes pix_to_face bit-identical
// to the reference CUDA kernels (see DESIGN.md "Arithmetic spec"; verified on B200 against
// oracle/_ref/ref_raster_cuda.so by tests/test_gpu_parity.py).
//
// Reference lines restated (nothing is copied; these are re-derivations):
//   pix_to_ndc          rasterize_points/rasterization_utils.cuh:15-41
//   edge_fn             utils/geometry_utils.cuh:37-40
//   BaryCoords          utils/geometry_utils.cuh:76-86
//   persp correction    utils/geometry_utils.cuh:172-185
//   bary clip           utils/geometry_utils.cuh:246-259
//   point-line dist     utils/geometry_utils.cuh:340-352
//   point-triangle dist utils/geometry_utils.cuh:397-408
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200r {

constexpr double kEps = 1e-8;  // a double in the reference (geometry_utils.cuh:18)

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__host__ __device__ __forceinline__ float ndc_range(int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)S1 * range) / (float)S2;  // rn(rn(S1*2)/S2); S1*2 is exact
  return range;
}

// NDC coordinate of the centre of pixel i (0 <= i < S1) along an axis of S1 pixels.
__device__ __forceinline__ float pix_to_ndc(int i, int S1, float range) {
  const float offset = fmul(range, 0.5f);
  return fsub(fdiv(ffma(range, (float)i, offset), (float)S1), offset);
}

// (p.x-a.x)*(b.y-a.y) - (p.y-a.y)*(b.x-a.x): first product fused, second rounded.
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  const float t = fmul(fsub(py, ay), fsub(bx, ax));
  return ffma(fsub(px, ax), fsub(by, ay), -t);
}

// a.x*b.x + a.y*b.y  ->  fma(a.x, b.x, rn(a.y*b.y))
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) {
  return ffma(ax, bx, fmul(ay, by));
}

// d.x*d.x + d.y*d.y of the final squared distances  ->  fma(d.y, d.y, rn(d.x*d.x))
// (the reference build fuses the other product here; read from its SASS)
__device__ __forceinline__ float sqnorm2(float dx, float dy) { return ffma(dy, dy, fmul(dx, dx)); }

struct Face {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

// float( double(E(v2,v0,v1)) + 1e-8 )
__device__ __forceinline__ float bary_denominator(const Face& f) {
  const float e = edge_fn(f.x2, f.y2, f.x0, f.y0, f.x1, f.y1);
  return __double2float_rn(__dadd_rn((double)e, kEps));
}

__device__ __forceinline__ void bary_coords(float px, float py, const Face& f, float den, float& w0, float& w1,
                                            float& w2) {
  w0 = fdiv(edge_fn(px, py, f.x1, f.y1, f.x2, f.y2), den);
  w1 = fdiv(edge_fn(px, py, f.x2, f.y2, f.x0, f.y0), den);
  w2 = fdiv(edge_fn(px, py, f.x0, f.y0, f.x1, f.y1), den);
}

__device__ __forceinline__ void bary_persp(float& w0, float& w1, float& w2, float z0, float z1, float z2) {
  const float t0 = fmul(fmul(w0, z1), z2);
  const float t1 = fmul(fmul(z0, w1), z2);
  const float t2 = fmul(fmul(z0, z1), w2);
  const float den = fmaxf(fadd(fadd(t0, t1), t2), 1e-8f);
  w0 = fdiv(t0, den);
  w1 = fdiv(t1, den);
  w2 = fdiv(t2, den);
}

__device__ __forceinline__ void bary_clip(float& w0, float& w1, float& w2) {
  const float c0 = fmaxf(w0, 0.0f), c1 = fmaxf(w1, 0.0f), c2 = fmaxf(w2, 0.0f);
  const float s = fmaxf(fadd(fadd(c0, c1), c2), 1e-5f);
  w0 = fdiv(c0, s);
  w1 = fdiv(c1, s);
  w2 = fdiv(c2, s);
}

__device__ __forceinline__ float point_line_dist(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = fsub(bx, ax), bay = fsub(by, ay);
  const float l2 = dot2(bax, bay, bax, bay);
  // the reference compares in double: (double)l2 <= 1e-8.  For a float l2 that is the same as l2 <= 1e-8f, because
  // 1e-8f (9.99999994e-9) is the largest float that is <= 1e-8 -- one FSETP instead of a conversion and a DSETP
  static_assert((double)1e-8f <= 1e-8, "1e-8f must round below 1e-8");
  if (l2 <= 1e-8f) {
    const float dx = fsub(px, bx), dy = fsub(py, by);
    return sqnorm2(dx, dy);
  }
  float t = fdiv(dot2(bax, bay, fsub(px, ax), fsub(py, ay)), l2);
  t = __saturatef(t);
  const float qx = ffma(t, bax, ax), qy = ffma(t, bay, ay);
  const float dx = fsub(qx, px), dy = fsub(qy, py);
  return sqnorm2(dx, dy);
}

__device__ __forceinline__ float point_tri_dist(float px, float py, const Face& f) {
  const float e01 = point_line_dist(px, py, f.x0, f.y0, f.x1, f.y1);
  const float e02 = point_line_dist(px, py, f.x0, f.y0, f.x2, f.y2);
  const float e12 = point_line_dist(px, py, f.x1, f.y1, f.x2, f.y2);
  return fminf(fminf(e01, e02), e12);
}

}  // namespace b200r
