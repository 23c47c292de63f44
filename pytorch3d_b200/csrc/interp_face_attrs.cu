// Barycentric interpolation of per-face vertex attributes at the rasterized pixels (forward + backward).
//
// "Next" row f3, the first consumer of Fragments in every shader: replaces
//   InterpFaceAttrsForwardKernel / InterpFaceAttrsBackwardKernel
//   (pytorch3d/csrc/interp_face_attrs/interp_face_attrs.cu:15-49, 86-124) behind
//   pytorch3d._C.interp_face_attrs_forward / _backward.
//
//   pix_attrs[p, d] = sum_i bary[p, i] * face_attrs[pix_to_face[p], i, d]        (0 where pix_to_face[p] < 0)
//
// Forward: one thread per (pixel-slot, attribute) like the reference (coalesced writes, broadcast index loads),
// the same FMA chain as the reference compiles to (fma(w2,a2, fma(w1,a1, fma(w0,a0,0)))) => bit-identical, and
// the zero of empty slots is written by the kernel (no at::zeros pre-pass).
// Backward: one thread per pixel-slot: grad_bary is a plain dot product over the attributes, written once (the
// reference issues 3*D atomics per slot for it); only grad_face_attrs -- a genuine scatter -- uses atomics.
#include "common.cuh"
#include "raster_math.cuh"

namespace b200r {

__global__ void __launch_bounds__(256)
    interp_face_attrs_forward_kernel(const int64_t* __restrict__ pix_to_face, const float* __restrict__ bary,
                                     const float* __restrict__ face_attrs, float* __restrict__ pix_attrs, int64_t P,
                                     int64_t D) {
  const int64_t total = P * D;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pd < total; pd += stride) {
    const int64_t p = pd / D, d = pd - p * D;
    const int64_t f = __ldg(pix_to_face + p);
    float v = 0.0f;
    if (f >= 0) {
      const float* a = face_attrs + f * 3 * D + d;
      v = ffma(__ldg(bary + p * 3 + 0), __ldg(a), 0.0f);
      v = ffma(__ldg(bary + p * 3 + 1), __ldg(a + D), v);
      v = ffma(__ldg(bary + p * 3 + 2), __ldg(a + 2 * D), v);
    }
    pix_attrs[pd] = v;
  }
}

__global__ void __launch_bounds__(256)
    interp_face_attrs_backward_kernel(const int64_t* __restrict__ pix_to_face, const float* __restrict__ bary,
                                      const float* __restrict__ face_attrs, const float* __restrict__ grad_pix_attrs,
                                      float* __restrict__ grad_bary, float* __restrict__ grad_face_attrs, int64_t P,
                                      int64_t D) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
    const int64_t f = __ldg(pix_to_face + p);
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    if (f >= 0) {
      const float w0 = __ldg(bary + p * 3 + 0), w1 = __ldg(bary + p * 3 + 1), w2 = __ldg(bary + p * 3 + 2);
      const float* a = face_attrs + f * 3 * D;
      float* ga = grad_face_attrs + f * 3 * D;
      const float* up = grad_pix_attrs + p * D;
      for (int64_t d = 0; d < D; ++d) {
        const float u = __ldg(up + d);
        g0 += __ldg(a + d) * u;          // grad_bary_down = vert_attr * upstream_grad   (:111)
        g1 += __ldg(a + D + d) * u;
        g2 += __ldg(a + 2 * D + d) * u;
        atomicAdd(ga + d, w0 * u);        // grad_face_down = weight * upstream_grad     (:112, :114)
        atomicAdd(ga + D + d, w1 * u);
        atomicAdd(ga + 2 * D + d, w2 * u);
      }
    }
    grad_bary[p * 3 + 0] = g0;
    grad_bary[p * 3 + 1] = g1;
    grad_bary[p * 3 + 2] = g2;
  }
}

}  // namespace b200r

using namespace b200r;

extern "C" int b200r_interp_face_attrs_forward(const int64_t* pix_to_face, const float* barycentric_coords,
                                               const float* face_attrs, int64_t P, int64_t F, int64_t D,
                                               float* pix_attrs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)F;
  if (P < 0 || F < 0 || D < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (P * D == 0) return B200R_OK;
  int64_t blocks = (P * D + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  interp_face_attrs_forward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pix_to_face, barycentric_coords, face_attrs,
                                                                       pix_attrs, P, D);
  B200R_LAUNCHED("interp_face_attrs_forward_kernel");
  return B200R_OK;
}

extern "C" int b200r_interp_face_attrs_backward(const int64_t* pix_to_face, const float* barycentric_coords,
                                                const float* face_attrs, const float* grad_pix_attrs, int64_t P,
                                                int64_t F, int64_t D, float* grad_barycentric_coords,
                                                float* grad_face_attrs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P < 0 || F < 0 || D < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (F * D > 0) B200R_CUDA_OK(cudaMemsetAsync(grad_face_attrs, 0, sizeof(float) * (size_t)(F * 3 * D), stream));
  if (P == 0) return B200R_OK;
  int64_t blocks = (P + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  interp_face_attrs_backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      pix_to_face, barycentric_coords, face_attrs, grad_pix_attrs, grad_barycentric_coords, grad_face_attrs, P, D);
  B200R_LAUNCHED("interp_face_attrs_backward_kernel");
  return B200R_OK;
}
