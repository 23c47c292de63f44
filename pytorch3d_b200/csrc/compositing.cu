// Alpha compositing of point features along the z-sorted hits of each pixel (forward + backward).
//
// First "next" row after the rasterizer itself (BASELINE config 3 names it): replaces
//   alphaCompositeCudaForwardKernel / alphaCompositeCudaBackwardKernel
//   (pytorch3d/csrc/compositing/alpha_composite.cu:24-70, 72-139) behind pytorch3d._C.accum_alphacomposite[_backward].
//
//   result[n,c,y,x]   = sum_k  feat[c, idx[n,k,y,x]] * cum_k * alpha[n,k,y,x],   cum_k = prod_{l<k, valid} (1 - alpha_l)
//
// Redesign: the reference runs one thread per (pixel, channel), recomputes the transmittance chain per channel,
// accumulates with atomics into a pre-zeroed result, and its backward issues O(K^2) atomics per (pixel, channel)
// on grad_alphas.  Here one thread owns a pixel: the chain is walked once, every output is written exactly once
// (no zero-fill, no atomics on result / grad_alphas), grad_alphas uses a suffix sum (O(K) per pixel), and only
// grad_features -- a genuine scatter -- uses atomics.  The forward value is bit-identical to the reference kernel:
// same products ((f * cum) * alpha), same ascending-k summation order.
// `alphas` / `points_idx` are addressed through element strides, because the renderer passes permuted views of the
// rasterizer's (N,H,W,K) outputs (pytorch3d/renderer/points/renderer.py:65-73); no copy is needed.
#include "common.cuh"
#include "raster_math.cuh"

namespace b200r {

struct Strides4 {
  int64_t n, k, y, x;
};

constexpr float kCompEps = 1e-9f;  // alpha_composite.cu:20

__global__ void __launch_bounds__(256)
    alpha_composite_forward_kernel(const float* __restrict__ features, int64_t C, int64_t P,
                                   const float* __restrict__ alphas, Strides4 sa,
                                   const int64_t* __restrict__ points_idx, Strides4 si, int N, int K, int H, int W,
                                   float* __restrict__ result) {
  const int64_t total = (int64_t)N * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += stride) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    const float* ap = alphas + n * sa.n + y * sa.y + x * sa.x;
    const int64_t* ip = points_idx + n * si.n + y * si.y + x * si.x;
    for (int64_t c = 0; c < C; ++c) {
      const float* fc = features + c * P;
      float acc = 0.0f, cum = 1.0f;
      for (int k = 0; k < K; ++k) {
        const int64_t id = ip[k * si.k];
        if (id < 0) continue;  // -1: no point overlaps the pixel in this slot (:54-57)
        const float a = ap[k * sa.k];
        acc = fadd(acc, fmul(fmul(__ldg(fc + id), cum), a));  // (:63-64): features * cum_alpha * alpha
        cum = fmul(cum, fsub(1.0f, a));
      }
      result[(((int64_t)n * C + c) * H + y) * W + x] = acc;
    }
  }
}

__global__ void __launch_bounds__(256)
    alpha_composite_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ features, int64_t C,
                                    int64_t P, const float* __restrict__ alphas, Strides4 sa,
                                    const int64_t* __restrict__ points_idx, Strides4 si, int N, int K, int H, int W,
                                    float* __restrict__ grad_features, float* __restrict__ grad_alphas) {
  const int64_t total = (int64_t)N * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t plane = (int64_t)H * W;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += stride) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / plane);
    const float* ap = alphas + n * sa.n + y * sa.y + x * sa.x;
    const int64_t* ip = points_idx + n * si.n + y * si.y + x * si.x;
    const float* go = grad_out + (int64_t)n * C * plane + (int64_t)y * W + x;  // + c * plane
    float* ga = grad_alphas + (int64_t)n * K * plane + (int64_t)y * W + x;     // + k * plane (contiguous N,K,H,W)
    // transmittance before the last valid slot, then walk the slots backwards keeping the suffix sum
    //   S_k = sum_{t>k} cum_t * alpha_t * A_t,   A_t = sum_c grad_out_c * feat[c, idx_t]
    // grad_alpha_k = cum_k * A_k - S_k / (1 - alpha_k + eps)          (alpha_composite.cu:112-134, summed over c)
    float cum = 1.0f;
    for (int k = 0; k < K; ++k) {
      if (ip[k * si.k] >= 0) cum *= 1.0f - ap[k * sa.k];
    }
    float suffix = 0.0f;
    for (int k = K - 1; k >= 0; --k) {
      const int64_t id = ip[k * si.k];
      if (id < 0) {
        ga[k * plane] = 0.0f;
        continue;
      }
      const float a = ap[k * sa.k];
      const float one_minus = 1.0f - a;
      // cum currently includes slot k: undo it (exactly what the forward chain had before slot k, up to rounding;
      // recomputed from scratch when the factor is ~0 to avoid dividing by it)
      float cum_k;
      if (fabsf(one_minus) > 1e-6f) {
        cum_k = cum / one_minus;
      } else {
        cum_k = 1.0f;
        for (int l = 0; l < k; ++l)
          if (ip[l * si.k] >= 0) cum_k *= 1.0f - ap[l * sa.k];
      }
      float A = 0.0f;
      const float w = cum_k * a;
      for (int64_t c = 0; c < C; ++c) {
        const float g = go[c * plane];
        A += g * __ldg(features + c * P + id);
        atomicAdd(grad_features + c * P + id, g * w);  // (:115-117)
      }
      ga[k * plane] = cum_k * A - suffix / (one_minus + kCompEps);
      suffix += w * A;
      cum = cum_k;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Weighted sum and normalised weighted sum of point features: replace
//   weightedSumCudaForwardKernel / BackwardKernel          (pytorch3d/csrc/compositing/weighted_sum.cu:22-61, 63-103)
//   weightedSumNormCudaForwardKernel / BackwardKernel      (norm_weighted_sum.cu:24-82, 84-160)
// behind pytorch3d._C.accum_weightedsum[_backward] / accum_weightedsumnorm[_backward].
//   result[n,c,y,x] = sum_k alpha_k * feat[c, idx_k]  (/ max(sum_k alpha_k, 1e-4) when NORM)
// Same redesign as above: one thread per pixel, every output written once (the reference zero-fills `result` and
// `grad_alphas` and accumulates both with atomics, one thread per (pixel, channel)); only grad_features scatters.
// Forward: the reference's operations in its order ((f * alpha) / total, summed over ascending k) -> identical bits.
// Backward of NORM: grad_alpha_k = sum_c g_c (f_ck S - sum_t alpha_t f_ct) / S^2 = A_k / S - T / S^2 with
// A_k = sum_c g_c f_ck and T = sum_t alpha_t A_t: two passes over the K slots instead of a per-channel rescan.
// ------------------------------------------------------------------------------------------------
constexpr float kNormEps = 1e-4f;  // norm_weighted_sum.cu:20

template <bool NORM>
__global__ void __launch_bounds__(256)
    weighted_sum_forward_kernel(const float* __restrict__ features, int64_t C, int64_t P,
                                const float* __restrict__ alphas, Strides4 sa, const int64_t* __restrict__ points_idx,
                                Strides4 si, int N, int K, int H, int W, float* __restrict__ result) {
  const int64_t total = (int64_t)N * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += stride) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    const float* ap = alphas + n * sa.n + y * sa.y + x * sa.x;
    const int64_t* ip = points_idx + n * si.n + y * si.y + x * si.x;
    float sum_alpha = 1.0f;
    if (NORM) {
      sum_alpha = 0.0f;
      for (int k = 0; k < K; ++k)
        if ((int)ip[k * si.k] >= 0) sum_alpha = fadd(sum_alpha, ap[k * sa.k]);  // (:54-62; index read into an int)
      if (sum_alpha < kNormEps) sum_alpha = kNormEps;
    }
    for (int64_t c = 0; c < C; ++c) {
      const float* fc = features + c * P;
      float acc = 0.0f;
      for (int k = 0; k < K; ++k) {
        const int id = (int)ip[k * si.k];
        if (id < 0) continue;
        const float t = fmul(__ldg(fc + id), ap[k * sa.k]);
        acc = fadd(acc, NORM ? fdiv(t, sum_alpha) : t);  // (norm_weighted_sum.cu:78-79, weighted_sum.cu:58)
      }
      result[(((int64_t)n * C + c) * H + y) * W + x] = acc;
    }
  }
}

template <bool NORM>
__global__ void __launch_bounds__(256)
    weighted_sum_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ features, int64_t C,
                                 int64_t P, const float* __restrict__ alphas, Strides4 sa,
                                 const int64_t* __restrict__ points_idx, Strides4 si, int N, int K, int H, int W,
                                 float* __restrict__ grad_features, float* __restrict__ grad_alphas) {
  const int64_t total = (int64_t)N * H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t plane = (int64_t)H * W;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += stride) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / plane);
    const float* ap = alphas + n * sa.n + y * sa.y + x * sa.x;
    const int64_t* ip = points_idx + n * si.n + y * si.y + x * si.x;
    const float* go = grad_out + (int64_t)n * C * plane + (int64_t)y * W + x;  // + c * plane
    float* ga = grad_alphas + (int64_t)n * K * plane + (int64_t)y * W + x;     // + k * plane (contiguous N,K,H,W)
    float S = 1.0f;
    if (NORM) {
      S = 0.0f;
      for (int k = 0; k < K; ++k)
        if ((int)ip[k * si.k] >= 0) S += ap[k * sa.k];
      if (S < kNormEps) S = kNormEps;
    }
    const float inv_s = 1.0f / S;
    float T = 0.0f;
    for (int k = 0; k < K; ++k) {
      const int id = (int)ip[k * si.k];
      if (id < 0) {
        ga[k * plane] = 0.0f;
        continue;
      }
      const float a = ap[k * sa.k];
      float A = 0.0f;
      for (int64_t c = 0; c < C; ++c) {
        const float g = go[c * plane];
        A += g * __ldg(features + c * P + id);
        atomicAdd(grad_features + c * P + id, NORM ? a * g * inv_s : a * g);  // (weighted_sum.cu:99-100, norm:155-157)
      }
      ga[k * plane] = NORM ? A * inv_s : A;
      T += a * A;
    }
    if (NORM) {
      const float corr = T * inv_s * inv_s;
      for (int k = 0; k < K; ++k)
        if ((int)ip[k * si.k] >= 0) ga[k * plane] -= corr;
    }
  }
}

}  // namespace b200r

using namespace b200r;

static int check_comp_args(int64_t C, int64_t P, int32_t N, int32_t K, int32_t H, int32_t W) {
  if (C < 0 || P < 0 || N < 0 || K < 0 || H < 0 || W < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  return B200R_OK;
}

extern "C" int b200r_alpha_composite_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                             const int64_t* alpha_strides, const int64_t* points_idx,
                                             const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                             float* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_comp_args(C, P, N, K, H, W);
  if (rc != B200R_OK) return rc;
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return B200R_OK;
  const Strides4 sa = {alpha_strides[0], alpha_strides[1], alpha_strides[2], alpha_strides[3]};
  const Strides4 si = {idx_strides[0], idx_strides[1], idx_strides[2], idx_strides[3]};
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  alpha_composite_forward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(features, C, P, alphas, sa, points_idx, si, N, K,
                                                                     H, W, result);
  B200R_LAUNCHED("alpha_composite_forward_kernel");
  return B200R_OK;
}

extern "C" int b200r_alpha_composite_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                              const float* alphas, const int64_t* alpha_strides,
                                              const int64_t* points_idx, const int64_t* idx_strides, int32_t N,
                                              int32_t K, int32_t H, int32_t W, float* grad_features,
                                              float* grad_alphas, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_comp_args(C, P, N, K, H, W);
  if (rc != B200R_OK) return rc;
  if (C * P > 0) B200R_CUDA_OK(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)(C * P), stream));
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || K == 0) return B200R_OK;
  const Strides4 sa = {alpha_strides[0], alpha_strides[1], alpha_strides[2], alpha_strides[3]};
  const Strides4 si = {idx_strides[0], idx_strides[1], idx_strides[2], idx_strides[3]};
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  alpha_composite_backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(grad_out, features, C, P, alphas, sa, points_idx,
                                                                      si, N, K, H, W, grad_features, grad_alphas);
  B200R_LAUNCHED("alpha_composite_backward_kernel");
  return B200R_OK;
}

template <bool NORM>
static int weighted_sum_forward_impl(const float* features, int64_t C, int64_t P, const float* alphas,
                                     const int64_t* alpha_strides, const int64_t* points_idx,
                                     const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                     float* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_comp_args(C, P, N, K, H, W);
  if (rc != B200R_OK) return rc;
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return B200R_OK;
  const Strides4 sa = {alpha_strides[0], alpha_strides[1], alpha_strides[2], alpha_strides[3]};
  const Strides4 si = {idx_strides[0], idx_strides[1], idx_strides[2], idx_strides[3]};
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  weighted_sum_forward_kernel<NORM><<<(unsigned)blocks, 256, 0, stream>>>(features, C, P, alphas, sa, points_idx, si, N,
                                                                        K, H, W, result);
  B200R_LAUNCHED("weighted_sum_forward_kernel");
  return B200R_OK;
}

template <bool NORM>
static int weighted_sum_backward_impl(const float* grad_out, const float* features, int64_t C, int64_t P,
                                      const float* alphas, const int64_t* alpha_strides, const int64_t* points_idx,
                                      const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                      float* grad_features, float* grad_alphas, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_comp_args(C, P, N, K, H, W);
  if (rc != B200R_OK) return rc;
  if (C * P > 0) B200R_CUDA_OK(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)(C * P), stream));
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || K == 0) return B200R_OK;
  const Strides4 sa = {alpha_strides[0], alpha_strides[1], alpha_strides[2], alpha_strides[3]};
  const Strides4 si = {idx_strides[0], idx_strides[1], idx_strides[2], idx_strides[3]};
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  weighted_sum_backward_kernel<NORM><<<(unsigned)blocks, 256, 0, stream>>>(
      grad_out, features, C, P, alphas, sa, points_idx, si, N, K, H, W, grad_features, grad_alphas);
  B200R_LAUNCHED("weighted_sum_backward_kernel");
  return B200R_OK;
}

extern "C" int b200r_weighted_sum_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                          const int64_t* alpha_strides, const int64_t* points_idx,
                                          const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                          float* result, void* stream) {
  return weighted_sum_forward_impl<false>(features, C, P, alphas, alpha_strides, points_idx, idx_strides, N, K, H, W,
                                          result, stream);
}
extern "C" int b200r_weighted_sum_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                           const float* alphas, const int64_t* alpha_strides,
                                           const int64_t* points_idx, const int64_t* idx_strides, int32_t N, int32_t K,
                                           int32_t H, int32_t W, float* grad_features, float* grad_alphas,
                                           void* stream) {
  return weighted_sum_backward_impl<false>(grad_out, features, C, P, alphas, alpha_strides, points_idx, idx_strides, N,
                                           K, H, W, grad_features, grad_alphas, stream);
}
extern "C" int b200r_norm_weighted_sum_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                               const int64_t* alpha_strides, const int64_t* points_idx,
                                               const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                               float* result, void* stream) {
  return weighted_sum_forward_impl<true>(features, C, P, alphas, alpha_strides, points_idx, idx_strides, N, K, H, W,
                                         result, stream);
}
extern "C" int b200r_norm_weighted_sum_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                                const float* alphas, const int64_t* alpha_strides,
                                                const int64_t* points_idx, const int64_t* idx_strides, int32_t N,
                                                int32_t K, int32_t H, int32_t W, float* grad_features,
                                                float* grad_alphas, void* stream) {
  return weighted_sum_backward_impl<true>(grad_out, features, C, P, alphas, alpha_strides, points_idx, idx_strides, N,
                                          K, H, W, grad_features, grad_alphas, stream);
}
