// 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier helpers for sm_100a.
//
// Used by the setup/bin kernels to pull a CTA's contiguous slice of the packed
// (F,3,3) face_verts / (P,3) points array into shared memory with one asynchronous
// bulk transaction instead of 9 stride-36B scalar loads per thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200r {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make the barrier init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Cooperative load of `count` 4-byte words starting at gmem word pointer `src` into smem `dst`.
// Uses one TMA bulk transaction for the 16B-aligned body and plain loads for the ragged head/tail.
// All threads of the CTA must call it; contains __syncthreads().  `bar` must have been initialised
// (mbar_init(bar, 1) + fence_mbar_init() + __syncthreads()) and `phase` is its current parity.
__device__ __forceinline__ void cta_load_words(float* dst, const float* src, int count, uint64_t* bar,
                                               uint32_t phase) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  // words to skip until src is 16B aligned (src is always 4B aligned)
  int head = static_cast<int>(((16 - (a & 15)) & 15) >> 2);
  if (head > count) head = count;
  // dst must be aligned the same way as src for the bulk part: caller passes dst with
  // (dst_word_index % 4) == (src_word_index % 4); we simply offset both by `head`.
  const int body = ((count - head) >> 2) << 2;
  const bool use_tma = body > 0 && ((reinterpret_cast<uintptr_t>(dst + head) & 15) == 0);
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, static_cast<uint32_t>(body) * 4u);
      bulk_g2s(dst + head, src + head, static_cast<uint32_t>(body) * 4u, bar);
    }
    for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    for (int i = head + body + threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
    mbar_wait(bar, phase);
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
}

}  // namespace b200r
