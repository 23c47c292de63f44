// Frame exchange between the GPUs of one node: the path's only collective ("NCCL over NVLink only to
// gather rendered frames"), written as ONE kernel that packs this rank's Fragments and pushes them straight into the
// memory of every peer over NVLink (plain st.global on CUDA-IPC-mapped peer pointers), plus the kernel that expands a
// received stream back into dense (N, H, W, K) buffers.
//
// Why pack: Fragments are mostly padding.  Every pixel holds `cnt` valid slots followed by K - cnt slots of -1 (the
// rasterizers write valid slots first, in depth order); on the north-star batch 12 % of the slots are valid.  A dense
// all-gather moves 24-28 B for every slot -- at 8 GPUs 2.8 GB into every rank per step, 10x the time of rasterizing --
// while the packed stream is 1 B per pixel (cnt) + 24 B per VALID slot (face i32, z, signed dist, 3 barycentrics):
// lossless, and the -1 padding is regenerated on arrival.
//
// Stream layout (one region per source rank in the receiver's arena; all offsets 16-byte aligned):
//   seg_offset int32 [nseg]        first payload entry of each segment of 128 consecutive pixels
//   counts     uint8 [npix]        valid slots per pixel
//   payload    6 x 4 B [entries]   (face, z, dist, b0, b1, b2) of the valid slots, segment by segment (segments are
//                                  placed by an atomic cursor, i.e. in arbitrary order; seg_offset finds them)
// The pack kernel stages a segment's payload in shared memory and copies it out with coalesced 8-byte stores, once per
// destination: every NVLink write is a contiguous run of a few KB.
#include <climits>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b200r {

constexpr int SEG = 128;  // pixels per segment = threads per CTA of the pack / unpack kernels
constexpr int MAX_PEERS = 16;
constexpr int PACK_MAX_K = 32;

struct PackedLayout {
  size_t off_seg, off_counts, off_payload, bytes;
  int64_t npix, nseg;
};

static inline PackedLayout packed_layout(int64_t n_images, int H, int W, int K) {
  PackedLayout l;
  l.npix = n_images * (int64_t)H * W;
  l.nseg = (l.npix + SEG - 1) / SEG;
  size_t off = 0;
  l.off_seg = off;
  off = align_up(off + sizeof(int32_t) * (size_t)(l.nseg > 0 ? l.nseg : 1), 16);
  l.off_counts = off;
  off = align_up(off + (size_t)(l.npix > 0 ? l.npix : 1), 16);
  l.off_payload = off;
  off = align_up(off + 24 * (size_t)(l.npix > 0 ? l.npix : 1) * (size_t)(K > 0 ? K : 1), 16);
  l.bytes = off;
  return l;
}

struct PeerPtrs {
  unsigned char* p[MAX_PEERS];
  int n;
};

// exclusive scan of one int per thread over the CTA's 128 threads; returns the exclusive prefix, total via `total`
__device__ __forceinline__ int cta_scan128(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int i = 0; i < SEG / 32; ++i) before += i < w ? warp_sums[i] : 0;
  total = warp_sums[0] + warp_sums[1] + warp_sums[2] + warp_sums[3];
  return before + inc - v;
}

__global__ void __launch_bounds__(SEG)
    fragments_pack_push_kernel(const int64_t* __restrict__ pix_to_face, const float* __restrict__ zbuf,
                               const float* __restrict__ bary, const float* __restrict__ dists, int64_t npix, int K,
                               PackedLayout lay, PeerPtrs dst, int* __restrict__ cursor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint2* stage = reinterpret_cast<uint2*>(smem_raw);  // SEG * K entries of 3 x uint2
  __shared__ int warp_sums[SEG / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x;
  const int64_t seg = blockIdx.x;
  const int64_t pix = seg * SEG + tid;
  int cnt = 0;
  if (pix < npix) {
    const int64_t* f = pix_to_face + pix * K;  // valid slots come first: their number = the number of slots >= 0
    if ((K & 1) == 0) {                        // independent 16-byte loads instead of a chain of dependent ones
      for (int k = 0; k < K; k += 2) {
        const longlong2 v = __ldcs(reinterpret_cast<const longlong2*>(f + k));
        cnt += (v.x >= 0) + (v.y >= 0);
      }
    } else {
      for (int k = 0; k < K; ++k) cnt += __ldcs(f + k) >= 0;
    }
  }
  int total;
  const int off = cta_scan128(cnt, warp_sums, total);
  if (tid == 0) s_base = total > 0 ? atomicAdd(cursor, total) : 0;
  for (int k = 0; k < cnt; ++k) {
    const int64_t i = pix * K + k;
    uint2* e = stage + (size_t)(off + k) * 3;
    e[0] = make_uint2((unsigned)(int)__ldcs(pix_to_face + i), __float_as_uint(__ldcs(zbuf + i)));
    e[1] = make_uint2(__float_as_uint(__ldcs(dists + i)), __float_as_uint(__ldcs(bary + i * 3)));
    e[2] = make_uint2(__float_as_uint(__ldcs(bary + i * 3 + 1)), __float_as_uint(__ldcs(bary + i * 3 + 2)));
  }
  __syncthreads();
  const int base = s_base;
  for (int d = 0; d < dst.n; ++d) {
    unsigned char* r = dst.p[d];
    if (pix < npix) r[lay.off_counts + pix] = (unsigned char)cnt;
    if (tid == 0) reinterpret_cast<int*>(r + lay.off_seg)[seg] = base;
    uint2* out = reinterpret_cast<uint2*>(r + lay.off_payload) + (size_t)base * 3;
    for (int i = tid; i < total * 3; i += SEG) out[i] = stage[i];
  }
}

// Expand one received stream into the dense full-batch buffers.  `image_index[j]` = position of the source's j-th
// image in the full batch, `face_shift[j]` = what turns its local packed face ids into global ones.
__global__ void __launch_bounds__(SEG)
    fragments_unpack_kernel(const unsigned char* __restrict__ region, PackedLayout lay, int64_t npix, int HW, int K,
                            const int32_t* __restrict__ image_index, const int64_t* __restrict__ face_shift,
                            int64_t* __restrict__ pix_to_face, float* __restrict__ zbuf, float* __restrict__ bary,
                            float* __restrict__ dists) {
  __shared__ int warp_sums[SEG / 32];
  const int tid = threadIdx.x;
  const int64_t seg = blockIdx.x;
  const int64_t pix = seg * SEG + tid;
  int cnt = 0;
  if (pix < npix) cnt = region[lay.off_counts + pix];
  int total;
  const int off = cta_scan128(cnt, warp_sums, total);
  if (pix >= npix) return;
  const int base = reinterpret_cast<const int*>(region + lay.off_seg)[seg];
  const uint2* e = reinterpret_cast<const uint2*>(region + lay.off_payload) + (size_t)(base + off) * 3;
  const int j = (int)(pix / HW);
  const int64_t o = ((int64_t)image_index[j] * HW + (pix - (int64_t)j * HW)) * K;
  const int64_t shift = face_shift[j];
  if ((K & 3) == 0) {
    // 4 slots at a time: 16-byte stores (32 for pix_to_face)
    for (int k0 = 0; k0 < K; k0 += 4) {
      long long id[4];
      float z[4], dd[4], b[12];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        id[u] = -1;
        z[u] = dd[u] = b[3 * u] = b[3 * u + 1] = b[3 * u + 2] = -1.0f;
        if (k0 + u < cnt) {
          const uint2 a = e[(k0 + u) * 3], c = e[(k0 + u) * 3 + 1], g = e[(k0 + u) * 3 + 2];
          id[u] = (long long)(int)a.x + shift;
          z[u] = __uint_as_float(a.y);
          dd[u] = __uint_as_float(c.x);
          b[3 * u] = __uint_as_float(c.y);
          b[3 * u + 1] = __uint_as_float(g.x);
          b[3 * u + 2] = __uint_as_float(g.y);
        }
      }
      __stcs(reinterpret_cast<longlong2*>(pix_to_face + o + k0), make_longlong2(id[0], id[1]));
      __stcs(reinterpret_cast<longlong2*>(pix_to_face + o + k0 + 2), make_longlong2(id[2], id[3]));
      __stcs(reinterpret_cast<float4*>(zbuf + o + k0), make_float4(z[0], z[1], z[2], z[3]));
      __stcs(reinterpret_cast<float4*>(dists + o + k0), make_float4(dd[0], dd[1], dd[2], dd[3]));
#pragma unroll
      for (int u = 0; u < 3; ++u)
        __stcs(reinterpret_cast<float4*>(bary + (o + k0) * 3) + u,
               make_float4(b[4 * u], b[4 * u + 1], b[4 * u + 2], b[4 * u + 3]));
    }
    return;
  }
  for (int k = 0; k < K; ++k) {
    long long id = -1;
    float z = -1.0f, dd = -1.0f, b0 = -1.0f, b1 = -1.0f, b2 = -1.0f;
    if (k < cnt) {
      const uint2 a = e[k * 3], c = e[k * 3 + 1], g = e[k * 3 + 2];
      id = (long long)(int)a.x + shift;
      z = __uint_as_float(a.y);
      dd = __uint_as_float(c.x);
      b0 = __uint_as_float(c.y);
      b1 = __uint_as_float(g.x);
      b2 = __uint_as_float(g.y);
    }
    pix_to_face[o + k] = id;
    zbuf[o + k] = z;
    dists[o + k] = dd;
    bary[(o + k) * 3] = b0;
    bary[(o + k) * 3 + 1] = b1;
    bary[(o + k) * 3 + 2] = b2;
  }
}

// The same for K % 4 == 0 with whole-sector stores: every thread first lays its pixel's K dense slots (valid entries,
// then the -1 padding) out in shared memory; the CTA then copies the four buffers out in 16-byte pieces, consecutive
// threads writing consecutive pieces -- a segment's 128 pixels are contiguous in the outputs, so every store
// instruction fills whole 32-byte sectors (the per-pixel stores of the kernel above fill half of each sector they
// touch, and the expansion is pure store bandwidth: 28 B per slot of the whole batch).
__global__ void __launch_bounds__(SEG)
    fragments_unpack_staged_kernel(const unsigned char* __restrict__ region, PackedLayout lay, int64_t npix, int HW,
                                   int K, const int32_t* __restrict__ image_index,
                                   const int64_t* __restrict__ face_shift, int64_t* __restrict__ pix_to_face,
                                   float* __restrict__ zbuf, float* __restrict__ bary, float* __restrict__ dists) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  long long* s_id = reinterpret_cast<long long*>(smem_raw);        // [SEG * K]
  float* s_z = reinterpret_cast<float*>(s_id + SEG * K);           // [SEG * K]
  float* s_d = s_z + SEG * K;                                      // [SEG * K]
  float* s_b = s_d + SEG * K;                                      // [SEG * K * 3]
  __shared__ long long s_base[SEG];                                // output slot offset of each pixel, -1 = none
  __shared__ int warp_sums[SEG / 32];
  const int tid = threadIdx.x;
  const int64_t seg = blockIdx.x;
  const int64_t pix = seg * SEG + tid;
  int cnt = 0;
  if (pix < npix) cnt = region[lay.off_counts + pix];
  int total;
  const int off = cta_scan128(cnt, warp_sums, total);
  long long obase = -1;
  if (pix < npix) {
    const int base = reinterpret_cast<const int*>(region + lay.off_seg)[seg];
    const uint2* e = reinterpret_cast<const uint2*>(region + lay.off_payload) + (size_t)(base + off) * 3;
    const int j = (int)(pix / HW);
    obase = ((long long)image_index[j] * HW + (pix - (int64_t)j * HW)) * K;
    const long long shift = face_shift[j];
    // four slots at a time, laid out with 16-byte shared-memory stores (7 per group of four slots)
    for (int k0 = 0; k0 < K; k0 += 4) {
      long long id[4];
      float z[4], dd[4], b[12];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        id[u] = -1;
        z[u] = dd[u] = b[3 * u] = b[3 * u + 1] = b[3 * u + 2] = -1.0f;
        if (k0 + u < cnt) {
          const uint2 a = e[(k0 + u) * 3], c = e[(k0 + u) * 3 + 1], g = e[(k0 + u) * 3 + 2];
          id[u] = (long long)(int)a.x + shift;
          z[u] = __uint_as_float(a.y);
          dd[u] = __uint_as_float(c.x);
          b[3 * u] = __uint_as_float(c.y);
          b[3 * u + 1] = __uint_as_float(g.x);
          b[3 * u + 2] = __uint_as_float(g.y);
        }
      }
      const int at = tid * K + k0;
      reinterpret_cast<longlong2*>(s_id + at)[0] = make_longlong2(id[0], id[1]);
      reinterpret_cast<longlong2*>(s_id + at)[1] = make_longlong2(id[2], id[3]);
      *reinterpret_cast<float4*>(s_z + at) = make_float4(z[0], z[1], z[2], z[3]);
      *reinterpret_cast<float4*>(s_d + at) = make_float4(dd[0], dd[1], dd[2], dd[3]);
#pragma unroll
      for (int u = 0; u < 3; ++u)
        reinterpret_cast<float4*>(s_b + at * 3)[u] = make_float4(b[4 * u], b[4 * u + 1], b[4 * u + 2], b[4 * u + 3]);
    }
  }
  s_base[tid] = obase;
  __syncthreads();
  // copy out: `u` 16-byte pieces per pixel of each buffer
  {
    const int u = K / 2;  // pix_to_face: 8 B per slot
    const float4* src = reinterpret_cast<const float4*>(s_id);
    for (int i = tid; i < SEG * u; i += SEG) {
      const long long o = s_base[i / u];
      if (o >= 0) __stcs(reinterpret_cast<float4*>(pix_to_face + o) + i % u, src[i]);
    }
  }
  {
    const int u = K / 4;  // zbuf, dists: 4 B per slot
    const float4 *srcz = reinterpret_cast<const float4*>(s_z), *srcd = reinterpret_cast<const float4*>(s_d);
    for (int i = tid; i < SEG * u; i += SEG) {
      const long long o = s_base[i / u];
      if (o >= 0) {
        __stcs(reinterpret_cast<float4*>(zbuf + o) + i % u, srcz[i]);
        __stcs(reinterpret_cast<float4*>(dists + o) + i % u, srcd[i]);
      }
    }
  }
  {
    const int u = 3 * K / 4;  // barycentrics: 12 B per slot
    const float4* src = reinterpret_cast<const float4*>(s_b);
    for (int i = tid; i < SEG * u; i += SEG) {
      const long long o = s_base[i / u];
      if (o >= 0) __stcs(reinterpret_cast<float4*>(bary + o * 3) + i % u, src[i]);
    }
  }
}

}  // namespace b200r

using namespace b200r;

// ------------------------------------------------------------------------------------------------ peer memory
extern "C" int b200r_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if (ptr == nullptr || handle64 == nullptr) return fail(B200R_ERR_INVALID_ARGUMENT, "null argument");
  void* p = nullptr;
  B200R_CUDA_OK(cudaMalloc(&p, bytes > 0 ? bytes : 16));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return check_cuda(e, "cudaIpcGetMemHandle");
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return B200R_OK;
}

extern "C" int b200r_peer_open(const unsigned char* handle64, void** ptr) {
  if (ptr == nullptr || handle64 == nullptr) return fail(B200R_ERR_INVALID_ARGUMENT, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200R_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200R_OK;
}

extern "C" int b200r_peer_close(void* ptr) {
  if (ptr) B200R_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return B200R_OK;
}

extern "C" int b200r_peer_free(void* ptr) {
  if (ptr) B200R_CUDA_OK(cudaFree(ptr));
  return B200R_OK;
}

// ------------------------------------------------------------------------------------------------ packed frames
extern "C" size_t b200r_packed_frames_bytes(int64_t n_images, int32_t H, int32_t W, int32_t K) {
  if (n_images < 0 || H < 0 || W < 0 || K < 0) return 0;
  return packed_layout(n_images, H, W, K).bytes;
}

static int launch_pack(const int64_t* pix_to_face, const float* zbuf, const float* bary, const float* dists,
                       int32_t n_images, int32_t H, int32_t W, int32_t K, int64_t n_images_layout,
                       void* const* dst_regions, int32_t n_dst, int32_t* cursor, cudaStream_t stream) {
  if (n_images < 0 || H < 0 || W < 0 || K < 0 || n_dst < 0 || n_images_layout < n_images)
    return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (n_dst > MAX_PEERS) return fail(B200R_ERR_INVALID_ARGUMENT, "at most 16 destinations");
  if (K > PACK_MAX_K) return fail(B200R_ERR_INVALID_ARGUMENT, "packed frame exchange supports K <= 32");
  const PackedLayout lay = packed_layout(n_images_layout, H, W, K);
  const int64_t npix = (int64_t)n_images * H * W;
  if ((npix + SEG - 1) / SEG > INT_MAX || npix * K > INT_MAX)
    return fail(B200R_ERR_INVALID_ARGUMENT, "too many slots for one packed stream");
  B200R_CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(int32_t), stream));
  if (npix == 0 || K == 0 || n_dst == 0) return B200R_OK;
  PeerPtrs dst;
  dst.n = n_dst;
  for (int i = 0; i < n_dst; ++i) dst.p[i] = static_cast<unsigned char*>(dst_regions[i]);
  const size_t smem = (size_t)SEG * K * 24;
  static bool configured[64] = {};
  int dev_ = 0;
  B200R_CUDA_OK(cudaGetDevice(&dev_));
  if (dev_ < 0 || dev_ >= 64 || !configured[dev_]) {
    B200R_CUDA_OK(cudaFuncSetAttribute(fragments_pack_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)SEG * PACK_MAX_K * 24)));
    if (dev_ >= 0 && dev_ < 64) configured[dev_] = true;
  }
  const int64_t nseg = (npix + SEG - 1) / SEG;
  fragments_pack_push_kernel<<<(unsigned)nseg, SEG, smem, stream>>>(pix_to_face, zbuf, bary, dists, npix, K, lay, dst,
                                                                    cursor);
  B200R_LAUNCHED("fragments_pack_push_kernel");
  return B200R_OK;
}

extern "C" int b200r_fragments_pack_push(const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                         const float* dists, int32_t n_images, int32_t H, int32_t W, int32_t K,
                                         int64_t n_images_layout, void* const* dst_regions, int32_t n_dst,
                                         int32_t* cursor, void* stream_) {
  return launch_pack(pix_to_face, zbuf, bary, dists, n_images, H, W, K, n_images_layout, dst_regions, n_dst, cursor,
                     static_cast<cudaStream_t>(stream_));
}

static int launch_unpack(const unsigned char* region, const PackedLayout& lay, int64_t npix, int HW, int K,
                         const int32_t* image_index, const int64_t* face_shift, int64_t* pix_to_face, float* zbuf,
                         float* bary, float* dists, cudaStream_t stream) {
  const int64_t nseg = (npix + SEG - 1) / SEG;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pix_to_face) | reinterpret_cast<uintptr_t>(zbuf) |
                         reinterpret_cast<uintptr_t>(bary) | reinterpret_cast<uintptr_t>(dists)) & 15u) == 0;
  if ((K & 3) == 0 && K <= PACK_MAX_K && aligned) {
    const size_t smem = (size_t)SEG * K * 28;
    static bool configured[64] = {};
    int dev_ = 0;
    B200R_CUDA_OK(cudaGetDevice(&dev_));
    if (dev_ < 0 || dev_ >= 64 || !configured[dev_]) {
      B200R_CUDA_OK(cudaFuncSetAttribute(fragments_unpack_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)SEG * PACK_MAX_K * 28)));
      if (dev_ >= 0 && dev_ < 64) configured[dev_] = true;
    }
    fragments_unpack_staged_kernel<<<(unsigned)nseg, SEG, smem, stream>>>(region, lay, npix, HW, K, image_index,
                                                                          face_shift, pix_to_face, zbuf, bary, dists);
  } else {
    fragments_unpack_kernel<<<(unsigned)nseg, SEG, 0, stream>>>(region, lay, npix, HW, K, image_index, face_shift,
                                                                pix_to_face, zbuf, bary, dists);
  }
  B200R_LAUNCHED("fragments_unpack_kernel");
  return B200R_OK;
}

extern "C" int b200r_fragments_unpack(const void* region, int32_t n_images, int32_t H, int32_t W, int32_t K,
                                      int64_t n_images_layout, const int32_t* image_index, const int64_t* face_shift,
                                      int64_t* pix_to_face, float* zbuf, float* bary, float* dists, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_images < 0 || H < 0 || W < 0 || K < 0 || n_images_layout < n_images)
    return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  const PackedLayout lay = packed_layout(n_images_layout, H, W, K);
  const int64_t npix = (int64_t)n_images * H * W;
  if ((int64_t)H * W > INT_MAX) return fail(B200R_ERR_INVALID_ARGUMENT, "image too large");
  if (npix == 0 || K == 0) return B200R_OK;
  return launch_unpack(static_cast<const unsigned char*>(region), lay, npix, H * W, K, image_index, face_shift,
                       pix_to_face, zbuf, bary, dists, stream);
}

// ------------------------------------------------------------------------------------------------ exchange context
// One object per process: the arenas of all ranks, the batch geometry and the CUDA events that order a step's pack
// (compute stream) -> cross-rank barrier (the caller's: e.g. a 4-byte NCCL all-reduce on the side stream) -> expansion
// (side stream), with arenas and result buffers double-buffered by step parity.  Keeps the host cost of a step to three
// calls (see pytorch3d_b200/peer.py for the protocol and the reuse argument).
namespace {
struct Exchange {
  int world = 0, rank = 0, H = 0, W = 0, K = 0;
  int64_t n_layout = 0;
  size_t region_bytes = 0, half_bytes = 0;
  std::vector<unsigned char*> arena;
  std::vector<int> n_images, first_image;
  int32_t* d_image_index = nullptr;
  int64_t* d_face_shift = nullptr;
  int32_t* d_cursor = nullptr;
  cudaEvent_t packed = nullptr, free_ev[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr},
              consumed = nullptr;
  bool have_free[2] = {false, false}, have_consumed = false;
  int64_t step = 0;
  unsigned char* region(int holder, int parity, int source) const {
    return arena[holder] + (size_t)parity * half_bytes + (size_t)source * region_bytes;
  }
};
}  // namespace

extern "C" int b200r_exchange_create(int32_t world, int32_t rank, int32_t H, int32_t W, int32_t K,
                                     int64_t n_images_layout, const int32_t* n_images_per_rank,
                                     const int32_t* image_index, const int64_t* face_shift, void* const* arenas,
                                     void** handle) {
  if (world <= 0 || world > MAX_PEERS || rank < 0 || rank >= world || handle == nullptr)
    return fail(B200R_ERR_INVALID_ARGUMENT, "bad world / rank");
  Exchange* ex = new Exchange();
  ex->world = world; ex->rank = rank; ex->H = H; ex->W = W; ex->K = K;
  ex->n_layout = n_images_layout;
  ex->region_bytes = packed_layout(n_images_layout, H, W, K).bytes;
  ex->half_bytes = ex->region_bytes * (size_t)world;
  int total = 0;
  for (int r = 0; r < world; ++r) {
    ex->arena.push_back(static_cast<unsigned char*>(arenas[r]));
    ex->n_images.push_back(n_images_per_rank[r]);
    ex->first_image.push_back(total);
    total += n_images_per_rank[r];
  }
  const int n = total > 0 ? total : 1;
  cudaError_t e = cudaMalloc(&ex->d_image_index, sizeof(int32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_face_shift, sizeof(int64_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_cursor, sizeof(int32_t));
  if (e == cudaSuccess && total > 0)
    e = cudaMemcpy(ex->d_image_index, image_index, sizeof(int32_t) * total, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && total > 0)
    e = cudaMemcpy(ex->d_face_shift, face_shift, sizeof(int64_t) * total, cudaMemcpyHostToDevice);
  cudaEvent_t* evs[] = {&ex->packed, &ex->free_ev[0], &ex->free_ev[1], &ex->done[0], &ex->done[1], &ex->consumed};
  for (cudaEvent_t* ev : evs)
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete ex;
    return check_cuda(e, "b200r_exchange_create");
  }
  *handle = ex;
  return B200R_OK;
}

extern "C" int b200r_exchange_destroy(void* handle) {
  Exchange* ex = static_cast<Exchange*>(handle);
  if (ex == nullptr) return B200R_OK;
  cudaFree(ex->d_image_index);
  cudaFree(ex->d_face_shift);
  cudaFree(ex->d_cursor);
  cudaEvent_t evs[] = {ex->packed, ex->free_ev[0], ex->free_ev[1], ex->done[0], ex->done[1], ex->consumed};
  for (cudaEvent_t ev : evs)
    if (ev) cudaEventDestroy(ev);
  delete ex;
  return B200R_OK;
}

// Step, part 1 (compute stream): pack this rank's frames and push them into every rank's arena half of this step's
// parity; afterwards `side_stream` waits for the pack.  `consumer_stream`: the stream that read the results of the
// previous steps (their buffers are rewritten two steps later): everything enqueued on it so far is ordered before that.
extern "C" int b200r_exchange_push(void* handle, const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                   const float* dists, void* compute_stream, void* side_stream,
                                   void* consumer_stream) {
  Exchange* ex = static_cast<Exchange*>(handle);
  if (ex == nullptr) return fail(B200R_ERR_INVALID_ARGUMENT, "null exchange");
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream), ss = static_cast<cudaStream_t>(side_stream);
  const int parity = (int)(ex->step & 1);
  B200R_CUDA_OK(cudaEventRecord(ex->consumed, static_cast<cudaStream_t>(consumer_stream)));
  ex->have_consumed = true;
  if (ex->have_free[parity]) B200R_CUDA_OK(cudaStreamWaitEvent(cs, ex->free_ev[parity], 0));
  void* dst[MAX_PEERS];
  for (int r = 0; r < ex->world; ++r) dst[r] = ex->region(r, parity, ex->rank);
  const int rc = launch_pack(pix_to_face, zbuf, bary, dists, ex->n_images[ex->rank], ex->H, ex->W, ex->K, ex->n_layout,
                             dst, ex->world, ex->d_cursor, cs);
  if (rc != B200R_OK) return rc;
  B200R_CUDA_OK(cudaEventRecord(ex->packed, cs));
  B200R_CUDA_OK(cudaStreamWaitEvent(ss, ex->packed, 0));
  return B200R_OK;
}

// Step, part 2 (side stream, enqueued behind the caller's cross-rank barrier): expand the streams of all ranks into
// the full-batch buffers.  Returns the step's parity in *parity_out (for b200r_exchange_wait).
extern "C" int b200r_exchange_expand(void* handle, int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                                     void* side_stream, int32_t* parity_out) {
  Exchange* ex = static_cast<Exchange*>(handle);
  if (ex == nullptr) return fail(B200R_ERR_INVALID_ARGUMENT, "null exchange");
  cudaStream_t ss = static_cast<cudaStream_t>(side_stream);
  const int parity = (int)(ex->step & 1);
  // the barrier of step i also proves that every rank has finished expanding step i-1: its arena half is free
  B200R_CUDA_OK(cudaEventRecord(ex->free_ev[parity ^ 1], ss));
  ex->have_free[parity ^ 1] = true;
  if (ex->have_consumed) B200R_CUDA_OK(cudaStreamWaitEvent(ss, ex->consumed, 0));
  const PackedLayout lay = packed_layout(ex->n_layout, ex->H, ex->W, ex->K);
  for (int r = 0; r < ex->world; ++r) {
    const int n_r = ex->n_images[r];
    if (n_r == 0) continue;
    const int rc = launch_unpack(ex->region(ex->rank, parity, r), lay, (int64_t)n_r * ex->H * ex->W, ex->H * ex->W,
                                 ex->K, ex->d_image_index + ex->first_image[r], ex->d_face_shift + ex->first_image[r],
                                 pix_to_face, zbuf, bary, dists, ss);
    if (rc != B200R_OK) return rc;
  }
  B200R_CUDA_OK(cudaEventRecord(ex->done[parity], ss));
  if (parity_out) *parity_out = parity;
  ex->step += 1;
  return B200R_OK;
}

extern "C" int b200r_exchange_wait(void* handle, int32_t parity, void* stream) {
  Exchange* ex = static_cast<Exchange*>(handle);
  if (ex == nullptr || parity < 0 || parity > 1) return fail(B200R_ERR_INVALID_ARGUMENT, "bad exchange / parity");
  B200R_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ex->done[parity], 0));
  return B200R_OK;
}
