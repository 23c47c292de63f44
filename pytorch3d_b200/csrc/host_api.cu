// Library globals and the host-buffer (`_host`) entry points of libb200raster.
//
// The `_host` functions are what a host runtime without its own device allocator binds (cgo / JNI /
// ctypes): inputs and outputs are HOST pointers; device staging buffers are kept in a small grow-only
// per-thread cache so that steady-state calls do no cudaMalloc.  Copies are issued on one stream in
// front of / behind the kernels and the call returns after the results have landed.
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200r {

std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}

std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load() != 0; }
static std::atomic<int> g_profiling{0};
bool profiling_enabled() { return g_profiling.load() != 0; }
PhaseTimer& phase_timer() {
  static thread_local PhaseTimer t;
  return t;
}
void PhaseTimer::record(int i, cudaStream_t s) {
  if (!ev[i]) cudaEventCreate(&ev[i]);
  cudaEventRecord(ev[i], s);
}

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int device = -1;
};

struct HostCtx {
  std::vector<DevBuf> bufs;
  cudaStream_t stream = nullptr;
  int device = -1;
};

HostCtx& ctx() {
  static thread_local HostCtx c;
  return c;
}

int ensure_ctx() {
  HostCtx& c = ctx();
  int dev = 0;
  B200R_CUDA_OK(cudaGetDevice(&dev));
  if (c.stream == nullptr || c.device != dev) {
    for (DevBuf& b : c.bufs)
      if (b.p) cudaFree(b.p);
    c.bufs.clear();
    if (c.stream) cudaStreamDestroy(c.stream);
    B200R_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = dev;
  }
  return B200R_OK;
}

// grow-only device buffer number `slot`
int dev_buf(size_t slot, size_t bytes, void** out) {
  HostCtx& c = ctx();
  if (c.bufs.size() <= slot) c.bufs.resize(slot + 1);
  DevBuf& b = c.bufs[slot];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) B200R_CUDA_OK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    B200R_CUDA_OK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
  }
  *out = b.p;
  return B200R_OK;
}

#define B200R_TRY(expr)              \
  do {                               \
    int _rc = (expr);                \
    if (_rc != B200R_OK) return _rc; \
  } while (0)

int h2d(void* d, const void* h, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return B200R_OK;
  B200R_CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
  return B200R_OK;
}
int d2h(void* h, const void* d, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return B200R_OK;
  B200R_CUDA_OK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
  return B200R_OK;
}

}  // namespace
}  // namespace b200r

using namespace b200r;

extern "C" const char* b200r_version(void) { return "b200raster 0.1.0 sm_100a"; }
extern "C" const char* b200r_last_error(void) { return last_error_ref().c_str(); }
extern "C" int64_t b200r_kernel_launch_count(void) { return launch_counter().load(); }
extern "C" void b200r_set_profiling(int32_t enabled) { g_profiling.store(enabled ? 1 : 0); }
extern "C" void b200r_set_pdl(int32_t enabled) { g_pdl.store(enabled ? 1 : 0); }
extern "C" int b200r_last_phase_ms(float out[3]) {
  PhaseTimer& t = phase_timer();
  out[0] = out[1] = out[2] = 0.0f;
  if (t.have_fwd) {
    B200R_CUDA_OK(cudaEventSynchronize(t.ev[2]));
    B200R_CUDA_OK(cudaEventElapsedTime(&out[0], t.ev[0], t.ev[1]));
    B200R_CUDA_OK(cudaEventElapsedTime(&out[1], t.ev[1], t.ev[2]));
  }
  if (t.have_bwd) {
    B200R_CUDA_OK(cudaEventSynchronize(t.ev[4]));
    B200R_CUDA_OK(cudaEventElapsedTime(&out[2], t.ev[3], t.ev[4]));
  }
  return B200R_OK;
}

extern "C" int b200r_rasterize_meshes_forward_host(const float* face_verts, int64_t F, const int64_t* first,
                                                   const int64_t* num, const int64_t* neighbor, int32_t N,
                                                   int32_t H, int32_t W, float blur_radius, int32_t K,
                                                   int32_t perspective_correct, int32_t clip_barycentric_coords,
                                                   int32_t cull_backfaces, int64_t* pix_to_face, float* zbuf,
                                                   float* bary, float* dists) {
  if (F < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  if (neighbor != nullptr) {  // only stage the neighbour table when it carries information
    bool any = false;
    for (int64_t i = 0; i < F && !any; ++i) any = neighbor[i] != -1;
    if (!any) neighbor = nullptr;
  }
  B200R_TRY(ensure_ctx());
  cudaStream_t s = ctx().stream;
  const size_t slots = (size_t)N * H * W * K;
  void *d_fv, *d_first, *d_num, *d_p2f, *d_z, *d_b, *d_d, *d_ws;
  B200R_TRY(dev_buf(0, sizeof(float) * 9 * (size_t)F, &d_fv));
  B200R_TRY(dev_buf(1, sizeof(int64_t) * (size_t)N, &d_first));
  B200R_TRY(dev_buf(2, sizeof(int64_t) * (size_t)N, &d_num));
  B200R_TRY(dev_buf(3, sizeof(int64_t) * slots, &d_p2f));
  B200R_TRY(dev_buf(4, sizeof(float) * slots, &d_z));
  B200R_TRY(dev_buf(5, sizeof(float) * 3 * slots, &d_b));
  B200R_TRY(dev_buf(6, sizeof(float) * slots, &d_d));
  const size_t ws_bytes = b200r_rasterize_meshes_workspace_bytes(F, N, H, W, 0);
  B200R_TRY(dev_buf(7, ws_bytes, &d_ws));
  B200R_TRY(h2d(d_fv, face_verts, sizeof(float) * 9 * (size_t)F, s));
  B200R_TRY(h2d(d_first, first, sizeof(int64_t) * (size_t)N, s));
  B200R_TRY(h2d(d_num, num, sizeof(int64_t) * (size_t)N, s));
  void* d_nb = nullptr;
  if (neighbor != nullptr) {
    B200R_TRY(dev_buf(10, sizeof(int64_t) * (size_t)F, &d_nb));
    B200R_TRY(h2d(d_nb, neighbor, sizeof(int64_t) * (size_t)F, s));
  }
  B200R_TRY(b200r_rasterize_meshes_forward(
      (const float*)d_fv, F, (const int64_t*)d_first, (const int64_t*)d_num, (const int64_t*)d_nb, N, H,
      W, blur_radius, K, 0, 0, perspective_correct, clip_barycentric_coords, cull_backfaces, (int64_t*)d_p2f,
      (float*)d_z, (float*)d_b, (float*)d_d, d_ws, ws_bytes, 0, s));
  B200R_TRY(d2h(pix_to_face, d_p2f, sizeof(int64_t) * slots, s));
  B200R_TRY(d2h(zbuf, d_z, sizeof(float) * slots, s));
  B200R_TRY(d2h(bary, d_b, sizeof(float) * 3 * slots, s));
  B200R_TRY(d2h(dists, d_d, sizeof(float) * slots, s));
  B200R_CUDA_OK(cudaStreamSynchronize(s));
  return B200R_OK;
}

extern "C" int b200r_rasterize_meshes_backward_host(const float* face_verts, int64_t F, const int64_t* pix_to_face,
                                                    const float* grad_zbuf, const float* grad_bary,
                                                    const float* grad_dists, int32_t N, int32_t H, int32_t W,
                                                    int32_t K, int32_t perspective_correct,
                                                    int32_t clip_barycentric_coords, float* grad_face_verts) {
  if (F < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  B200R_TRY(ensure_ctx());
  cudaStream_t s = ctx().stream;
  const size_t slots = (size_t)N * H * W * K;
  void *d_fv, *d_p2f, *d_gz, *d_gb, *d_gd, *d_out;
  B200R_TRY(dev_buf(0, sizeof(float) * 9 * (size_t)F, &d_fv));
  B200R_TRY(dev_buf(3, sizeof(int64_t) * slots, &d_p2f));
  B200R_TRY(dev_buf(4, sizeof(float) * slots, &d_gz));
  B200R_TRY(dev_buf(5, sizeof(float) * 3 * slots, &d_gb));
  B200R_TRY(dev_buf(6, sizeof(float) * slots, &d_gd));
  B200R_TRY(dev_buf(8, sizeof(float) * 9 * (size_t)F, &d_out));
  B200R_TRY(h2d(d_fv, face_verts, sizeof(float) * 9 * (size_t)F, s));
  B200R_TRY(h2d(d_p2f, pix_to_face, sizeof(int64_t) * slots, s));
  B200R_TRY(h2d(d_gz, grad_zbuf, sizeof(float) * slots, s));
  B200R_TRY(h2d(d_gb, grad_bary, sizeof(float) * 3 * slots, s));
  B200R_TRY(h2d(d_gd, grad_dists, sizeof(float) * slots, s));
  B200R_TRY(b200r_rasterize_meshes_backward((const float*)d_fv, F, (const int64_t*)d_p2f, (const float*)d_gz,
                                            (const float*)d_gb, (const float*)d_gd, N, H, W, K,
                                            perspective_correct, clip_barycentric_coords, (float*)d_out, s));
  B200R_TRY(d2h(grad_face_verts, d_out, sizeof(float) * 9 * (size_t)F, s));
  B200R_CUDA_OK(cudaStreamSynchronize(s));
  return B200R_OK;
}

extern "C" int b200r_rasterize_points_forward_host(const float* points, int64_t P, const int64_t* first,
                                                   const int64_t* num, const float* radius, int32_t N, int32_t H,
                                                   int32_t W, int32_t K, int32_t* idx, float* zbuf, float* dists) {
  if (P < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  B200R_TRY(ensure_ctx());
  cudaStream_t s = ctx().stream;
  const size_t slots = (size_t)N * H * W * K;
  void *d_pts, *d_first, *d_num, *d_rad, *d_idx, *d_z, *d_d, *d_ws;
  B200R_TRY(dev_buf(0, sizeof(float) * 3 * (size_t)P, &d_pts));
  B200R_TRY(dev_buf(1, sizeof(int64_t) * (size_t)N, &d_first));
  B200R_TRY(dev_buf(2, sizeof(int64_t) * (size_t)N, &d_num));
  B200R_TRY(dev_buf(9, sizeof(float) * (size_t)P, &d_rad));
  B200R_TRY(dev_buf(3, sizeof(int32_t) * slots, &d_idx));
  B200R_TRY(dev_buf(4, sizeof(float) * slots, &d_z));
  B200R_TRY(dev_buf(6, sizeof(float) * slots, &d_d));
  const size_t ws_bytes = b200r_rasterize_points_workspace_bytes(P, N, H, W, 0);
  B200R_TRY(dev_buf(7, ws_bytes, &d_ws));
  B200R_TRY(h2d(d_pts, points, sizeof(float) * 3 * (size_t)P, s));
  B200R_TRY(h2d(d_first, first, sizeof(int64_t) * (size_t)N, s));
  B200R_TRY(h2d(d_num, num, sizeof(int64_t) * (size_t)N, s));
  B200R_TRY(h2d(d_rad, radius, sizeof(float) * (size_t)P, s));
  B200R_TRY(b200r_rasterize_points_forward((const float*)d_pts, P, (const int64_t*)d_first, (const int64_t*)d_num,
                                           (const float*)d_rad, N, H, W, K, 0, 0, (int32_t*)d_idx, (float*)d_z,
                                           (float*)d_d, d_ws, ws_bytes, 0, s));
  B200R_TRY(d2h(idx, d_idx, sizeof(int32_t) * slots, s));
  B200R_TRY(d2h(zbuf, d_z, sizeof(float) * slots, s));
  B200R_TRY(d2h(dists, d_d, sizeof(float) * slots, s));
  B200R_CUDA_OK(cudaStreamSynchronize(s));
  return B200R_OK;
}

extern "C" int b200r_rasterize_points_backward_host(const float* points, int64_t P, const int32_t* idxs,
                                                    const float* grad_zbuf, const float* grad_dists, int32_t N,
                                                    int32_t H, int32_t W, int32_t K, float* grad_points) {
  if (P < 0 || N < 0 || H < 0 || W < 0 || K < 0) return fail(B200R_ERR_INVALID_ARGUMENT, "negative size");
  B200R_TRY(ensure_ctx());
  cudaStream_t s = ctx().stream;
  const size_t slots = (size_t)N * H * W * K;
  void *d_pts, *d_idx, *d_gz, *d_gd, *d_out;
  B200R_TRY(dev_buf(0, sizeof(float) * 3 * (size_t)P, &d_pts));
  B200R_TRY(dev_buf(3, sizeof(int32_t) * slots, &d_idx));
  B200R_TRY(dev_buf(4, sizeof(float) * slots, &d_gz));
  B200R_TRY(dev_buf(6, sizeof(float) * slots, &d_gd));
  B200R_TRY(dev_buf(8, sizeof(float) * 3 * (size_t)P, &d_out));
  B200R_TRY(h2d(d_pts, points, sizeof(float) * 3 * (size_t)P, s));
  B200R_TRY(h2d(d_idx, idxs, sizeof(int32_t) * slots, s));
  B200R_TRY(h2d(d_gz, grad_zbuf, sizeof(float) * slots, s));
  B200R_TRY(h2d(d_gd, grad_dists, sizeof(float) * slots, s));
  B200R_TRY(b200r_rasterize_points_backward((const float*)d_pts, P, (const int32_t*)d_idx, (const float*)d_gz,
                                            (const float*)d_gd, N, H, W, K, (float*)d_out, s));
  B200R_TRY(d2h(grad_points, d_out, sizeof(float) * 3 * (size_t)P, s));
  B200R_CUDA_OK(cudaStreamSynchronize(s));
  return B200R_OK;
}
