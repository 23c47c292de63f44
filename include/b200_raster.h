/*
 * b200_raster.h -- C ABI of the B200-native differentiable rasterizer (libb200raster.so).
 *
 * This is the drop-in boundary for PyTorch3D's native rasterizer ops.  The reference reaches its
 * kernels through the pybind11 module pytorch3d._C (pytorch3d/csrc/ext.cpp:53-56); the four entry
 * points below take exactly the arguments of those ops, flattened to plain pointers and sizes
 * (no torch types), so any host language can bind them.  INTEGRATION.md shows the binding a
 * PyTorch3D maintainer would add; pytorch3d_b200/_C.py is that binding for this repo.
 *
 * Conventions
 *  - all pointers are DEVICE pointers on the current CUDA device unless the function name ends in
 *    `_host` (then they are host pointers and the call performs the H2D/D2H copies itself);
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous and
 *    never synchronise the host (the reference ops do not either, rasterize_meshes.cu:383-384);
 *  - tensors are dense/contiguous in the layouts documented per argument (the reference ops call
 *    .contiguous() themselves, rasterize_meshes.cu:802-804; the Python host does it here);
 *  - outputs are fully written by the kernels, including the -1 padding of empty slots
 *    (the reference pre-fills with at::full, rasterize_meshes.cu:788-791);
 *  - return value: 0 = ok, otherwise an error code; b200r_last_error() gives the message.  Messages
 *    for argument errors match the reference's TORCH_CHECK / AT_ERROR texts.
 */
#ifndef B200_RASTER_H_
#define B200_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200R_OK 0
#define B200R_ERR_INVALID_ARGUMENT 1
#define B200R_ERR_CUDA 2
#define B200R_ERR_WORKSPACE 3

/* kMaxPointsPerPixel, pytorch3d/csrc/rasterize_points/rasterization_utils.cuh:48 */
#define B200R_MAX_K 150

/* Library / build identification ("b200raster <version> sm_100a"). */
const char* b200r_version(void);

/* Message of the last failing call on the calling thread ("" if none). */
const char* b200r_last_error(void);

/* ------------------------------------------------------------------ meshes ------------------ */

/*
 * Scratch bytes needed by b200r_rasterize_meshes_forward for F packed faces, N meshes and an
 * H x W image.  pair_capacity = number of (tile, face) pairs the bin lists can hold; pass <= 0
 * for the default (min(F * tiles_per_image, 32*F + 64*N*tiles_per_image)).  If the real number of
 * pairs exceeds the capacity the affected tiles transparently fall back to testing every face of
 * their mesh, so results never depend on it (the reference drops faces and prints a warning
 * instead, rasterize_coarse.cu:186-201).
 */
size_t b200r_rasterize_meshes_workspace_bytes(int64_t F, int32_t N, int32_t H, int32_t W, int64_t pair_capacity);

/*
 * Replaces pytorch3d._C.rasterize_meshes
 *   (RasterizeMeshes, pytorch3d/csrc/rasterize_meshes/rasterize_meshes.h:513-562;
 *    call site pytorch3d/renderer/mesh/rasterize_meshes.py:297-310).
 *
 *  face_verts                 float32 (F,3,3)   packed faces in NDC (+X left, +Y up, z = depth)
 *  (mesh_to_face_first_idx / num_faces_per_mesh must describe ASCENDING, NON-OVERLAPPING ranges of the packed array --
 *   what Meshes produces: every face belongs to exactly one image.  The reference's kernels loop over each image's
 *   range and so also accept overlapping ranges; here images other than a face's owner would not see it.)
 *  mesh_to_face_first_idx     int64   (N,)      first packed face of each mesh (ascending)
 *  num_faces_per_mesh         int64   (N,)
 *  clipped_faces_neighbor_idx int64   (F,)      -1 or index of the other half of a clipped face
 *                                               (may be NULL = all -1)
 *  blur_radius, faces_per_pixel (K <= 150), perspective_correct, clip_barycentric_coords,
 *  cull_backfaces             as in the reference.
 *  bin_size, max_faces_per_bin  accepted for signature compatibility; they are performance hints
 *                             in the reference ("should not affect the output",
 *                             rasterize_meshes.py:73-80) and are ignored here: tiling is internal
 *                             and exact.
 * Outputs (all fully written):
 *  pix_to_face int64 (N,H,W,K); zbuf float32 (N,H,W,K); bary float32 (N,H,W,K,3);
 *  dists float32 (N,H,W,K); empty slots = -1.  For every pixel the K nearest faces are returned
 *  in increasing (z, face index) order.
 *  workspace: >= b200r_rasterize_meshes_workspace_bytes(...) bytes of device memory, 16B aligned.
 */
int b200r_rasterize_meshes_forward(const float* face_verts, int64_t F, const int64_t* mesh_to_face_first_idx,
                                   const int64_t* num_faces_per_mesh, const int64_t* clipped_faces_neighbor_idx,
                                   int32_t N, int32_t H, int32_t W, float blur_radius, int32_t faces_per_pixel,
                                   int32_t bin_size, int32_t max_faces_per_bin, int32_t perspective_correct,
                                   int32_t clip_barycentric_coords, int32_t cull_backfaces, int64_t* pix_to_face,
                                   float* zbuf, float* bary, float* dists, void* workspace, size_t workspace_bytes,
                                   int64_t pair_capacity, void* stream);

/*
 * Replaces pytorch3d._C.rasterize_meshes_backward
 *   (RasterizeMeshesBackward, rasterize_meshes.h:211-218; call site rasterize_meshes.py:334-342).
 *  grad_face_verts float32 (F,3,3) is zeroed and accumulated by the call.
 */
int b200r_rasterize_meshes_backward(const float* face_verts, int64_t F, const int64_t* pix_to_face,
                                    const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                    int32_t N, int32_t H, int32_t W, int32_t K, int32_t perspective_correct,
                                    int32_t clip_barycentric_coords, float* grad_face_verts, void* stream);

/*
 * Fused gather entry points: the same rasterization, but taking the packed mesh itself.
 * Replace the `face_verts = verts_packed[faces_packed]` autograd node of the reference's wrapper
 * (pytorch3d/renderer/mesh/rasterize_meshes.py:144-148) together with the op it feeds: the setup pass gathers
 * the vertices of each face, and the backward adds the per-face gradient straight into the vertices.
 *
 *  verts  float32 (V,3) packed vertices;  faces int64 (F,3) packed vertex indices (0 <= index < V; a face with
 *         an out-of-range index is rasterized as NaN, i.e. never hit -- the reference raises a device assert)
 *  face_verts_out  float32 (F,3,3), written by the forward call: the gathered faces, to be passed to the
 *         backward call (what the reference's autograd saves)
 *  grad_verts      float32 (V,3), zeroed and accumulated by the backward call
 *  grad_face_verts_scratch unused since the backward kernel adds every contribution straight to the vertices of its
 *         face (kept in the signature; may be NULL)
 * All other arguments as in b200r_rasterize_meshes_forward / _backward (same workspace size).
 */
int b200r_rasterize_meshes_forward_indexed(const float* verts, int64_t V, const int64_t* faces, int64_t F,
                                           const int64_t* mesh_to_face_first_idx, const int64_t* num_faces_per_mesh,
                                           const int64_t* clipped_faces_neighbor_idx, int32_t N, int32_t H,
                                           int32_t W, float blur_radius, int32_t faces_per_pixel,
                                           int32_t perspective_correct, int32_t clip_barycentric_coords,
                                           int32_t cull_backfaces, int64_t* pix_to_face, float* zbuf, float* bary,
                                           float* dists, float* face_verts_out, void* workspace,
                                           size_t workspace_bytes, int64_t pair_capacity, void* stream);

int b200r_rasterize_meshes_backward_indexed(const float* face_verts, const int64_t* faces, int64_t F, int64_t V,
                                            const int64_t* pix_to_face, const float* grad_zbuf,
                                            const float* grad_bary, const float* grad_dists, int32_t N, int32_t H,
                                            int32_t W, int32_t K, int32_t perspective_correct,
                                            int32_t clip_barycentric_coords, float* grad_verts,
                                            float* grad_face_verts_scratch, void* stream);

/* ------------------------------------------------------------------ points ------------------ */

size_t b200r_rasterize_points_workspace_bytes(int64_t P, int32_t N, int32_t H, int32_t W, int64_t pair_capacity);

/*
 * Replaces pytorch3d._C.rasterize_points
 *   (RasterizePoints, pytorch3d/csrc/rasterize_points/rasterize_points.h:343-374;
 *    call site pytorch3d/renderer/points/rasterize_points.py:200-212).
 *  points float32 (P,3); cloud_to_packed_first_idx / num_points_per_cloud int64 (N,);
 *  radius float32 (P,) in NDC units; points_per_pixel K <= 150.
 * Outputs: idx int32 (N,H,W,K); zbuf float32 (N,H,W,K); dists float32 (N,H,W,K) (squared xy distance).
 */
int b200r_rasterize_points_forward(const float* points, int64_t P, const int64_t* cloud_to_packed_first_idx,
                                   const int64_t* num_points_per_cloud, const float* radius, int32_t N, int32_t H,
                                   int32_t W, int32_t points_per_pixel, int32_t bin_size,
                                   int32_t max_points_per_bin, int32_t* idx, float* zbuf, float* dists,
                                   void* workspace, size_t workspace_bytes, int64_t pair_capacity, void* stream);

/*
 * Replaces pytorch3d._C.rasterize_points_backward
 *   (RasterizePointsBackward, rasterize_points.h:281-285; call site rasterize_points.py:229-231).
 */
int b200r_rasterize_points_backward(const float* points, int64_t P, const int32_t* idxs, const float* grad_zbuf,
                                    const float* grad_dists, int32_t N, int32_t H, int32_t W, int32_t K,
                                    float* grad_points, void* stream);

/* ------------------------------------------------------------------ compositing ------------- */

/*
 * Replaces pytorch3d._C.accum_alphacomposite
 *   (alphaCompositeForward, pytorch3d/csrc/compositing/alpha_composite.h:59-82;
 *    call site pytorch3d/renderer/compositing.py:47-49).
 *  features float32 (C,P) contiguous; alphas float32 and points_idx int64, logical shape (N,K,H,W), addressed
 *  through the four ELEMENT strides given (the renderer passes permuted views of (N,H,W,K) tensors);
 *  result float32 (N,C,H,W) contiguous, fully written.
 */
int b200r_alpha_composite_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                  const int64_t* alpha_strides, const int64_t* points_idx,
                                  const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                  float* result, void* stream);

/*
 * Replaces pytorch3d._C.accum_alphacomposite_backward
 *   (alphaCompositeBackward, alpha_composite.h:84-116; call site compositing.py:58-60).
 *  grad_out (N,C,H,W) contiguous; grad_features (C,P) is zeroed and accumulated; grad_alphas (N,K,H,W)
 *  contiguous, fully written.
 */
int b200r_alpha_composite_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                   const float* alphas, const int64_t* alpha_strides, const int64_t* points_idx,
                                   const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                   float* grad_features, float* grad_alphas, void* stream);

/* The same with the features addressed through element strides: feature (c, p) at features[c * stride_c + p * stride_p]
 * (the renderer passes `features_packed().permute(1, 0)`, a (C,P) view of point-major memory); grad_features uses the
 * same strides.  The two entry points above are the contiguous (C,P) case (stride_c = P, stride_p = 1). */
int b200r_alpha_composite_forward_strided(const float* features, int64_t C, int64_t P, int64_t feature_stride_c,
                                          int64_t feature_stride_p, const float* alphas, const int64_t* alpha_strides,
                                          const int64_t* points_idx, const int64_t* idx_strides, int32_t N, int32_t K,
                                          int32_t H, int32_t W, float* result, void* stream);
int b200r_alpha_composite_backward_strided(const float* grad_out, const float* features, int64_t C, int64_t P,
                                           int64_t feature_stride_c, int64_t feature_stride_p, const float* alphas,
                                           const int64_t* alpha_strides, const int64_t* points_idx,
                                           const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                           float* grad_features, float* grad_alphas, void* stream);

/*
 * Replace pytorch3d._C.accum_weightedsum / accum_weightedsum_backward
 *   (weightedSumForward / Backward, pytorch3d/csrc/compositing/weighted_sum.h:57-78, 80-110) and
 * pytorch3d._C.accum_weightedsumnorm / accum_weightedsumnorm_backward
 *   (weightedSumNormForward / Backward, compositing/norm_weighted_sum.h:57-79, 81-112):
 *   result[n,c,y,x] = sum_k alpha[n,k,y,x] * features[c, idx[n,k,y,x]]   (norm: / max(sum_k alpha, 1e-4)),
 * slots with idx < 0 skipped.  Arguments and layouts exactly as b200r_alpha_composite_forward / _backward.
 */
int b200r_weighted_sum_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                               const int64_t* alpha_strides, const int64_t* points_idx, const int64_t* idx_strides,
                               int32_t N, int32_t K, int32_t H, int32_t W, float* result, void* stream);
int b200r_weighted_sum_backward(const float* grad_outputs, const float* features, int64_t C, int64_t P,
                                const float* alphas, const int64_t* alpha_strides, const int64_t* points_idx,
                                const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                float* grad_features, float* grad_alphas, void* stream);
int b200r_norm_weighted_sum_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                    const int64_t* alpha_strides, const int64_t* points_idx,
                                    const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                    float* result, void* stream);
int b200r_norm_weighted_sum_backward(const float* grad_outputs, const float* features, int64_t C, int64_t P,
                                     const float* alphas, const int64_t* alpha_strides, const int64_t* points_idx,
                                     const int64_t* idx_strides, int32_t N, int32_t K, int32_t H, int32_t W,
                                     float* grad_features, float* grad_alphas, void* stream);

/* ------------------------------------------------------------------ face attribute interpolation */

/*
 * Replaces pytorch3d._C.interp_face_attrs_forward
 *   (InterpFaceAttrsForward, pytorch3d/csrc/interp_face_attrs/interp_face_attrs.h:45-66;
 *    call site pytorch3d/ops/interp_face_attrs.py:66).
 *  pix_to_face int64 (P,); barycentric_coords float32 (P,3); face_attrs float32 (F,3,D); pix_attrs float32 (P,D),
 *  fully written (0 where pix_to_face < 0).
 */
int b200r_interp_face_attrs_forward(const int64_t* pix_to_face, const float* barycentric_coords,
                                    const float* face_attrs, int64_t P, int64_t F, int64_t D, float* pix_attrs,
                                    void* stream);

/*
 * Replaces pytorch3d._C.interp_face_attrs_backward
 *   (InterpFaceAttrsBackward, interp_face_attrs.h:88-118; call site interp_face_attrs.py:74).
 *  grad_barycentric_coords (P,3) fully written; grad_face_attrs (F,3,D) zeroed and accumulated.
 */
int b200r_interp_face_attrs_backward(const int64_t* pix_to_face, const float* barycentric_coords,
                                     const float* face_attrs, const float* grad_pix_attrs, int64_t P, int64_t F,
                                     int64_t D, float* grad_barycentric_coords, float* grad_face_attrs, void* stream);

/* ------------------------------------------------------------------ frame exchange between GPUs ---------- */

/*
 * The path's only collective (BASELINE.json north_star: "NCCL over NVLink only to gather rendered frames"; the
 * reference has no counterpart, tests/test_render_multigpu.py:120-185 only moves modules between devices).
 * Fragments are exchanged in a packed, lossless form -- 1 byte per pixel (number of valid slots) + 24 bytes per
 * VALID slot -- that one kernel writes straight into the memory of every peer over NVLink.
 *
 * Peer memory: b200r_peer_alloc cudaMallocs `bytes` on the current device and returns a 64-byte CUDA IPC handle
 * that another process of the same node turns into a device pointer with b200r_peer_open (peer access is enabled
 * on first use).  b200r_peer_close / b200r_peer_free undo them.
 */
int b200r_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int b200r_peer_open(const unsigned char* handle64, void** ptr);
int b200r_peer_close(void* ptr);
int b200r_peer_free(void* ptr);

/* Bytes of one packed-stream region sized for `n_images` frames of H x W x K slots (worst case: every slot valid). */
size_t b200r_packed_frames_bytes(int64_t n_images, int32_t H, int32_t W, int32_t K);

/*
 * Pack the Fragments of `n_images` local frames (the four outputs of b200r_rasterize_meshes_forward; valid slots
 * first in every pixel, as the rasterizer writes them) and store the stream into each of the `n_dst` regions
 * (HOST array of device pointers: local memory or b200r_peer_open'ed peer memory; each region laid out for
 * `n_images_layout` >= n_images frames).  `cursor`: one device int32 of scratch.  K <= 32, n_dst <= 16.
 * Completion of the kernel on `stream` + any cross-rank synchronisation (e.g. an NCCL barrier enqueued behind it)
 * makes the stream readable on the destination.
 */
int b200r_fragments_pack_push(const int64_t* pix_to_face, const float* zbuf, const float* bary, const float* dists,
                              int32_t n_images, int32_t H, int32_t W, int32_t K, int64_t n_images_layout,
                              void* const* dst_regions, int32_t n_dst, int32_t* cursor, void* stream);

/*
 * Expand a packed stream (device pointer `region`, written by b200r_fragments_pack_push with the same H, W, K and
 * n_images_layout) into dense full-batch buffers: frame j of the stream lands at batch position image_index[j]
 * (device int32 (n_images,)), its face ids shifted by face_shift[j] (device int64 (n_images,): first global packed face
 * of the mesh minus its first face in the sender's local packing); empty slots are written as -1.
 */
int b200r_fragments_unpack(const void* region, int32_t n_images, int32_t H, int32_t W, int32_t K,
                           int64_t n_images_layout, const int32_t* image_index, const int64_t* face_shift,
                           int64_t* pix_to_face, float* zbuf, float* bary, float* dists, void* stream);

/*
 * Exchange context: one per process, holding the arenas of all ranks (own memory + b200r_peer_open'ed peers; each
 * arena = 2 halves x `world` regions of b200r_packed_frames_bytes(n_images_layout, H, W, K) bytes), the batch geometry
 * (HOST arrays: images per rank; for every image of every rank, in rank order, its position in the full batch and its
 * face-id shift) and the events that order one step:
 *   b200r_exchange_push    on compute_stream: pack this rank's frames, push them into every arena; side_stream then
 *                          waits for the pack.  (`consumer_stream`: the stream that reads the results.)
 *   <cross-rank barrier>   the caller's, enqueued on side_stream: e.g. a 4-byte NCCL all-reduce
 *   b200r_exchange_expand  on side_stream: expand all ranks' streams into the dense full-batch buffers
 *   b200r_exchange_wait    make a stream wait for that expansion
 * Arena halves and result buffers alternate with the step's parity; a half is rewritten only behind the NEXT step's
 * barrier, which every rank enqueues behind its own expansion of this step.
 */
int b200r_exchange_create(int32_t world, int32_t rank, int32_t H, int32_t W, int32_t K, int64_t n_images_layout,
                          const int32_t* n_images_per_rank, const int32_t* image_index, const int64_t* face_shift,
                          void* const* arenas, void** handle);
int b200r_exchange_destroy(void* handle);
int b200r_exchange_push(void* handle, const int64_t* pix_to_face, const float* zbuf, const float* bary,
                        const float* dists, void* compute_stream, void* side_stream, void* consumer_stream);
int b200r_exchange_expand(void* handle, int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                          void* side_stream, int32_t* parity_out);
int b200r_exchange_wait(void* handle, int32_t parity, void* stream);

/* ------------------------------------------------------------------ host-buffer entry points - */

/*
 * Same operators with HOST buffers: the call stages inputs to the device (pinned staging is the
 * caller's choice), runs the kernels and copies the results back, synchronising before returning.
 * This is what a non-PyTorch host (cgo / JNI / ctypes) would bind, and what bench.py times as the
 * end-to-end number.
 */
int b200r_rasterize_meshes_forward_host(const float* face_verts, int64_t F, const int64_t* mesh_to_face_first_idx,
                                        const int64_t* num_faces_per_mesh,
                                        const int64_t* clipped_faces_neighbor_idx, int32_t N, int32_t H, int32_t W,
                                        float blur_radius, int32_t faces_per_pixel, int32_t perspective_correct,
                                        int32_t clip_barycentric_coords, int32_t cull_backfaces,
                                        int64_t* pix_to_face, float* zbuf, float* bary, float* dists);

int b200r_rasterize_meshes_backward_host(const float* face_verts, int64_t F, const int64_t* pix_to_face,
                                         const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                         int32_t N, int32_t H, int32_t W, int32_t K, int32_t perspective_correct,
                                         int32_t clip_barycentric_coords, float* grad_face_verts);

int b200r_rasterize_points_forward_host(const float* points, int64_t P, const int64_t* cloud_to_packed_first_idx,
                                        const int64_t* num_points_per_cloud, const float* radius, int32_t N,
                                        int32_t H, int32_t W, int32_t points_per_pixel, int32_t* idx, float* zbuf,
                                        float* dists);

int b200r_rasterize_points_backward_host(const float* points, int64_t P, const int32_t* idxs,
                                         const float* grad_zbuf, const float* grad_dists, int32_t N, int32_t H,
                                         int32_t W, int32_t K, float* grad_points);

/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t b200r_kernel_launch_count(void);

/*
 * Phase timing for bench.py's roofline line.  When enabled, the forward / backward entry points record
 * CUDA events on the caller's stream around their phases; b200r_last_phase_ms synchronises on those
 * events and returns the durations of the most recent call on this thread:
 *   out[0] = binning (setup+count, scan, fill, sort)   out[1] = fine kernel   out[2] = backward kernel
 * (entries of phases that did not run are 0).  Disabled by default; adds no work when disabled.
 */
void b200r_set_profiling(int32_t enabled);
int b200r_last_phase_ms(float out[3]);

/*
 * Fused point rendering (additional entry points, no counterpart in pytorch3d._C): weights = 1 - dists / radius2 and
 * alpha compositing of the point features in one kernel per direction -- what PointsRenderer.forward does between the
 * rasterizer and the image (pytorch3d/renderer/points/renderer.py:63-73) -- reading the rasterizer's outputs as they
 * are: idx int32 (N,H,W,K), dists float32 (N,H,W,K); images float32 (N,C,H,W); feature (c, p) at
 * features[c * feature_stride_c + p * feature_stride_p] (the renderer's `features_packed().permute(1, 0)` is a view of
 * point-major memory: strides (1, C); grad_features uses the same strides).
 * Forward values are bit-identical to the unfused chain; the backward zero-fills grad_features (C,P), accumulates it
 * with atomics and writes grad_dists (N,H,W,K) = d loss / d dists.
 */
int b200r_points_alpha_render_forward(const float* features, int64_t C, int64_t P, int64_t feature_stride_c,
                                      int64_t feature_stride_p, const int32_t* idx, const float* dists,
                                      float radius2, int32_t N, int32_t K, int32_t H, int32_t W, float* images,
                                      void* stream);
int b200r_points_alpha_render_backward(const float* grad_images, const float* features, int64_t C, int64_t P,
                                       int64_t feature_stride_c, int64_t feature_stride_p, const int32_t* idx,
                                       const float* dists, float radius2, int32_t N, int32_t K, int32_t H, int32_t W,
                                       float* grad_features, float* grad_dists, void* stream);

/*
 * Test hooks of the reference's coarse stage (pytorch3d._C._rasterize_meshes_coarse / _rasterize_points_coarse,
 * pytorch3d/csrc/ext.cpp:69-73; RasterizeMeshesCoarse rasterize_meshes.h:292-318, RasterizePointsCoarse
 * rasterize_points.h:140-166): the dense table bin_faces / bin_points int32 (N, BH, BW, M), BH = 1 + (H-1)/bin_size,
 * -1 padded, elements of a bin in arbitrary order (sort for a canonical form).  bin_counts int32 (N, BH, BW) scratch;
 * *overflow (device int32) is set to 1 if a bin received more than M elements (the reference prints a warning,
 * rasterize_coarse.cu:186-201).  Not used by the rasterizer itself, whose tile lists are compact and exact.
 */
int b200r_rasterize_meshes_coarse(const float* face_verts, int64_t F, const int64_t* mesh_to_face_first_idx,
                                  const int64_t* num_faces_per_mesh, int32_t N, int32_t H, int32_t W,
                                  float blur_radius, int32_t bin_size, int32_t max_faces_per_bin, int32_t* bin_faces,
                                  int32_t* bin_counts, int32_t* overflow, void* stream);
int b200r_rasterize_points_coarse(const float* points, int64_t P, const int64_t* cloud_to_packed_first_idx,
                                  const int64_t* num_points_per_cloud, const float* radius, int32_t N, int32_t H,
                                  int32_t W, int32_t bin_size, int32_t max_points_per_bin, int32_t* bin_points,
                                  int32_t* bin_counts, int32_t* overflow, void* stream);

/*
 * Programmatic dependent launch between the kernels of one call (setup -> scan -> fill -> fine; backward -> scatter):
 * the next kernel is made resident while its predecessor drains.  On by default; results never depend on it.
 */
void b200r_set_pdl(int32_t enabled);

#ifdef __cplusplus
}
#endif
#endif /* B200_RASTER_H_ */
