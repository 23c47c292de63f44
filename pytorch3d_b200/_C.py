"""Operator-level drop-in for the rasterizer part of `pytorch3d._C`.

Same names, positional arguments, return values and error behaviour as the pybind11 ops registered
in pytorch3d/csrc/ext.cpp:53-56:

    rasterize_meshes, rasterize_meshes_backward, rasterize_points, rasterize_points_backward

implemented by calling the C ABI of libb200raster.so (include/b200_raster.h) on the tensors' device
pointers and the current CUDA stream -- through the torch C++ extension csrc/torch_ext.cpp (the binding the reference
uses) or, for the ops it does not cover, through ctypes.  PyTorch is used only for device memory and streams.
There is no CPU path: CPU tensors raise RuntimeError, like a CUDA-less build of the reference does
for CUDA tensors (rasterize_meshes.h:137-139, mirrored).
"""
from typing import Tuple

import torch

from . import _lib

kMaxPointsPerPixel = 150  # rasterization_utils.cuh:48

# Capacity (in (tile, element) pairs) of the bin lists; 0 = library default.  Results never depend on it:
# tiles whose list does not fit fall back to testing every element of their mesh / cloud.  Tests lower it
# to exercise that path.
PAIR_CAPACITY = 0


# The hot ops go through the torch C++ extension over the C ABI (csrc/torch_ext.cpp, built by
# `python -m pytorch3d_b200.build` next to libb200raster.so) -- the same kind of binding `pytorch3d._C` is (ext.cpp:53-56).
# Without it (library built alone) they call the C ABI through ctypes; both run the same kernels of the same library.
_EXT = None
_EXT_TRIED = False
USE_EXT = not _lib.DEV_OVERRIDE  # tests switch it off to exercise the ctypes binding of the same entry points


def _ext():
    global _EXT, _EXT_TRIED
    if not USE_EXT:
        return None
    if not _EXT_TRIED:
        _EXT_TRIED = True
        import importlib.util
        import os
        from . import build as _build
        path = _build.ext_path()
        if os.path.exists(path):
            _lib.load()  # fails loudly if libb200raster.so itself is missing
            spec = importlib.util.spec_from_file_location(_build.EXT_NAME, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _EXT = mod
    return _EXT


def binding():
    """'torch-extension' or 'ctypes': which binding of the C ABI the hot ops use in this process."""
    return "torch-extension" if _ext() is not None else "ctypes"


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else None


def _require_cuda(*named):
    dev = None
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(
                "%s must be a CUDA tensor: pytorch3d_b200 is a B200-native (sm_100a) rasterizer and has "
                "no CPU implementation." % name)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(
                "Expected all tensors to be on the same device (%s is on %s, expected %s)" % (name, t.device, dev))
    return dev


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def rasterize_meshes(
    face_verts: torch.Tensor,
    mesh_to_face_first_idx: torch.Tensor,
    num_faces_per_mesh: torch.Tensor,
    clipped_faces_neighbor_idx: torch.Tensor,
    image_size: Tuple[int, int],
    blur_radius: float,
    faces_per_pixel: int,
    bin_size: int,
    max_faces_per_bin: int,
    perspective_correct: bool,
    clip_barycentric_coords: bool,
    cull_backfaces: bool,
):
    """pytorch3d._C.rasterize_meshes (RasterizeMeshes, rasterize_meshes.h:513-562)."""
    if face_verts.dim() != 3 or face_verts.shape[1] != 3 or face_verts.shape[2] != 3:
        raise RuntimeError("face_verts must have dimensions (num_faces, 3, 3)")
    if num_faces_per_mesh.shape[0] != mesh_to_face_first_idx.shape[0]:
        raise RuntimeError(
            "num_faces_per_mesh must have save size first dimension as mesh_to_faces_packed_first_idx")
    if clipped_faces_neighbor_idx.shape[0] != face_verts.shape[0]:
        raise RuntimeError("clipped_faces_neighbor_idx must have save size first dimension as face_verts")
    if faces_per_pixel > kMaxPointsPerPixel:
        raise RuntimeError("Must have points_per_pixel <= %d" % kMaxPointsPerPixel)
    if face_verts.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float but found %s" % face_verts.dtype)
    dev = _require_cuda(("face_verts", face_verts), ("mesh_to_faces_packed_first_idx", mesh_to_face_first_idx),
                        ("num_faces_per_mesh", num_faces_per_mesh),
                        ("clipped_faces_neighbor_idx", clipped_faces_neighbor_idx))
    ext = _ext()
    if ext is not None:
        tagged = getattr(clipped_faces_neighbor_idx, "_b200_all_minus_one", False)
        return ext.rasterize_meshes(face_verts, mesh_to_face_first_idx, num_faces_per_mesh,
                                    None if tagged else clipped_faces_neighbor_idx,
                                    (int(image_size[0]), int(image_size[1])), float(blur_radius), int(faces_per_pixel),
                                    int(bin_size), int(max_faces_per_bin), bool(perspective_correct),
                                    bool(clip_barycentric_coords), bool(cull_backfaces), int(PAIR_CAPACITY))
    lib = _lib.load()
    H, W = int(image_size[0]), int(image_size[1])
    K = int(faces_per_pixel)
    N, F = int(num_faces_per_mesh.shape[0]), int(face_verts.shape[0])
    fv = face_verts.contiguous()
    first = mesh_to_face_first_idx.contiguous().to(torch.int64)
    num = num_faces_per_mesh.contiguous().to(torch.int64)
    # Clipped-face neighbours (only produced by clip_faces): the kernel variant with the neighbour logic is
    # used unless the caller tagged the tensor as all -1 (our own wrapper does); no host sync either way.
    nb = None
    if F > 0 and not getattr(clipped_faces_neighbor_idx, "_b200_all_minus_one", False):
        nb = clipped_faces_neighbor_idx.contiguous().to(torch.int64)
    with torch.cuda.device(dev):
        pix_to_face = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
        zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        bary = torch.empty((N, H, W, K, 3), dtype=torch.float32, device=dev)
        dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        if pix_to_face.numel() == 0:
            return pix_to_face, zbuf, bary, dists
        ws_bytes = lib.b200r_rasterize_meshes_workspace_bytes(F, N, H, W, PAIR_CAPACITY)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.b200r_rasterize_meshes_forward(
            _ptr(fv), F, _ptr(first), _ptr(num), _ptr(nb), N, H, W, float(blur_radius), K, int(bin_size),
            int(max_faces_per_bin), int(bool(perspective_correct)), int(bool(clip_barycentric_coords)),
            int(bool(cull_backfaces)), _ptr(pix_to_face), _ptr(zbuf), _ptr(bary), _ptr(dists), _ptr(ws),
            ws_bytes, PAIR_CAPACITY, _stream_ptr(dev)))
        # the workspace is only read by kernels already enqueued on this stream
        ws.record_stream(torch.cuda.current_stream(dev))
    return pix_to_face, zbuf, bary, dists


def rasterize_meshes_backward(
    face_verts: torch.Tensor,
    pix_to_face: torch.Tensor,
    grad_zbuf: torch.Tensor,
    grad_bary: torch.Tensor,
    grad_dists: torch.Tensor,
    perspective_correct: bool,
    clip_barycentric_coords: bool,
):
    """pytorch3d._C.rasterize_meshes_backward (RasterizeMeshesBackward, rasterize_meshes.h:211-218)."""
    dev = _require_cuda(("face_verts", face_verts), ("pix_to_face", pix_to_face), ("grad_zbuf", grad_zbuf),
                        ("grad_bary", grad_bary), ("grad_dists", grad_dists))
    for name, t in (("face_verts", face_verts), ("grad_zbuf", grad_zbuf), ("grad_bary", grad_bary),
                    ("grad_dists", grad_dists)):
        if t.dtype != torch.float32:
            raise RuntimeError("Expected tensor for %s to have scalar type Float; but got %s" % (name, t.dtype))
    # same non-determinism contract as the reference (rasterize_meshes.cu:587)
    if torch.are_deterministic_algorithms_enabled() and not torch.is_deterministic_algorithms_warn_only_enabled():
        raise RuntimeError(
            "RasterizeMeshesBackwardCuda does not have a deterministic implementation, but you set "
            "'torch.use_deterministic_algorithms(True)'.")
    ext = _ext()
    if ext is not None:
        return ext.rasterize_meshes_backward(face_verts, pix_to_face, grad_zbuf, grad_bary, grad_dists,
                                             bool(perspective_correct), bool(clip_barycentric_coords))
    lib = _lib.load()
    N, H, W, K = (int(s) for s in pix_to_face.shape)
    F = int(face_verts.shape[0])
    fv = face_verts.contiguous()
    p2f = pix_to_face.contiguous()
    gz, gb, gd = grad_zbuf.contiguous(), grad_bary.contiguous(), grad_dists.contiguous()
    with torch.cuda.device(dev):
        grad_face_verts = torch.empty((F, 3, 3), dtype=torch.float32, device=dev)
        if F == 0:
            return grad_face_verts
        _lib.check(lib.b200r_rasterize_meshes_backward(
            _ptr(fv), F, _ptr(p2f), _ptr(gz), _ptr(gb), _ptr(gd), N, H, W, K, int(bool(perspective_correct)),
            int(bool(clip_barycentric_coords)), _ptr(grad_face_verts), _stream_ptr(dev)))
    return grad_face_verts


def rasterize_meshes_indexed(
    verts_packed: torch.Tensor,
    faces_packed: torch.Tensor,
    mesh_to_face_first_idx: torch.Tensor,
    num_faces_per_mesh: torch.Tensor,
    image_size: Tuple[int, int],
    blur_radius: float,
    faces_per_pixel: int,
    perspective_correct: bool,
    clip_barycentric_coords: bool,
    cull_backfaces: bool,
):
    """Fused `rasterize_meshes(verts_packed[faces_packed], ...)` (no counterpart in pytorch3d._C).

    Returns (pix_to_face, zbuf, bary, dists, face_verts): the four Fragments buffers and the gathered (F,3,3)
    faces that `rasterize_meshes_backward_indexed` needs.
    """
    if verts_packed.dim() != 2 or verts_packed.shape[1] != 3:
        raise RuntimeError("verts_packed must have dimensions (num_verts, 3)")
    if faces_packed.dim() != 2 or faces_packed.shape[1] != 3:
        raise RuntimeError("faces_packed must have dimensions (num_faces, 3)")
    if num_faces_per_mesh.shape[0] != mesh_to_face_first_idx.shape[0]:
        raise RuntimeError(
            "num_faces_per_mesh must have save size first dimension as mesh_to_faces_packed_first_idx")
    if faces_per_pixel > kMaxPointsPerPixel:
        raise RuntimeError("Must have points_per_pixel <= %d" % kMaxPointsPerPixel)
    if verts_packed.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float but found %s" % verts_packed.dtype)
    dev = _require_cuda(("verts_packed", verts_packed), ("faces_packed", faces_packed),
                        ("mesh_to_faces_packed_first_idx", mesh_to_face_first_idx),
                        ("num_faces_per_mesh", num_faces_per_mesh))
    ext = _ext()
    if ext is not None:
        return ext.rasterize_meshes_indexed(verts_packed, faces_packed, mesh_to_face_first_idx, num_faces_per_mesh,
                                            (int(image_size[0]), int(image_size[1])), float(blur_radius),
                                            int(faces_per_pixel), bool(perspective_correct),
                                            bool(clip_barycentric_coords), bool(cull_backfaces), int(PAIR_CAPACITY))
    lib = _lib.load()
    H, W = int(image_size[0]), int(image_size[1])
    K = int(faces_per_pixel)
    N, F, V = int(num_faces_per_mesh.shape[0]), int(faces_packed.shape[0]), int(verts_packed.shape[0])
    verts = verts_packed.contiguous()
    faces = faces_packed.contiguous().to(torch.int64)
    first = mesh_to_face_first_idx.contiguous().to(torch.int64)
    num = num_faces_per_mesh.contiguous().to(torch.int64)
    with torch.cuda.device(dev):
        pix_to_face = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
        zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        bary = torch.empty((N, H, W, K, 3), dtype=torch.float32, device=dev)
        dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        face_verts = torch.empty((F, 3, 3), dtype=torch.float32, device=dev)
        ws_bytes = lib.b200r_rasterize_meshes_workspace_bytes(F, N, H, W, PAIR_CAPACITY)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.b200r_rasterize_meshes_forward_indexed(
            _ptr(verts), V, _ptr(faces), F, _ptr(first), _ptr(num), None, N, H, W, float(blur_radius), K,
            int(bool(perspective_correct)), int(bool(clip_barycentric_coords)), int(bool(cull_backfaces)),
            _ptr(pix_to_face), _ptr(zbuf), _ptr(bary), _ptr(dists), _ptr(face_verts), _ptr(ws), ws_bytes,
            PAIR_CAPACITY, _stream_ptr(dev)))
        ws.record_stream(torch.cuda.current_stream(dev))
    return pix_to_face, zbuf, bary, dists, face_verts


def rasterize_meshes_backward_indexed(
    face_verts: torch.Tensor,
    faces_packed: torch.Tensor,
    num_verts: int,
    pix_to_face: torch.Tensor,
    grad_zbuf: torch.Tensor,
    grad_bary: torch.Tensor,
    grad_dists: torch.Tensor,
    perspective_correct: bool,
    clip_barycentric_coords: bool,
):
    """Backward of `rasterize_meshes_indexed`: the gradient w.r.t. verts_packed, (V, 3)."""
    dev = _require_cuda(("face_verts", face_verts), ("faces_packed", faces_packed), ("pix_to_face", pix_to_face),
                        ("grad_zbuf", grad_zbuf), ("grad_bary", grad_bary), ("grad_dists", grad_dists))
    for name, t in (("face_verts", face_verts), ("grad_zbuf", grad_zbuf), ("grad_bary", grad_bary),
                    ("grad_dists", grad_dists)):
        if t.dtype != torch.float32:
            raise RuntimeError("Expected tensor for %s to have scalar type Float; but got %s" % (name, t.dtype))
    if torch.are_deterministic_algorithms_enabled() and not torch.is_deterministic_algorithms_warn_only_enabled():
        raise RuntimeError(
            "RasterizeMeshesBackwardCuda does not have a deterministic implementation, but you set "
            "'torch.use_deterministic_algorithms(True)'.")
    ext = _ext()
    if ext is not None:
        return ext.rasterize_meshes_backward_indexed(face_verts, faces_packed, int(num_verts), pix_to_face, grad_zbuf,
                                                     grad_bary, grad_dists, bool(perspective_correct),
                                                     bool(clip_barycentric_coords))
    lib = _lib.load()
    N, H, W, K = (int(s) for s in pix_to_face.shape)
    F, V = int(face_verts.shape[0]), int(num_verts)
    fv = face_verts.contiguous()
    faces = faces_packed.contiguous().to(torch.int64)
    p2f = pix_to_face.contiguous()
    gz, gb, gd = grad_zbuf.contiguous(), grad_bary.contiguous(), grad_dists.contiguous()
    with torch.cuda.device(dev):
        grad_verts = torch.empty((V, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.b200r_rasterize_meshes_backward_indexed(
            _ptr(fv), _ptr(faces), F, V, _ptr(p2f), _ptr(gz), _ptr(gb), _ptr(gd), N, H, W, K,
            int(bool(perspective_correct)), int(bool(clip_barycentric_coords)), _ptr(grad_verts), None,
            _stream_ptr(dev)))
    return grad_verts


def rasterize_points(
    points: torch.Tensor,
    cloud_to_packed_first_idx: torch.Tensor,
    num_points_per_cloud: torch.Tensor,
    image_size: Tuple[int, int],
    radius: torch.Tensor,
    points_per_pixel: int,
    bin_size: int,
    max_points_per_bin: int,
):
    """pytorch3d._C.rasterize_points (RasterizePoints, rasterize_points.h:343-374)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    if num_points_per_cloud.shape[0] != cloud_to_packed_first_idx.shape[0]:
        raise RuntimeError(
            "num_points_per_cloud must have same size first dimension as cloud_to_packed_first_idx")
    if radius.dim() != 1 or radius.shape[0] != points.shape[0]:
        raise RuntimeError("radius must be of shape (P,)")
    if points_per_pixel > kMaxPointsPerPixel:
        raise RuntimeError("Must have num_closest <= %d" % kMaxPointsPerPixel)
    if points.dtype != torch.float32 or radius.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    dev = _require_cuda(("points", points), ("cloud_to_packed_first_idx", cloud_to_packed_first_idx),
                        ("num_points_per_cloud", num_points_per_cloud), ("radius", radius))
    ext = _ext()
    if ext is not None:
        return ext.rasterize_points(points, cloud_to_packed_first_idx, num_points_per_cloud,
                                    (int(image_size[0]), int(image_size[1])), radius, int(points_per_pixel),
                                    int(bin_size), int(max_points_per_bin), int(PAIR_CAPACITY))
    lib = _lib.load()
    H, W = int(image_size[0]), int(image_size[1])
    K = int(points_per_pixel)
    N, P = int(num_points_per_cloud.shape[0]), int(points.shape[0])
    pts = points.contiguous()
    first = cloud_to_packed_first_idx.contiguous().to(torch.int64)
    num = num_points_per_cloud.contiguous().to(torch.int64)
    rad = radius.contiguous()
    with torch.cuda.device(dev):
        idx = torch.empty((N, H, W, K), dtype=torch.int32, device=dev)
        zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        if idx.numel() == 0:
            return idx, zbuf, dists
        ws_bytes = lib.b200r_rasterize_points_workspace_bytes(P, N, H, W, PAIR_CAPACITY)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.b200r_rasterize_points_forward(
            _ptr(pts), P, _ptr(first), _ptr(num), _ptr(rad), N, H, W, K, int(bin_size), int(max_points_per_bin),
            _ptr(idx), _ptr(zbuf), _ptr(dists), _ptr(ws), ws_bytes, PAIR_CAPACITY, _stream_ptr(dev)))
        ws.record_stream(torch.cuda.current_stream(dev))
    return idx, zbuf, dists


def rasterize_points_backward(points: torch.Tensor, idxs: torch.Tensor, grad_zbuf: torch.Tensor,
                              grad_dists: torch.Tensor):
    """pytorch3d._C.rasterize_points_backward (RasterizePointsBackward, rasterize_points.h:281-285)."""
    dev = _require_cuda(("points", points), ("idxs", idxs), ("grad_zbuf", grad_zbuf), ("grad_dists", grad_dists))
    if idxs.dtype != torch.int32:
        raise RuntimeError("expected scalar type Int but found %s" % idxs.dtype)
    if torch.are_deterministic_algorithms_enabled() and not torch.is_deterministic_algorithms_warn_only_enabled():
        raise RuntimeError(
            "RasterizePointsBackwardCuda does not have a deterministic implementation, but you set "
            "'torch.use_deterministic_algorithms(True)'.")
    ext = _ext()
    if ext is not None:
        return ext.rasterize_points_backward(points, idxs, grad_zbuf, grad_dists)
    lib = _lib.load()
    N, H, W, K = (int(s) for s in idxs.shape)
    P = int(points.shape[0])
    pts, idx = points.contiguous(), idxs.contiguous()
    gz, gd = grad_zbuf.contiguous(), grad_dists.contiguous()
    with torch.cuda.device(dev):
        grad_points = torch.empty((P, 3), dtype=torch.float32, device=dev)
        if P == 0:
            return grad_points
        _lib.check(lib.b200r_rasterize_points_backward(
            _ptr(pts), P, _ptr(idx), _ptr(gz), _ptr(gd), N, H, W, K, _ptr(grad_points), _stream_ptr(dev)))
    return grad_points


def _strides4(t):
    import ctypes
    return (ctypes.c_int64 * 4)(*[int(v) for v in t.stride()])


def _composite_forward(fn_name, features, alphas, points_idx):
    dev = _require_cuda(("features", features), ("alphas", alphas), ("points_idx", points_idx))
    if features.dtype != torch.float32 or alphas.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    if points_idx.dtype != torch.int64:
        raise RuntimeError("expected scalar type Long but found %s" % points_idx.dtype)
    if features.dim() != 2 or alphas.dim() != 4 or points_idx.dim() != 4 or alphas.shape != points_idx.shape:
        raise RuntimeError("features must be (C, P); alphas and points_idx must both be (N, K, H, W)")
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, K, H, W = (int(v) for v in points_idx.shape)
    feat = features.contiguous()
    with torch.cuda.device(dev):
        result = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
        if result.numel() == 0:
            return result
        if K == 0:
            return result.zero_()
        _lib.check(getattr(lib, fn_name)(
            _ptr(feat), C, P, alphas.data_ptr(), _strides4(alphas), points_idx.data_ptr(), _strides4(points_idx),
            N, K, H, W, _ptr(result), _stream_ptr(dev)))
    return result


def _composite_backward(fn_name, grad_outputs, features, alphas, points_idx):
    dev = _require_cuda(("grad_outputs", grad_outputs), ("features", features), ("alphas", alphas),
                        ("points_idx", points_idx))
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, K, H, W = (int(v) for v in points_idx.shape)
    feat, go = features.contiguous(), grad_outputs.contiguous()
    with torch.cuda.device(dev):
        grad_features = torch.empty((C, P), dtype=torch.float32, device=dev)
        grad_alphas = torch.empty((N, K, H, W), dtype=torch.float32, device=dev)
        if C * P == 0 or grad_alphas.numel() == 0:
            return grad_features.zero_(), grad_alphas.zero_()
        _lib.check(getattr(lib, fn_name)(
            _ptr(go), _ptr(feat), C, P, alphas.data_ptr(), _strides4(alphas), points_idx.data_ptr(),
            _strides4(points_idx), N, K, H, W, _ptr(grad_features), _ptr(grad_alphas), _stream_ptr(dev)))
    return grad_features, grad_alphas


def accum_alphacomposite(features: torch.Tensor, alphas: torch.Tensor, points_idx: torch.Tensor):
    """pytorch3d._C.accum_alphacomposite (alphaCompositeForward, csrc/compositing/alpha_composite.h:59-82).

    features (C,P) f32 (a (C,P) view of point-major memory -- what the renderer passes -- is read in place), alphas
    (N,K,H,W) f32, points_idx (N,K,H,W) i64 (any strides) -> (N,C,H,W) f32."""
    dev = _require_cuda(("features", features), ("alphas", alphas), ("points_idx", points_idx))
    if features.dtype != torch.float32 or alphas.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    if points_idx.dtype != torch.int64:
        raise RuntimeError("expected scalar type Long but found %s" % points_idx.dtype)
    if features.dim() != 2 or alphas.dim() != 4 or points_idx.dim() != 4 or alphas.shape != points_idx.shape:
        raise RuntimeError("features must be (C, P); alphas and points_idx must both be (N, K, H, W)")
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, K, H, W = (int(v) for v in points_idx.shape)
    feat, fs_c, fs_p = _feature_layout(features)
    with torch.cuda.device(dev):
        result = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
        if result.numel() == 0:
            return result
        if K == 0:
            return result.zero_()
        _lib.check(lib.b200r_alpha_composite_forward_strided(
            _ptr(feat), C, P, fs_c, fs_p, alphas.data_ptr(), _strides4(alphas), points_idx.data_ptr(),
            _strides4(points_idx), N, K, H, W, _ptr(result), _stream_ptr(dev)))
    return result


def accum_alphacomposite_backward(grad_outputs: torch.Tensor, features: torch.Tensor, alphas: torch.Tensor,
                                  points_idx: torch.Tensor):
    """pytorch3d._C.accum_alphacomposite_backward (alpha_composite.h:84-116) -> (grad_features, grad_alphas);
    grad_features (C,P) has the memory layout of `features`."""
    dev = _require_cuda(("grad_outputs", grad_outputs), ("features", features), ("alphas", alphas),
                        ("points_idx", points_idx))
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, K, H, W = (int(v) for v in points_idx.shape)
    feat, fs_c, fs_p = _feature_layout(features)
    go = grad_outputs.contiguous()
    with torch.cuda.device(dev):
        if fs_c == 1 and C > 1:
            grad_features = torch.empty((P, C), dtype=torch.float32, device=dev).permute(1, 0)
        else:
            grad_features = torch.empty((C, P), dtype=torch.float32, device=dev)
        grad_alphas = torch.empty((N, K, H, W), dtype=torch.float32, device=dev)
        if C * P == 0 or grad_alphas.numel() == 0:
            return grad_features.zero_(), grad_alphas.zero_()
        _lib.check(lib.b200r_alpha_composite_backward_strided(
            _ptr(go), _ptr(feat), C, P, fs_c, fs_p, alphas.data_ptr(), _strides4(alphas), points_idx.data_ptr(),
            _strides4(points_idx), N, K, H, W, grad_features.data_ptr(), _ptr(grad_alphas), _stream_ptr(dev)))
    return grad_features, grad_alphas


def _feature_layout(features):
    """(tensor to read, stride_c, stride_p): a (C, P) view of point-major memory (`features_packed().permute(1, 0)`,
    what the renderer passes) is read in place; anything else as a contiguous (C, P) array."""
    C, P = (int(v) for v in features.shape)
    if features.stride(0) == 1 and features.stride(1) == C and P > 0:
        return features, 1, C
    f = features.contiguous()
    return f, P, 1


def points_alpha_render(features: torch.Tensor, idx: torch.Tensor, dists: torch.Tensor, radius: float):
    """Fused `accum_alphacomposite(features, 1 - dists / radius**2, idx)` on the rasterizer's own layout (no counterpart
    in pytorch3d._C): features (C,P) f32, idx (N,H,W,K) i32, dists (N,H,W,K) f32 -> (N,C,H,W) f32."""
    dev = _require_cuda(("features", features), ("idx", idx), ("dists", dists))
    if features.dtype != torch.float32 or dists.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    if idx.dtype != torch.int32:
        raise RuntimeError("expected scalar type Int but found %s" % idx.dtype)
    if features.dim() != 2 or idx.dim() != 4 or idx.shape != dists.shape:
        raise RuntimeError("features must be (C, P); idx and dists must both be (N, H, W, K)")
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, H, W, K = (int(v) for v in idx.shape)
    feat, fs_c, fs_p = _feature_layout(features)
    ii, dd = idx.contiguous(), dists.contiguous()
    with torch.cuda.device(dev):
        images = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
        if images.numel() == 0:
            return images
        _lib.check(lib.b200r_points_alpha_render_forward(_ptr(feat), C, P, fs_c, fs_p, _ptr(ii), _ptr(dd),
                                                         float(radius) * float(radius), N, K, H, W, _ptr(images),
                                                         _stream_ptr(dev)))
    return images


def points_alpha_render_backward(grad_images: torch.Tensor, features: torch.Tensor, idx: torch.Tensor,
                                 dists: torch.Tensor, radius: float):
    """Backward of `points_alpha_render` -> (grad_features (C,P) with the memory layout of `features`, grad_dists
    (N,H,W,K))."""
    dev = _require_cuda(("grad_images", grad_images), ("features", features), ("idx", idx), ("dists", dists))
    lib = _lib.load()
    C, P = (int(v) for v in features.shape)
    N, H, W, K = (int(v) for v in idx.shape)
    feat, fs_c, fs_p = _feature_layout(features)
    go, ii, dd = grad_images.contiguous(), idx.contiguous(), dists.contiguous()
    with torch.cuda.device(dev):
        if fs_c == 1 and C > 1:
            grad_features = torch.empty((P, C), dtype=torch.float32, device=dev).permute(1, 0)  # point-major, like feat
        else:
            grad_features = torch.empty((C, P), dtype=torch.float32, device=dev)
        grad_dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
        if C * P == 0 or grad_dists.numel() == 0:
            return grad_features.zero_(), grad_dists.zero_()
        _lib.check(lib.b200r_points_alpha_render_backward(_ptr(go), _ptr(feat), C, P, fs_c, fs_p, _ptr(ii), _ptr(dd),
                                                          float(radius) * float(radius), N, K, H, W,
                                                          grad_features.data_ptr(), _ptr(grad_dists), _stream_ptr(dev)))
    return grad_features, grad_dists


def accum_weightedsum(features: torch.Tensor, alphas: torch.Tensor, points_idx: torch.Tensor):
    """pytorch3d._C.accum_weightedsum (weightedSumForward, csrc/compositing/weighted_sum.h:57-78)."""
    return _composite_forward("b200r_weighted_sum_forward", features, alphas, points_idx)


def accum_weightedsum_backward(grad_outputs: torch.Tensor, features: torch.Tensor, alphas: torch.Tensor,
                               points_idx: torch.Tensor):
    """pytorch3d._C.accum_weightedsum_backward (weighted_sum.h:80-110) -> (grad_features, grad_alphas)."""
    return _composite_backward("b200r_weighted_sum_backward", grad_outputs, features, alphas, points_idx)


def accum_weightedsumnorm(features: torch.Tensor, alphas: torch.Tensor, points_idx: torch.Tensor):
    """pytorch3d._C.accum_weightedsumnorm (weightedSumNormForward, csrc/compositing/norm_weighted_sum.h:57-79)."""
    return _composite_forward("b200r_norm_weighted_sum_forward", features, alphas, points_idx)


def accum_weightedsumnorm_backward(grad_outputs: torch.Tensor, features: torch.Tensor, alphas: torch.Tensor,
                                   points_idx: torch.Tensor):
    """pytorch3d._C.accum_weightedsumnorm_backward (norm_weighted_sum.h:81-112) -> (grad_features, grad_alphas)."""
    return _composite_backward("b200r_norm_weighted_sum_backward", grad_outputs, features, alphas, points_idx)


def interp_face_attrs_forward(pix_to_face: torch.Tensor, barycentric_coords: torch.Tensor, face_attrs: torch.Tensor):
    """pytorch3d._C.interp_face_attrs_forward (csrc/interp_face_attrs/interp_face_attrs.h:45-66):
    pix_to_face (P,) i64, barycentric_coords (P,3) f32, face_attrs (F,3,D) f32 -> (P,D) f32."""
    dev = _require_cuda(("pix_to_face", pix_to_face), ("barycentric_coords", barycentric_coords),
                        ("face_attributes", face_attrs))
    if barycentric_coords.dtype != torch.float32 or face_attrs.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    if pix_to_face.dtype != torch.int64:
        raise RuntimeError("expected scalar type Long but found %s" % pix_to_face.dtype)
    lib = _lib.load()
    P, (F, _, D) = int(pix_to_face.shape[0]), (int(v) for v in face_attrs.shape)
    p2f, bary, attrs = pix_to_face.contiguous(), barycentric_coords.contiguous(), face_attrs.contiguous()
    with torch.cuda.device(dev):
        out = torch.empty((P, D), dtype=torch.float32, device=dev)
        if out.numel() == 0:
            return out
        _lib.check(lib.b200r_interp_face_attrs_forward(_ptr(p2f), _ptr(bary), _ptr(attrs), P, F, D, _ptr(out),
                                                       _stream_ptr(dev)))
    return out


def interp_face_attrs_backward(pix_to_face: torch.Tensor, barycentric_coords: torch.Tensor, face_attrs: torch.Tensor,
                               grad_pix_attrs: torch.Tensor):
    """pytorch3d._C.interp_face_attrs_backward (interp_face_attrs.h:88-118) -> (grad_bary (P,3), grad_attrs (F,3,D))."""
    dev = _require_cuda(("pix_to_face", pix_to_face), ("barycentric_coords", barycentric_coords),
                        ("face_attributes", face_attrs), ("pix_attrs", grad_pix_attrs))
    lib = _lib.load()
    P, (F, _, D) = int(pix_to_face.shape[0]), (int(v) for v in face_attrs.shape)
    p2f, bary, attrs = pix_to_face.contiguous(), barycentric_coords.contiguous(), face_attrs.contiguous()
    gp = grad_pix_attrs.contiguous()
    with torch.cuda.device(dev):
        grad_bary = torch.empty((P, 3), dtype=torch.float32, device=dev)
        grad_attrs = torch.empty((F, 3, D), dtype=torch.float32, device=dev)
        if grad_attrs.numel() == 0 or P == 0:
            return grad_bary.zero_(), grad_attrs.zero_()
        _lib.check(lib.b200r_interp_face_attrs_backward(_ptr(p2f), _ptr(bary), _ptr(attrs), _ptr(gp), P, F, D,
                                                        _ptr(grad_bary), _ptr(grad_attrs), _stream_ptr(dev)))
    return grad_bary, grad_attrs


# ------------------------------------------------------------------------------------------------ test hooks
# pytorch3d/csrc/ext.cpp:69-73: "These are only visible for testing; users should not call them directly".  Provided so
# that the reference's own tests of these entry points can run against this build; none of them is on the product path.

def _coarse(fn_name, elems, first, num, image_size, bin_size, max_per_bin, blur_radius=None, radius=None):
    dev = _require_cuda(("elements", elems), ("first_idx", first), ("num_per_batch", num))
    lib = _lib.load()
    H, W = int(image_size[0]), int(image_size[1])
    N, E = int(num.shape[0]), int(elems.shape[0])
    bin_size, M = int(bin_size), int(max_per_bin)
    if bin_size <= 0:
        raise RuntimeError("bin_size must be positive for the coarse stage")
    BH, BW = 1 + (H - 1) // bin_size, 1 + (W - 1) // bin_size
    if BH >= 22 or BW >= 22:  # kMaxItemsPerBin (rasterize_coarse.cu:244-249)
        raise RuntimeError("In RasterizeCoarseCuda got num_bins_y: %d, num_bins_x: %d, too many bins" % (BH, BW))
    el = elems.contiguous()
    f64, n64 = first.contiguous().to(torch.int64), num.contiguous().to(torch.int64)
    with torch.cuda.device(dev):
        bins = torch.empty((N, BH, BW, M), dtype=torch.int32, device=dev)
        counts = torch.empty((N, BH, BW), dtype=torch.int32, device=dev)
        overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
        if blur_radius is not None:
            _lib.check(lib.b200r_rasterize_meshes_coarse(_ptr(el), E, _ptr(f64), _ptr(n64), N, H, W, float(blur_radius),
                                                         bin_size, M, bins.data_ptr(), counts.data_ptr(),
                                                         overflow.data_ptr(), _stream_ptr(dev)))
        else:
            rad = radius.contiguous()
            _lib.check(lib.b200r_rasterize_points_coarse(_ptr(el), E, _ptr(f64), _ptr(n64), _ptr(rad), N, H, W, bin_size,
                                                         M, bins.data_ptr(), counts.data_ptr(), overflow.data_ptr(),
                                                         _stream_ptr(dev)))
        if int(overflow.item()) != 0:
            import warnings
            warnings.warn("Bin size was too small in the coarse rasterization phase. This caused an overflow, meaning "
                          "output may be incomplete. To solve, try increasing max_faces_per_bin / max_points_per_bin, "
                          "decreasing bin_size, or setting bin_size to 0 to use the naive rasterization.")
        # canonical form: ascending element index inside every bin, -1 padding last
        big = torch.iinfo(torch.int32).max
        bins = torch.where(bins < 0, torch.full_like(bins, big), bins).sort(dim=-1).values
        bins = torch.where(bins == big, torch.full_like(bins, -1), bins)
    return bins


def _rasterize_meshes_coarse(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, blur_radius, bin_size,
                             max_faces_per_bin):
    """pytorch3d._C._rasterize_meshes_coarse (RasterizeMeshesCoarse, rasterize_meshes.h:292-318) -> bin_faces
    (N, BH, BW, M) int32, -1 padded, ascending inside every bin."""
    if face_verts.dim() != 3 or face_verts.shape[1] != 3 or face_verts.shape[2] != 3:
        raise RuntimeError("face_verts must have dimensions (num_faces, 3, 3)")
    return _coarse("meshes", face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, bin_size,
                   max_faces_per_bin, blur_radius=blur_radius)


def _rasterize_points_coarse(points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, radius, bin_size,
                             max_points_per_bin):
    """pytorch3d._C._rasterize_points_coarse (RasterizePointsCoarse, rasterize_points.h:140-166)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    return _coarse("points", points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, bin_size,
                   max_points_per_bin, radius=radius)


def _rasterize_meshes_naive(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, clipped_faces_neighbor_idx,
                            image_size, blur_radius, faces_per_pixel, perspective_correct, clip_barycentric_coords,
                            cull_backfaces):
    """pytorch3d._C._rasterize_meshes_naive (RasterizeMeshesNaive, rasterize_meshes.h:109-145): this build has one path;
    it returns the naive kernel's result for every bin_size."""
    return rasterize_meshes(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, clipped_faces_neighbor_idx,
                            image_size, blur_radius, faces_per_pixel, 0, 0, perspective_correct,
                            clip_barycentric_coords, cull_backfaces)


def _rasterize_points_naive(points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, radius,
                            points_per_pixel):
    """pytorch3d._C._rasterize_points_naive (RasterizePointsNaive, rasterize_points.h:65-95)."""
    return rasterize_points(points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, radius,
                            points_per_pixel, 0, 0)


def _rasterize_meshes_fine(face_verts, bin_faces, clipped_faces_neighbor_idx, image_size, blur_radius, bin_size,
                           faces_per_pixel, perspective_correct, clip_barycentric_coords, cull_backfaces):
    """pytorch3d._C._rasterize_meshes_fine (RasterizeMeshesFine, rasterize_meshes.h:407-452).  The reference's fine stage
    looks only at the faces listed in a pixel's bin; this build bins exactly by itself, so `bin_faces` only tells which
    image a face belongs to (the index range of the faces listed for image n) -- for a table produced by the coarse stage
    (every face in every bin it can touch) the result is the same."""
    if bin_faces.dim() != 4:
        raise RuntimeError("bin_faces must have 4 dimensions")
    N = int(bin_faces.shape[0])
    flat = bin_faces.reshape(N, -1).to(torch.int64)
    big = torch.iinfo(torch.int64).max
    lo = torch.where(flat >= 0, flat, torch.full_like(flat, big)).min(dim=1).values.tolist()  # (host sync: test hook)
    hi = flat.max(dim=1).values.tolist()
    first_l, num_l, at = [], [], 0
    for a, b in zip(lo, hi):  # ascending ranges; an image without listed faces gets an empty range
        if b >= 0:
            first_l.append(int(a))
            num_l.append(int(b - a + 1))
            at = int(b) + 1
        else:
            first_l.append(at)
            num_l.append(0)
    first = torch.tensor(first_l, dtype=torch.int64, device=face_verts.device)
    num = torch.tensor(num_l, dtype=torch.int64, device=face_verts.device)
    return rasterize_meshes(face_verts, first, num, clipped_faces_neighbor_idx, image_size, blur_radius,
                            faces_per_pixel, bin_size, int(bin_faces.shape[3]), perspective_correct,
                            clip_barycentric_coords, cull_backfaces)
