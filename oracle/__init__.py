"""TEST INFRASTRUCTURE ONLY -- CPU parity oracle for the rasterizer hot path.

`oracle/raster_oracle.c` is a plain-C restatement of the reference algorithm
(pytorch3d/csrc/rasterize_meshes/rasterize_meshes_cpu.cpp, rasterize_points_cpu.cpp and the
arithmetic of utils/geometry_utils.{h,cuh}); this module compiles it with gcc and wraps it with
ctypes + numpy.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product (pytorch3d_b200) never does.

Parity status: pinned (see header of raster_oracle.c and tests/test_oracle_*.py).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "raster_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "libraster_oracle.so")

ARITH_CPU, ARITH_CUDA = 0, 1
SELECT_CPU, SELECT_CUDA = 0, 1

_lib = None


def build(force=False):
    """gcc-compile the C restatement (no FMA contraction, no fast-math)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", LIB, SRC,
             "-lm", "-lpthread"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def rasterize_meshes(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, blur_radius=0.0,
                     faces_per_pixel=8, perspective_correct=False, clip_barycentric_coords=False,
                     cull_backfaces=False, clipped_faces_neighbor_idx=None, arith=ARITH_CPU, select=SELECT_CPU,
                     rows=None, nthreads=None):
    """Naive (all faces per pixel) forward.  Returns (pix_to_face i64, zbuf, bary, dists) numpy arrays."""
    fv = _f32(face_verts).reshape(-1, 3, 3)
    first, num = _i64(mesh_to_face_first_idx), _i64(num_faces_per_mesh)
    nb = _i64(clipped_faces_neighbor_idx) if clipped_faces_neighbor_idx is not None else None
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    N, K = len(first), int(faces_per_pixel)
    p2f = np.full((N, H, W, K), -1, np.int64)
    zbuf = np.full((N, H, W, K), -1, np.float32)
    bary = np.full((N, H, W, K, 3), -1, np.float32)
    dists = np.full((N, H, W, K), -1, np.float32)
    r0, r1 = rows if rows is not None else (0, H)
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().oracle_rasterize_meshes_forward(
        _p(fv, ctypes.c_float), _p(first, ctypes.c_int64), _p(num, ctypes.c_int64), _p(nb, ctypes.c_int64),
        N, H, W, ctypes.c_float(blur_radius), K, int(perspective_correct), int(clip_barycentric_coords),
        int(cull_backfaces), int(arith), int(select), int(r0), int(r1), int(nthreads),
        _p(p2f, ctypes.c_int64), _p(zbuf, ctypes.c_float), _p(bary, ctypes.c_float), _p(dists, ctypes.c_float))
    if rc != 0:
        raise RuntimeError("Must have points_per_pixel <= 150")
    return p2f, zbuf, bary, dists


def rasterize_meshes_backward(face_verts, pix_to_face, grad_zbuf, grad_bary, grad_dists,
                              perspective_correct=False, clip_barycentric_coords=False, arith=ARITH_CPU,
                              clip_bwd_uncorrected=False, rows=None):
    fv = _f32(face_verts).reshape(-1, 3, 3)
    p2f = _i64(pix_to_face)
    gz, gb, gd = _f32(grad_zbuf), _f32(grad_bary), _f32(grad_dists)
    N, H, W, K = p2f.shape
    out = np.zeros_like(fv)
    r0, r1 = rows if rows is not None else (0, H)
    lib().oracle_rasterize_meshes_backward(
        _p(fv, ctypes.c_float), ctypes.c_int64(fv.shape[0]), _p(p2f, ctypes.c_int64), _p(gz, ctypes.c_float),
        _p(gb, ctypes.c_float), _p(gd, ctypes.c_float), N, H, W, K, int(perspective_correct),
        int(clip_barycentric_coords), int(arith), int(clip_bwd_uncorrected), int(r0), int(r1),
        _p(out, ctypes.c_float))
    return out


def rasterize_points(points, cloud_to_packed_first_idx, num_points_per_cloud, image_size, radius,
                     points_per_pixel=8, arith=ARITH_CPU, select=SELECT_CPU, rows=None, nthreads=None):
    pts = _f32(points).reshape(-1, 3)
    first, num = _i64(cloud_to_packed_first_idx), _i64(num_points_per_cloud)
    rad = _f32(radius)
    if rad.ndim == 0:
        rad = np.full((pts.shape[0],), float(rad), np.float32)
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    N, K = len(first), int(points_per_pixel)
    idx = np.full((N, H, W, K), -1, np.int32)
    zbuf = np.full((N, H, W, K), -1, np.float32)
    dists = np.full((N, H, W, K), -1, np.float32)
    r0, r1 = rows if rows is not None else (0, H)
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().oracle_rasterize_points_forward(
        _p(pts, ctypes.c_float), _p(first, ctypes.c_int64), _p(num, ctypes.c_int64), _p(rad, ctypes.c_float),
        N, H, W, K, int(arith), int(select), int(r0), int(r1), int(nthreads), _p(idx, ctypes.c_int32),
        _p(zbuf, ctypes.c_float), _p(dists, ctypes.c_float))
    if rc != 0:
        raise RuntimeError("Must have num_closest <= 150")
    return idx, zbuf, dists


def rasterize_points_backward(points, idxs, grad_zbuf, grad_dists, arith=ARITH_CPU):
    pts = _f32(points).reshape(-1, 3)
    idx = np.ascontiguousarray(np.asarray(idxs, dtype=np.int32))
    gz, gd = _f32(grad_zbuf), _f32(grad_dists)
    N, H, W, K = idx.shape
    out = np.zeros_like(pts)
    lib().oracle_rasterize_points_backward(
        _p(pts, ctypes.c_float), ctypes.c_int64(pts.shape[0]), _p(idx, ctypes.c_int32), _p(gz, ctypes.c_float),
        _p(gd, ctypes.c_float), N, H, W, K, int(arith), _p(out, ctypes.c_float))
    return out


def alpha_composite(features, alphas, points_idx, arith=ARITH_CPU):
    """features (C,P), alphas / points_idx (N,K,H,W) -> (N,C,H,W)."""
    f, a = _f32(features), _f32(alphas)
    idx = _i64(points_idx)
    C, P = f.shape
    N, K, H, W = idx.shape
    out = np.zeros((N, C, H, W), np.float32)
    lib().oracle_alpha_composite_forward(_p(f, ctypes.c_float), ctypes.c_int64(C), ctypes.c_int64(P),
                                         _p(a, ctypes.c_float), _p(idx, ctypes.c_int64), N, K, H, W, int(arith),
                                         _p(out, ctypes.c_float))
    return out


def alpha_composite_backward(grad_out, features, alphas, points_idx):
    g, f, a = _f32(grad_out), _f32(features), _f32(alphas)
    idx = _i64(points_idx)
    C, P = f.shape
    N, K, H, W = idx.shape
    gf, ga = np.zeros_like(f), np.zeros_like(a)
    lib().oracle_alpha_composite_backward(_p(g, ctypes.c_float), _p(f, ctypes.c_float), ctypes.c_int64(C),
                                          ctypes.c_int64(P), _p(a, ctypes.c_float), _p(idx, ctypes.c_int64), N, K, H,
                                          W, _p(gf, ctypes.c_float), _p(ga, ctypes.c_float))
    return gf, ga


def weighted_sum(features, alphas, points_idx, norm=False):
    """features (C,P), alphas / points_idx (N,K,H,W) -> (N,C,H,W); norm: divided by max(sum of alphas, 1e-4)."""
    f, a = _f32(features), _f32(alphas)
    idx = _i64(points_idx)
    C, P = f.shape
    N, K, H, W = idx.shape
    out = np.zeros((N, C, H, W), np.float32)
    lib().oracle_weighted_sum_forward(_p(f, ctypes.c_float), ctypes.c_int64(C), ctypes.c_int64(P),
                                      _p(a, ctypes.c_float), _p(idx, ctypes.c_int64), N, K, H, W, int(bool(norm)),
                                      _p(out, ctypes.c_float))
    return out


def weighted_sum_backward(grad_out, features, alphas, points_idx, norm=False):
    g, f, a = _f32(grad_out), _f32(features), _f32(alphas)
    idx = _i64(points_idx)
    C, P = f.shape
    N, K, H, W = idx.shape
    gf, ga = np.zeros_like(f), np.zeros_like(a)
    lib().oracle_weighted_sum_backward(_p(g, ctypes.c_float), _p(f, ctypes.c_float), ctypes.c_int64(C),
                                       ctypes.c_int64(P), _p(a, ctypes.c_float), _p(idx, ctypes.c_int64), N, K, H, W,
                                       int(bool(norm)), _p(gf, ctypes.c_float), _p(ga, ctypes.c_float))
    return gf, ga


def interp_face_attrs(pix_to_face, bary, attrs, arith=ARITH_CPU):
    """pix_to_face (P,), bary (P,3), attrs (F,3,D) -> (P,D)."""
    p2f, b, a = _i64(pix_to_face).reshape(-1), _f32(bary).reshape(-1, 3), _f32(attrs)
    P, D = p2f.shape[0], a.shape[2]
    out = np.zeros((P, D), np.float32)
    lib().oracle_interp_face_attrs_forward(_p(p2f, ctypes.c_int64), _p(b, ctypes.c_float), _p(a, ctypes.c_float),
                                           ctypes.c_int64(P), ctypes.c_int64(D), int(arith), _p(out, ctypes.c_float))
    return out


def interp_face_attrs_backward(pix_to_face, bary, attrs, grad_out):
    p2f, b, a, g = _i64(pix_to_face).reshape(-1), _f32(bary).reshape(-1, 3), _f32(attrs), _f32(grad_out)
    P, (F, _, D) = p2f.shape[0], a.shape
    gb, ga = np.zeros((P, 3), np.float32), np.zeros_like(a)
    lib().oracle_interp_face_attrs_backward(_p(p2f, ctypes.c_int64), _p(b, ctypes.c_float), _p(a, ctypes.c_float),
                                            _p(g, ctypes.c_float), ctypes.c_int64(P), ctypes.c_int64(F),
                                            ctypes.c_int64(D), _p(gb, ctypes.c_float), _p(ga, ctypes.c_float))
    return gb, ga


def load_reference(cuda=False):
    """Import the UNMODIFIED reference ops built by oracle/build_ref.py (None if absent)."""
    import importlib.util
    import torch  # noqa: F401  (the .so links against libtorch)
    name = "ref_raster_cuda" if cuda else "ref_raster_cpu"
    path = os.path.join(HERE, "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
