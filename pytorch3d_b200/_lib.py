"""ctypes loader of libb200raster.so -- the C-ABI boundary (include/b200_raster.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing the import of the ops
fails loudly with the build command to run.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb200raster.so")
# development: B200R_LIB=<path> runs everything against another build of the library (through the ctypes binding:
# the torch extension is linked against the in-tree library)
DEV_OVERRIDE = os.environ.get("B200R_LIB")
if DEV_OVERRIDE:
    LIB_PATH = os.path.abspath(DEV_OVERRIDE)

_c_f = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i32, _i64, _f32, _sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/b200_raster.h one to one.
SIGNATURES = {
    "b200r_version": (ctypes.c_char_p, []),
    "b200r_last_error": (ctypes.c_char_p, []),
    "b200r_kernel_launch_count": (_i64, []),
    "b200r_set_profiling": (None, [_i32]),
    "b200r_set_pdl": (None, [_i32]),
    "b200r_last_phase_ms": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float)]),
    "b200r_rasterize_meshes_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i64]),
    "b200r_rasterize_meshes_forward": (
        ctypes.c_int,
        [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _i32, _i32, _i32,
         _vp, _vp, _vp, _vp, _vp, _sz, _i64, _vp]),
    "b200r_rasterize_meshes_backward": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_rasterize_meshes_forward_indexed": (
        ctypes.c_int,
        [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _i32,
         _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i64, _vp]),
    "b200r_rasterize_meshes_backward_indexed": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_rasterize_points_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i64]),
    "b200r_rasterize_points_forward": (
        ctypes.c_int,
        [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _i64, _vp]),
    "b200r_rasterize_points_backward": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_alpha_composite_forward": (
        ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_alpha_composite_backward": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_alpha_composite_forward_strided": (
        ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_alpha_composite_backward_strided": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_weighted_sum_forward": (
        ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_weighted_sum_backward": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_norm_weighted_sum_forward": (
        ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_norm_weighted_sum_backward": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_points_alpha_render_forward": (
        ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _f32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200r_points_alpha_render_backward": (
        ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _f32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_interp_face_attrs_forward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "b200r_interp_face_attrs_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "b200r_rasterize_meshes_coarse": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200r_rasterize_points_coarse": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200r_peer_alloc": (ctypes.c_int, [_sz, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "b200r_peer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "b200r_peer_close": (ctypes.c_int, [_vp]),
    "b200r_peer_free": (ctypes.c_int, [_vp]),
    "b200r_packed_frames_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "b200r_fragments_pack_push": (
        ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, ctypes.POINTER(_vp), _i32, _vp, _vp]),
    "b200r_fragments_unpack": (
        ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200r_exchange_create": (
        ctypes.c_int, [_i32, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "b200r_exchange_destroy": (ctypes.c_int, [_vp]),
    "b200r_exchange_push": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200r_exchange_expand": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i32)]),
    "b200r_exchange_wait": (ctypes.c_int, [_vp, _i32, _vp]),
    "b200r_rasterize_meshes_forward_host": (
        ctypes.c_int,
        [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200r_rasterize_meshes_backward_host": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200r_rasterize_points_forward_host": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200r_rasterize_points_backward_host": (
        ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle with typed prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pytorch3d_b200: %s is missing. Build it with `python -m pytorch3d_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().b200r_last_error().decode()


def check(rc):
    """Turn a C-ABI status into the exception the reference op would raise (RuntimeError)."""
    if rc != 0:
        raise RuntimeError(last_error() or ("libb200raster error %d" % rc))
