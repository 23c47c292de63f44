"""Rebind an installed PyTorch3D onto the B200-native rasterizer ops.

PyTorch3D reaches its native rasterizer through module attributes:
    pytorch3d/renderer/mesh/rasterize_meshes.py:14      from pytorch3d import _C
    pytorch3d/renderer/points/rasterize_points.py:13    from pytorch3d import _C
    pytorch3d/renderer/compositing.py:10                from pytorch3d import _C
    pytorch3d/ops/interp_face_attrs.py:10               from pytorch3d import _C
and calls `_C.rasterize_meshes[_backward]`, `_C.rasterize_points[_backward]`, `_C.accum_alphacomposite[_backward]`,
`_C.accum_weightedsum[_backward]`, `_C.accum_weightedsumnorm[_backward]`, `_C.interp_face_attrs_forward/_backward`.
`install()` replaces that `_C` name *in those modules only* with a proxy that serves these ops from
`pytorch3d_b200._C` for CUDA tensors and forwards everything else (including CPU tensors) to the original module,
so `MeshRasterizer` / `PointsRasterizer` / `MeshRenderer` / `PointsRenderer` work unchanged.
`uninstall()` restores the originals.
"""
import types

from . import _C as _b200_C

_OPS = ("rasterize_meshes", "rasterize_meshes_backward", "rasterize_points", "rasterize_points_backward",
        "accum_alphacomposite", "accum_alphacomposite_backward", "accum_weightedsum", "accum_weightedsum_backward",
        "accum_weightedsumnorm", "accum_weightedsumnorm_backward", "interp_face_attrs_forward",
        "interp_face_attrs_backward")
_MODULES = ("pytorch3d.renderer.mesh.rasterize_meshes", "pytorch3d.renderer.points.rasterize_points",
            "pytorch3d.renderer.compositing", "pytorch3d.ops.interp_face_attrs")
_saved = {}


class _Proxy(types.ModuleType):
    def __init__(self, original):
        super().__init__("pytorch3d_b200._C_proxy")
        self.__dict__["_original"] = original

    def __getattr__(self, name):
        original = self.__dict__["_original"]
        if name in _OPS:
            ours = getattr(_b200_C, name)
            theirs = getattr(original, name, None)

            def dispatch(*args, **kwargs):
                first = args[0] if args else None
                if theirs is not None and first is not None and not getattr(first, "is_cuda", False):
                    return theirs(*args, **kwargs)  # CPU tensors keep the reference's CPU path
                return ours(*args, **kwargs)

            return dispatch
        return getattr(original, name)


def install():
    """Patch pytorch3d (must be importable).  Returns the list of patched module names."""
    import importlib
    patched = []
    for modname in _MODULES:
        mod = importlib.import_module(modname)
        if modname not in _saved:
            _saved[modname] = mod._C
            mod._C = _Proxy(mod._C)
        patched.append(modname)
    return patched


def uninstall():
    import importlib
    for modname, original in list(_saved.items()):
        importlib.import_module(modname)._C = original
        del _saved[modname]
