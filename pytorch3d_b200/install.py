"""Rebind an installed PyTorch3D onto the B200-native rasterizer ops.

PyTorch3D reaches its native rasterizer through module attributes:
    pytorch3d/renderer/mesh/rasterize_meshes.py:14      from pytorch3d import _C
    pytorch3d/renderer/points/rasterize_points.py:13    from pytorch3d import _C
and calls `_C.rasterize_meshes`, `_C.rasterize_meshes_backward`, `_C.rasterize_points`,
`_C.rasterize_points_backward`.  `install()` replaces that `_C` name *in those two modules only* with a proxy that
serves the four ops from `pytorch3d_b200._C` for CUDA tensors and forwards everything else (including CPU tensors)
to the original module, so `MeshRasterizer` / `PointsRasterizer` / `MeshRenderer` work unchanged.
`uninstall()` restores the originals.
"""
import types

from . import _C as _b200_C

_OPS = ("rasterize_meshes", "rasterize_meshes_backward", "rasterize_points", "rasterize_points_backward")
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
    for modname in ("pytorch3d.renderer.mesh.rasterize_meshes", "pytorch3d.renderer.points.rasterize_points"):
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
