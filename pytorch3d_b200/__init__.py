"""pytorch3d_b200 -- B200-native (sm_100a) differentiable rasterizer, a drop-in for the rasterization
path of PyTorch3D (rasterize_meshes / rasterize_points and their backward ops).

Layers (mirroring the reference's):
    _lib     ctypes binding of the C ABI (include/b200_raster.h, pytorch3d_b200/lib/libb200raster.so)
    _C       operator level: same names / positional signatures as pytorch3d._C
    rasterize_meshes / rasterize_points   reference keyword API + autograd
    rasterizer   Fragments, RasterizationSettings, MeshRasterizer, PointsRasterizer
    parallel     batch sharding over the GPUs of one node (NCCL gather of rendered frames)
    install      rebinds an installed PyTorch3D onto these kernels
"""
from .rasterize_meshes import rasterize_meshes  # noqa: F401
from .rasterize_points import rasterize_points  # noqa: F401
from .rasterizer import (  # noqa: F401
    Fragments,
    MeshRasterizer,
    PointFragments,
    PointsRasterizationSettings,
    PointsRasterizer,
    RasterizationSettings,
)
from .structures import PackedMeshes, PackedPointclouds  # noqa: F401

__version__ = "0.1.0"
