"""Module-level surface: Fragments / RasterizationSettings / MeshRasterizer and the points analogue.

Same fields, defaults and forward() protocol as pytorch3d/renderer/mesh/rasterizer.py:19-276 and
pytorch3d/renderer/points/rasterizer.py:21-169, so that a `MeshRenderer(rasterizer=..., shader=...)`
of PyTorch3D accepts these modules (a rasterizer is any nn.Module whose forward returns Fragments).
Cameras are duck-typed (the camera stack itself is outside the hot path): any object with the
PyTorch3D camera protocol works; `cameras=None` means the input is already in NDC.
"""
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .rasterize_meshes import rasterize_meshes
from .rasterize_points import rasterize_points


@dataclass(frozen=True)
class Fragments:
    """Outputs of the mesh rasterizer (rasterizer.py:19-71 of the reference)."""

    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: Optional[torch.Tensor]

    def detach(self) -> "Fragments":
        return Fragments(
            pix_to_face=self.pix_to_face,
            zbuf=self.zbuf.detach(),
            bary_coords=self.bary_coords.detach(),
            dists=self.dists.detach() if self.dists is not None else self.dists,
        )


@dataclass
class RasterizationSettings:
    """Same fields and defaults as the reference dataclass (rasterizer.py:74-136)."""

    image_size: Union[int, Tuple[int, int]] = 256
    blur_radius: float = 0.0
    faces_per_pixel: int = 1
    bin_size: Optional[int] = None
    max_faces_per_bin: Optional[int] = None
    perspective_correct: Optional[bool] = None
    clip_barycentric_coords: Optional[bool] = None
    cull_backfaces: bool = False
    z_clip_value: Optional[float] = None
    cull_to_frustum: bool = False


def _world_to_ndc(cameras, pts_world, kwargs):
    """World -> NDC xy with the view-space z kept as depth (rasterizer.py:196-214, points/rasterizer.py:126-142 of
    the reference).  Per-call camera overrides in `kwargs` (R, T, focal_length, ...) reach every transform: the
    projection is taken from `get_projection_transform(**kwargs)` and applied to the view-space points; cameras
    without a projection matrix (NotImplementedError) go through `transform_points` instead."""
    eps = kwargs.get("eps", None)
    pts_view = cameras.get_world_to_view_transform(**kwargs).transform_points(pts_world, eps=eps)
    to_ndc = cameras.get_ndc_camera_transform(**kwargs)
    try:
        projection = cameras.get_projection_transform(**kwargs)
    except NotImplementedError:
        projection = None
    if projection is not None:
        pts_ndc = projection.compose(to_ndc).transform_points(pts_view, eps=eps)
    else:
        pts_proj = cameras.transform_points(pts_world, eps=eps)
        pts_ndc = to_ndc.transform_points(pts_proj, eps=eps)
    pts_ndc[..., 2] = pts_view[..., 2]
    return pts_ndc


class MeshRasterizer(nn.Module):
    """Rasterizes a batch of heterogeneous meshes (rasterizer.py:139-276 of the reference)."""

    def __init__(self, cameras=None, raster_settings=None) -> None:
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else RasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def transform(self, meshes_world, **kwargs):
        """World -> view -> NDC with the view-space z kept as depth (rasterizer.py:171-217)."""
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass of "
                             "MeshRasterizer")  # (rasterizer.py:183-186 of the reference)
        n_cameras = len(cameras)
        if n_cameras != 1 and n_cameras != len(meshes_world):
            raise ValueError("Wrong number (%r) of cameras for %r meshes" % (n_cameras, len(meshes_world)))
        verts_ndc = _world_to_ndc(cameras, meshes_world.verts_padded(), kwargs)
        return meshes_world.update_padded(new_verts_padded=verts_ndc)

    def forward(self, meshes_world, **kwargs) -> Fragments:
        meshes_proj = self.transform(meshes_world, **kwargs)
        rs = kwargs.get("raster_settings", self.raster_settings)
        clip_barycentric_coords = rs.clip_barycentric_coords
        if clip_barycentric_coords is None:
            clip_barycentric_coords = rs.blur_radius > 0.0
        cameras = kwargs.get("cameras", self.cameras)
        if rs.perspective_correct is not None:
            perspective_correct = rs.perspective_correct
        else:
            perspective_correct = cameras.is_perspective()
        if rs.z_clip_value is not None:
            z_clip = rs.z_clip_value
        else:  # (rasterizer.py:240-246 of the reference)
            znear = cameras.get_znear()
            if isinstance(znear, torch.Tensor):
                znear = znear.min().item()
            z_clip = None if not perspective_correct or znear is None else znear / 2
        pix_to_face, zbuf, bary_coords, dists = rasterize_meshes(
            meshes_proj,
            image_size=rs.image_size,
            blur_radius=rs.blur_radius,
            faces_per_pixel=rs.faces_per_pixel,
            bin_size=rs.bin_size,
            max_faces_per_bin=rs.max_faces_per_bin,
            clip_barycentric_coords=clip_barycentric_coords,
            perspective_correct=perspective_correct,
            cull_backfaces=rs.cull_backfaces,
            z_clip_value=z_clip,
            cull_to_frustum=rs.cull_to_frustum,
        )
        return Fragments(pix_to_face=pix_to_face, zbuf=zbuf, bary_coords=bary_coords, dists=dists)


@dataclass(frozen=True)
class PointFragments:
    """Outputs of the point rasterizer (points/rasterizer.py:21-48 of the reference)."""

    idx: torch.Tensor
    zbuf: torch.Tensor
    dists: torch.Tensor

    def detach(self) -> "PointFragments":
        return PointFragments(idx=self.idx, zbuf=self.zbuf.detach(), dists=self.dists.detach())


@dataclass
class PointsRasterizationSettings:
    """Same fields and defaults as the reference dataclass (points/rasterizer.py:51-78)."""

    image_size: Union[int, Tuple[int, int]] = 256
    radius: Union[float, torch.Tensor] = 0.01
    points_per_pixel: int = 8
    bin_size: Optional[int] = None
    max_points_per_bin: Optional[int] = None


class PointsRasterizer(nn.Module):
    """Rasterizes a batch of point clouds (points/rasterizer.py:81-169 of the reference)."""

    def __init__(self, cameras=None, raster_settings=None) -> None:
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else PointsRasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def transform(self, point_clouds, **kwargs):
        cameras = kwargs.get("cameras", self.cameras)
        if cameras is None:
            raise ValueError("Cameras must be specified either at initialization or in the forward pass of "
                             "PointsRasterizer")  # (points/rasterizer.py:115-118 of the reference)
        return point_clouds.update_padded(_world_to_ndc(cameras, point_clouds.points_padded(), kwargs))

    def forward(self, point_clouds, **kwargs) -> PointFragments:
        points_proj = self.transform(point_clouds, **kwargs)
        rs = kwargs.get("raster_settings", self.raster_settings)
        idx, zbuf, dists2 = rasterize_points(
            points_proj,
            image_size=rs.image_size,
            radius=rs.radius,
            points_per_pixel=rs.points_per_pixel,
            bin_size=rs.bin_size,
            max_points_per_bin=rs.max_points_per_bin,
        )
        return PointFragments(idx=idx, zbuf=zbuf, dists=dists2)
