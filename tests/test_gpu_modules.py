"""Module surface: MeshRasterizer / PointsRasterizer forward() against the reference's golden
masks (tests/test_rasterizer.py:57-166, 445-511 of the reference; masks stored by tests/golden/make_module_masks.py),
with a self-contained FoV-perspective camera stand-in; and -- when the CPU-only reference package travelled to the box
in baseline/_ref -- the REAL pytorch3d.renderer.MeshRasterizer / PointsRasterizer after pytorch3d_b200.install(),
on CUDA tensors, against the reference's own CPU render of the same call."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def masks():
    data = np.load(os.path.join(ROOT, "tests", "golden", "module_masks.npz"))
    out = {}
    for key in data.files:
        name, field = key.rsplit("/", 1)
        out.setdefault(name, {})[field] = data[key]
    return {n: np.unpackbits(v["bits"])[: int(np.prod(v["shape"]))].reshape(tuple(v["shape"])).astype(bool)
            for n, v in out.items()}


# ------------------------------------------------------------------------------ camera / batch stand-ins

class _Transform:
    """Row-vector homogeneous transform with the transform_points / compose protocol of Transform3d."""

    def __init__(self, M):
        self.M = M

    def compose(self, other):
        return _Transform(self.M @ other.M)

    def transform_points(self, points, eps=None):
        ones = torch.ones_like(points[..., :1])
        ph = torch.cat([points, ones], -1) @ self.M
        denom = ph[..., 3:]
        if eps is not None:
            sign = denom.sign() + (denom == 0.0).type_as(denom)
            denom = sign * torch.clamp(denom.abs(), eps)
        return ph[..., :3] / denom


class _FoVCamera:
    """FoV perspective camera (fov 60 deg, znear 1, zfar 100, aspect 1) looking at the origin from (0, 0, dist):
    what look_at_view_transform(dist, 0, 0) + FoVPerspectiveCameras() of the reference give."""

    def __init__(self, dist, device):
        self.device = device
        self.R = torch.tensor([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]], device=device)
        self.T = torch.tensor([0.0, 0.0, float(dist)], device=device)

    def __len__(self):
        return 1

    def to(self, device):
        self.device = device
        self.R, self.T = self.R.to(device), self.T.to(device)
        return self

    def is_perspective(self):
        return True

    def get_znear(self):
        return 1.0

    def get_world_to_view_transform(self, **kwargs):
        R, T = kwargs.get("R", self.R), kwargs.get("T", self.T)
        M = torch.eye(4, device=self.device)
        M[:3, :3] = R.reshape(3, 3)
        M[3, :3] = T.reshape(3)
        return _Transform(M)

    def get_projection_transform(self, **kwargs):
        s = 1.0 / math.tan(math.radians(60.0) / 2.0)
        znear, zfar = 1.0, 100.0
        f1, f2 = zfar / (zfar - znear), -(zfar * znear) / (zfar - znear)
        M = torch.tensor([[s, 0, 0, 0], [0, s, 0, 0], [0, 0, f1, 1.0], [0, 0, f2, 0]], device=self.device)
        return _Transform(M)

    def get_ndc_camera_transform(self, **kwargs):
        return _Transform(torch.eye(4, device=self.device))

    def transform_points(self, points, eps=None, **kwargs):
        return self.get_world_to_view_transform(**kwargs).compose(self.get_projection_transform(**kwargs)) \
            .transform_points(points, eps=eps)


class _Meshes:
    """N copies of one topology with the padded + packed accessors MeshRasterizer uses."""

    def __init__(self, verts_padded, faces):
        self._vp, self._faces = verts_padded, faces
        self._N, self._V = verts_padded.shape[0], verts_padded.shape[1]
        self._F = faces.shape[0]

    def __len__(self):
        return self._N

    def verts_padded(self):
        return self._vp

    def update_padded(self, new_verts_padded):
        return _Meshes(new_verts_padded, self._faces)

    def verts_packed(self):
        return self._vp.reshape(-1, 3)

    def faces_packed(self):
        off = (torch.arange(self._N, device=self._faces.device) * self._V).view(-1, 1, 1)
        return (self._faces[None] + off).reshape(-1, 3)

    def mesh_to_faces_packed_first_idx(self):
        return torch.arange(self._N, device=self._faces.device) * self._F

    def num_faces_per_mesh(self):
        return torch.full((self._N,), self._F, dtype=torch.int64, device=self._faces.device)


class _Clouds:
    def __init__(self, points_padded):
        self._pp = points_padded
        self._N, self._P = points_padded.shape[0], points_padded.shape[1]

    def points_padded(self):
        return self._pp

    def update_padded(self, new_points_padded):
        return _Clouds(new_points_padded)

    def points_packed(self):
        return self._pp.reshape(-1, 3)

    def cloud_to_packed_first_idx(self):
        return torch.arange(self._N, device=self._pp.device) * self._P

    def num_points_per_cloud(self):
        return torch.full((self._N,), self._P, dtype=torch.int64, device=self._pp.device)

    def padded_to_packed_idx(self):
        return torch.arange(self._N * self._P, device=self._pp.device)


def _mismatch(mask, golden):
    return int((mask != golden).sum())


def test_mesh_rasterizer_sphere_against_reference_goldens(built_lib, dev, masks):
    """tests/test_rasterizer.py:67-166 of the reference (_simple_sphere): single mesh, batch of 10, per-call R / T
    overrides (zoom), cameras passed at call time, and the error without cameras."""
    import pytorch3d_b200 as p3b
    from pytorch3d_b200 import synthetic
    verts, faces = synthetic.ico_sphere(5)
    assert faces.shape[0] == 20480
    mesh = _Meshes(verts[None].to(dev), faces.to(dev))
    cameras = _FoVCamera(2.7, dev)
    rs = p3b.RasterizationSettings(image_size=512, blur_radius=0.0, faces_per_pixel=1, bin_size=0)
    rasterizer = p3b.MeshRasterizer(cameras=cameras, raster_settings=rs)
    golden = masks["test_rasterized_sphere_MeshRasterizer"]
    frag = rasterizer(mesh)
    assert isinstance(frag, p3b.Fragments) and frag.pix_to_face.shape == (1, 512, 512, 1)
    mask = (frag.pix_to_face[0, ..., 0] >= 0).cpu().numpy()
    # (our icosphere is built in float64 and rounded once; a handful of silhouette pixels may differ from the image
    #  the reference rendered from its float32 icosphere)
    assert _mismatch(mask, golden) <= 40, "%d pixels differ from the reference image" % _mismatch(mask, golden)
    # perspective camera -> perspective-correct barycentrics and z-clipping at znear / 2 are switched on by default
    assert (frag.zbuf[frag.pix_to_face >= 0] > 1.5).all() and (frag.zbuf[frag.pix_to_face >= 0] < 2.8).all()
    batch = _Meshes(verts[None].expand(10, -1, -1).contiguous().to(dev), faces.to(dev))
    fb = rasterizer(batch)
    for i in range(10):
        assert np.array_equal((fb.pix_to_face[i, ..., 0] >= 0).cpu().numpy(), mask)
    assert torch.equal(fb.pix_to_face[3] - 3 * 20480, frag.pix_to_face[0] - 0) or \
        torch.equal(torch.where(fb.pix_to_face[3] >= 0, fb.pix_to_face[3] - 3 * 20480, fb.pix_to_face[3]),
                    frag.pix_to_face[0])
    # kwargs reach BOTH the view transform (depth) and the projection (xy): zoomed-out view
    T20 = torch.tensor([0.0, 0.0, 20.0], device=dev)
    fz = rasterizer(mesh, R=cameras.R, T=T20)
    zoom = (fz.pix_to_face[0, ..., 0] >= 0).cpu().numpy()
    assert _mismatch(zoom, masks["test_rasterized_sphere_zoom_MeshRasterizer"]) <= 12
    assert (fz.zbuf[fz.pix_to_face >= 0] > 18.9).all()
    # cameras only at call time / not at all
    bare = p3b.MeshRasterizer(raster_settings=rs)
    with pytest.raises(ValueError, match="Cameras must be specified"):
        bare(mesh)
    assert np.array_equal((bare(mesh, cameras=cameras).pix_to_face[0, ..., 0] >= 0).cpu().numpy(), mask)
    # the module is differentiable w.r.t. the world-space vertices
    vw = verts[None].to(dev).clone().requires_grad_(True)
    soft = p3b.RasterizationSettings(image_size=64, blur_radius=1e-3, faces_per_pixel=4)
    out = p3b.MeshRasterizer(cameras=cameras, raster_settings=soft)(_Meshes(vw, faces.to(dev)))
    (out.zbuf.clamp_min(0).sum() + out.dists.clamp(-1, 1).sum()).backward()
    assert torch.isfinite(vw.grad).all() and vw.grad.abs().sum() > 0


def test_points_rasterizer_sphere_against_reference_golden(built_lib, dev, masks):
    """tests/test_rasterizer.py:445-511 of the reference."""
    import pytorch3d_b200 as p3b
    from pytorch3d_b200 import synthetic
    verts, _ = synthetic.ico_sphere(1)
    pts = verts.clone()
    pts[:, 0] += 0.2
    pts[:, 1] += 0.2
    clouds = _Clouds(pts[None].to(dev))
    cameras = _FoVCamera(2.7, dev)
    rs = p3b.PointsRasterizationSettings(image_size=256, radius=5e-2, points_per_pixel=1)
    rasterizer = p3b.PointsRasterizer()
    with pytest.raises(ValueError, match="Cameras must be specified"):
        rasterizer(clouds)
    frag = rasterizer(clouds, cameras=cameras, raster_settings=rs)
    assert isinstance(frag, p3b.PointFragments) and frag.idx.dtype == torch.int32
    mask = (frag.idx[0, ..., 0] >= 0).cpu().numpy()
    golden = masks["test_simple_pointcloud_sphere"]
    assert _mismatch(mask, golden) <= 10, "%d pixels differ from the reference image" % _mismatch(mask, golden)
    batch = _Clouds(pts[None].expand(10, -1, -1).contiguous().to(dev))
    fb = rasterizer(batch, cameras=cameras, raster_settings=rs)
    for i in range(10):
        assert np.array_equal((fb.idx[i, ..., 0] >= 0).cpu().numpy(), mask)


# ------------------------------------------------------------------------------ the real package, re-bound

@pytest.fixture(scope="module")
def real_pytorch3d():
    """The unmodified reference package (CPU-only build, installed by `pip install --target baseline/_ref`, see
    DESIGN.md); git-ignored, shipped to the GPU box by gpurun."""
    path = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(path, "pytorch3d")):
        pytest.skip("baseline/_ref/pytorch3d not present")
    sys.path.insert(0, path)
    try:
        import pytorch3d  # noqa: F401
        from pytorch3d import renderer  # noqa: F401
    except Exception as ex:  # e.g. a missing optional dependency of the reference
        sys.path.remove(path)
        pytest.skip("reference package not importable here: %s" % str(ex)[:200])
    yield path
    from pytorch3d_b200 import install as inst
    inst.uninstall()
    sys.path.remove(path)


def test_install_rebinds_the_real_pytorch3d(built_lib, dev, masks, real_pytorch3d):
    from pytorch3d.renderer import (FoVPerspectiveCameras, MeshRasterizer, PointsRasterizationSettings,
                                    PointsRasterizer, RasterizationSettings, look_at_view_transform)
    from pytorch3d.renderer.compositing import alpha_composite, norm_weighted_sum
    from pytorch3d.structures import Pointclouds
    from pytorch3d.utils import ico_sphere

    from pytorch3d_b200 import install as inst
    R, T = look_at_view_transform(2.7, 0, 0)
    rs = RasterizationSettings(image_size=512, blur_radius=0.0, faces_per_pixel=1, bin_size=0)
    # the reference, untouched, on the host cores
    cpu_mesh = ico_sphere(4)
    cpu_frag = MeshRasterizer(cameras=FoVPerspectiveCameras(R=R, T=T), raster_settings=rs)(cpu_mesh)
    assert len(inst.install()) == 4
    # the same user code on CUDA tensors now runs on the B200-native ops
    from pytorch3d_b200 import _lib
    before = _lib.load().b200r_kernel_launch_count()
    cameras = FoVPerspectiveCameras(device=dev, R=R, T=T)
    rasterizer = MeshRasterizer(cameras=cameras, raster_settings=rs)
    frag = rasterizer(ico_sphere(4, dev))
    assert _lib.load().b200r_kernel_launch_count() > before, "the re-bound ops did not run"
    same = (frag.pix_to_face.cpu() == cpu_frag.pix_to_face)
    assert same.float().mean() >= 0.9999, "%d pixels differ from the reference CPU render" % int((~same).sum())
    # (the reference's own CPU-vs-CUDA tolerances, tests/test_rasterize_meshes.py:543-594: zbuf rtol 1e-4, bary rtol 1e-3;
    #  perspective-corrected values of faces seen edge-on at the silhouette are the ill-conditioned ones)
    zerr = (frag.zbuf.cpu() - cpu_frag.zbuf).abs()[same]
    assert (zerr <= 1e-4 * cpu_frag.zbuf.abs()[same] + 1e-6).all(), float(zerr.max())
    berr = (frag.bary_coords.cpu() - cpu_frag.bary_coords).abs()[same[..., None].expand(-1, -1, -1, -1, 3)]
    assert (berr <= 2e-3).all(), float(berr.max())
    # the reference's golden image (ico_sphere(5)) -- exactly, as its own CUDA test demands (test_rasterizer.py:101)
    f5 = rasterizer(ico_sphere(5, dev))
    assert np.array_equal((f5.pix_to_face[0, ..., 0] >= 0).cpu().numpy(), masks["test_rasterized_sphere_MeshRasterizer"])
    Rz, Tz = look_at_view_transform(20.0, 0, 0, device=dev)
    fz = rasterizer(ico_sphere(5, dev), R=Rz, T=Tz)
    assert np.array_equal((fz.pix_to_face[0, ..., 0] >= 0).cpu().numpy(),
                          masks["test_rasterized_sphere_zoom_MeshRasterizer"])
    # our own module with the real cameras and Meshes gives the same Fragments (fresh cameras: the reference's
    # get_world_to_view_transform stores per-call R / T overrides in the camera object, cameras.py:196-209)
    import pytorch3d_b200 as p3b
    cameras = FoVPerspectiveCameras(device=dev, R=R, T=T)
    mine = p3b.MeshRasterizer(cameras=cameras, raster_settings=p3b.RasterizationSettings(
        image_size=512, blur_radius=0.0, faces_per_pixel=1, bin_size=0))(ico_sphere(5, dev))
    assert torch.equal(mine.pix_to_face, f5.pix_to_face) and torch.equal(mine.zbuf, f5.zbuf)
    # backward through the real package's autograd Function
    sphere = ico_sphere(3, dev)
    verts = sphere.verts_padded().clone().requires_grad_(True)
    soft = RasterizationSettings(image_size=64, blur_radius=1e-3, faces_per_pixel=4)
    out = MeshRasterizer(cameras=cameras, raster_settings=soft)(sphere.update_padded(verts))
    (out.zbuf.clamp_min(0).sum() + out.dists.clamp(-1, 1).sum()).backward()
    assert torch.isfinite(verts.grad).all() and verts.grad.abs().sum() > 0
    # points + compositing
    pv = ico_sphere(1, dev).verts_padded().clone()
    pv[..., :2] += 0.2
    prs = PointsRasterizationSettings(image_size=256, radius=5e-2, points_per_pixel=4)
    pf = PointsRasterizer(cameras=cameras, raster_settings=prs)(Pointclouds(points=pv))
    assert np.array_equal((pf.idx[0, ..., 0] >= 0).cpu().numpy(), masks["test_simple_pointcloud_sphere"])
    w = (1 - pf.dists / (5e-2 ** 2)).permute(0, 3, 1, 2)
    feats = torch.rand(3, pv.shape[1], device=dev)
    img = alpha_composite(pf.idx.long().permute(0, 3, 1, 2), w, feats)
    img2 = norm_weighted_sum(pf.idx.long().permute(0, 3, 1, 2), w, feats)
    assert img.shape == (1, 3, 256, 256) and torch.isfinite(img).all() and torch.isfinite(img2).all()
    inst.uninstall()
