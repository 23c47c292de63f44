"""Why do pytorch3d_b200.MeshRasterizer and the re-bound pytorch3d.renderer.MeshRasterizer differ? (development aid)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import oracle  # noqa: E402
import pytorch3d_b200 as p3b  # noqa: E402
from pytorch3d.renderer import FoVPerspectiveCameras, MeshRasterizer, RasterizationSettings, look_at_view_transform  # noqa: E402
from pytorch3d.utils import ico_sphere  # noqa: E402
from pytorch3d_b200 import _C, install as inst  # noqa: E402

dev = torch.device("cuda:0")
R, T = look_at_view_transform(2.7, 0, 0)
cameras = FoVPerspectiveCameras(device=dev, R=R, T=T)
rs = RasterizationSettings(image_size=512, blur_radius=0.0, faces_per_pixel=1, bin_size=0)
inst.install()
mesh = ico_sphere(5, dev)
rast = MeshRasterizer(cameras=cameras, raster_settings=rs)
f5 = rast(mesh)
mine = p3b.MeshRasterizer(cameras=cameras, raster_settings=p3b.RasterizationSettings(
    image_size=512, blur_radius=0.0, faces_per_pixel=1, bin_size=0))(mesh)
d = mine.pix_to_face != f5.pix_to_face
print("pix_to_face differs in", int(d.sum()), "pixels; zbuf differs in", int((mine.zbuf != f5.zbuf).sum()))
ndc = rast.transform(mesh)
fv = ndc.verts_packed()[ndc.faces_packed()].contiguous()
ndc2 = p3b.MeshRasterizer(cameras=cameras).transform(mesh)
fv2 = ndc2.verts_packed()[ndc2.faces_packed()].contiguous()
print("NDC face_verts equal:", bool(torch.equal(fv, fv2)), "max abs diff", float((fv - fv2).abs().max()))
first, num = ndc.mesh_to_faces_packed_first_idx(), ndc.num_faces_per_mesh()
nbt = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
a = _C.rasterize_meshes(fv, first, num, nbt, (512, 512), 0.0, 1, 0, 0, True, False, False)  # untagged: NB kernel
nbt2 = nbt.clone()
nbt2._b200_all_minus_one = True
b = _C.rasterize_meshes(fv, first, num, nbt2, (512, 512), 0.0, 1, 0, 0, True, False, False)  # tagged: optimistic walk
print("NB kernel vs optimistic kernel: p2f diff", int((a[0] != b[0]).sum()), "zbuf diff", int((a[1] != b[1]).sum()))
o = oracle.rasterize_meshes(fv.cpu().numpy(), first.cpu().numpy(), num.cpu().numpy(), (512, 512), 0.0, 1, True, False, False,
                            arith=oracle.ARITH_CUDA, select=oracle.SELECT_CUDA)
print("NB kernel vs oracle: p2f diff", int((a[0].cpu().numpy() != o[0]).sum()),
      "| optimistic vs oracle:", int((b[0].cpu().numpy() != o[0]).sum()))
print("module f5 vs NB kernel", int((f5.pix_to_face != a[0]).sum()), "| mine vs optimistic kernel", int((mine.pix_to_face != b[0]).sum()))
if d.any():
    idx = d.nonzero()[:5]
    for i in idx:
        n, y, x, k = [int(v) for v in i]
        print("pixel", y, x, "mine", int(mine.pix_to_face[n, y, x, k]), float(mine.zbuf[n, y, x, k]), "real", int(f5.pix_to_face[n, y, x, k]),
              float(f5.zbuf[n, y, x, k]), "oracle", int(o[0][n, y, x, k]), float(o[1][n, y, x, k]))
