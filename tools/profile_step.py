"""A few forward+backward steps of one workload, for use under ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, synthetic  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ns"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
if name == "c3":
    pc = synthetic.random_pointclouds(8, 100000, seed=0)
    pts = pc.points_packed().to(dev)
    pf, pn = pc.cloud_to_packed_first_idx().to(dev), pc.num_points_per_cloud().to(dev)
    rad = torch.full((pts.shape[0],), 0.01, device=dev)
    out = _C.rasterize_points(pts, pf, pn, (512, 512), rad, 10, 0, 0)
    gz, gd = torch.randn_like(out[1]), torch.randn_like(out[2])
    for _ in range(steps):
        out = _C.rasterize_points(pts, pf, pn, (512, 512), rad, 10, 0, 0)
        _C.rasterize_points_backward(pts, out[0], gz, gd)
else:
    meshes, (nm, F1, H, W, K, blur) = bench.build_workload(name, 0)
    fv = synthetic.face_verts_of(meshes).to(dev)
    first, num = meshes.mesh_to_faces_packed_first_idx().to(dev), meshes.num_faces_per_mesh().to(dev)
    nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
    nb._b200_all_minus_one = True
    out = _C.rasterize_meshes(fv, first, num, nb, (H, W), blur, K, 0, 0, False, False, False)
    gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])
    for _ in range(steps):
        out = _C.rasterize_meshes(fv, first, num, nb, (H, W), blur, K, 0, 0, False, False, False)
        _C.rasterize_meshes_backward(fv, out[0], gz, gb, gd, False, False)
torch.cuda.synchronize()
print("done")
