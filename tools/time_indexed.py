"""North-star batch: plain (face_verts) ops vs the fused (verts, faces) entry points, forward and backward, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = synthetic.torus_batch(8, 187, 187, seed=0)
verts, faces = m.verts_packed().to(dev), m.faces_packed().to(dev)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb._b200_all_minus_one = True
out = _C.rasterize_meshes_indexed(verts, faces, first, num, (512, 512), 0.0, 8, False, False, False)
gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])


def t(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("forward  plain   %7.1f us" % t(lambda: _C.rasterize_meshes(fv, first, num, nb, (512, 512), 0.0, 8, 0, 0, False, False, False)))
print("forward  indexed %7.1f us" % t(lambda: _C.rasterize_meshes_indexed(verts, faces, first, num, (512, 512), 0.0, 8, False, False, False)))
print("backward plain   %7.1f us" % t(lambda: _C.rasterize_meshes_backward(fv, out[0], gz, gb, gd, False, False)))
print("backward indexed %7.1f us" % t(lambda: _C.rasterize_meshes_backward_indexed(out[4], faces, verts.shape[0], out[0], gz, gb, gd, False, False)))
