"""Small invocations of every kernel family, for `compute-sanitizer --tool memcheck|racecheck` (see profiles/README.md)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, _lib, parallel, synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = synthetic.torus_batch(2, 40, 40, seed=0)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb._b200_all_minus_one = True
nb2 = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb2[0:200:2] = torch.arange(1, 200, 2, device=dev)
nb2[1:200:2] = torch.arange(0, 200, 2, device=dev)
for (size, blur, K, neigh) in [((64, 64), 0.0, 8, nb), ((48, 80), 1e-3, 4, nb), ((33, 47), 1e-3, 16, nb),
                               ((64, 64), 0.0, 16, nb), ((32, 32), 1e-2, 40, nb), ((40, 40), 1e-3, 8, nb2),
                               ((16, 16), 1e-2, 8, nb)]:  # (the last one: > 256 faces per tile, in-kernel list sort)
    out = _C.rasterize_meshes(fv, first, num, neigh, size, blur, K, 0, 0, False, False, False)
    gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])
    _C.rasterize_meshes_backward(fv, out[0], gz, gb, gd, False, False)
    print("meshes", size, blur, K, int((out[0] >= 0).sum()), flush=True)
# many blocks of 256 faces in the setup pass, the last one ragged, one straddling the two meshes
mb = synthetic.torus_batch(2, 331, 331, seed=1)
fvb = synthetic.face_verts_of(mb).to(dev)
nbb = torch.full((fvb.shape[0],), -1, dtype=torch.int64, device=dev)
nbb._b200_all_minus_one = True
outb = _C.rasterize_meshes(fvb, mb.mesh_to_faces_packed_first_idx().to(dev), mb.num_faces_per_mesh().to(dev), nbb, (32, 32),
                           0.0, 4, 0, 0, False, False, False)
print("meshes (many blocks)", fvb.shape[0], int((outb[0] >= 0).sum()), flush=True)
del mb, fvb, nbb, outb
verts, faces = m.verts_packed().to(dev), m.faces_packed().to(dev)
out = _C.rasterize_meshes_indexed(verts, faces, first, num, (64, 64), 0.0, 8, False, False, False)
_C.rasterize_meshes_backward_indexed(out[4], faces, verts.shape[0], out[0], torch.randn_like(out[1]),
                                     torch.randn_like(out[2]), torch.randn_like(out[3]), False, False)
pc = synthetic.random_pointclouds(2, 4000, seed=0)
pts = pc.points_packed().to(dev)
pf, pn = pc.cloud_to_packed_first_idx().to(dev), pc.num_points_per_cloud().to(dev)
rad = torch.full((pts.shape[0],), 0.05, device=dev)
for (size, K) in [((64, 64), 10), ((33, 47), 7), ((16, 16), 4), ((24, 24), 40)]:
    o = _C.rasterize_points(pts, pf, pn, size, rad, K, 0, 0)
    _C.rasterize_points_backward(pts, o[0], torch.randn_like(o[1]), torch.randn_like(o[2]))
    print("points", size, K, int((o[0] >= 0).sum()), flush=True)
# compositing: drop-in alpha composite (strided point-major features) and the fused point renderer op
o = _C.rasterize_points(pts, pf, pn, (33, 47), rad, 6, 0, 0)
feats = torch.rand(pts.shape[0], 4, device=dev).permute(1, 0)
w = (1 - o[2] / (0.05 * 0.05)).permute(0, 3, 1, 2)
il = o[0].long().permute(0, 3, 1, 2)
img = _C.accum_alphacomposite(feats, w, il)
_C.accum_alphacomposite_backward(torch.randn_like(img), feats, w, il)
img2 = _C.points_alpha_render(feats, o[0], o[2], 0.05)
_C.points_alpha_render_backward(torch.randn_like(img2), feats, o[0], o[2], 0.05)
assert torch.equal(img, img2)
print("compositing", tuple(img.shape), flush=True)
# packed frame exchange kernels (one GPU: the region is local memory)
lib = _lib.load()
out = _C.rasterize_meshes(fv, first, num, nb, (64, 64), 0.0, 8, 0, 0, False, False, False)
rb = int(lib.b200r_packed_frames_bytes(2, 64, 64, 8))
arena = torch.zeros(rb + 64, dtype=torch.uint8, device=dev)
base = (arena.data_ptr() + 15) // 16 * 16
cursor = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
dst = (ctypes.c_void_p * 1)(base)
_lib.check(lib.b200r_fragments_pack_push(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                         2, 64, 64, 8, 2, dst, 1, cursor.data_ptr(), stream))
full = [torch.empty_like(t) for t in out]
idx = torch.arange(2, dtype=torch.int32, device=dev)
shift = torch.zeros(2, dtype=torch.int64, device=dev)
_lib.check(lib.b200r_fragments_unpack(base, 2, 64, 64, 8, 2, idx.data_ptr(), shift.data_ptr(), full[0].data_ptr(),
                                      full[1].data_ptr(), full[2].data_ptr(), full[3].data_ptr(), stream))
torch.cuda.synchronize()
assert all(torch.equal(a, b) for a, b in zip(full, out))
print("done", flush=True)
