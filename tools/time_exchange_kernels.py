"""Times the pack / expand kernels of the frame exchange on one GPU (the regions are local memory); development aid."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
nm, H, W, K = 8, 512, 512, 8
sources = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = synthetic.torus_batch(nm, 187, 187, seed=0)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb._b200_all_minus_one = True
f = _C.rasterize_meshes(fv, first, num, nb, (H, W), 0.0, K, 0, 0, False, False, False)
rb = int(lib.b200r_packed_frames_bytes(nm, H, W, K))
arena = torch.zeros(sources * rb + 64, dtype=torch.uint8, device=dev)
base = (arena.data_ptr() + 15) // 16 * 16
cursor = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
dst = (ctypes.c_void_p * sources)(*[base + r * rb for r in range(sources)])
N = nm * sources
full = [torch.empty((N, H, W, K), dtype=torch.int64, device=dev), torch.empty((N, H, W, K), device=dev),
        torch.empty((N, H, W, K, 3), device=dev), torch.empty((N, H, W, K), device=dev)]
idx = [torch.arange(r * nm, (r + 1) * nm, dtype=torch.int32, device=dev) for r in range(sources)]
shift = torch.zeros(nm, dtype=torch.int64, device=dev)


def pack():
    _lib.check(lib.b200r_fragments_pack_push(f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(), f[3].data_ptr(), nm, H, W, K,
                                             nm, dst, sources, cursor.data_ptr(), stream))


def unpack():
    for r in range(sources):
        _lib.check(lib.b200r_fragments_unpack(base + r * rb, nm, H, W, K, nm, idx[r].data_ptr(), shift.data_ptr(),
                                              full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr(),
                                              full[3].data_ptr(), stream))


def ms(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


tp = ms(pack)
tu = ms(unpack)
for r in range(sources):
    for a, b in zip(full, f):
        assert torch.equal(a[r * nm:(r + 1) * nm], b)
hits = int((f[0] >= 0).sum())
dense = N * H * W * K * 28
print("pack to %d local regions: %.3f ms | expand %d sources (%d frames): %.3f ms = %.0f GB/s of dense output | hits %d"
      % (sources, tp, sources, N, tu, dense / tu / 1e6, hits), flush=True)
