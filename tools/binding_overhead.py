"""Host-side cost of one forward+backward call through each binding of the C ABI (tiny workload: launch-bound)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = synthetic.torus_batch(2, 24, 24, seed=0)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb._b200_all_minus_one = True
out = _C.rasterize_meshes(fv, first, num, nb, (64, 64), 0.0, 8, 0, 0, False, False, False)
gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])
for use_ext in (True, False, True, False):
    _C.USE_EXT = use_ext
    for _ in range(50):
        o = _C.rasterize_meshes(fv, first, num, nb, (64, 64), 0.0, 8, 0, 0, False, False, False)
        _C.rasterize_meshes_backward(fv, o[0], gz, gb, gd, False, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 2000
    for _ in range(n):
        o = _C.rasterize_meshes(fv, first, num, nb, (64, 64), 0.0, 8, 0, 0, False, False, False)
        _C.rasterize_meshes_backward(fv, o[0], gz, gb, gd, False, False)
    t1 = time.perf_counter()  # host time to ENQUEUE n steps
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-16s host enqueue %.1f us/step, with drain %.1f us/step" % (_C.binding(), 1e6 * (t1 - t0) / n, 1e6 * (t2 - t0) / n),
          flush=True)
_C.USE_EXT = True
