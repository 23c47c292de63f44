"""Binning / fine / backward phase times (CUDA events inside the library) of one mesh workload of bench.py."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorch3d_b200 import _C, _lib, synthetic  # noqa: E402

args = sys.argv[1:]
if args and args[0] == "--lib":  # development: time another build of the library
    _lib.LIB_PATH = os.path.abspath(args[1])
    args = args[2:]
dev = torch.device("cuda:0")
lib = _lib.load()
buf = (ctypes.c_float * 3)()
for name in args or ["ns", "c2"]:
    nm, rings, sides, H, W, K, blur = bench.WORKLOADS[name]
    meshes = synthetic.torus_batch(nm, rings, sides, seed=0)
    fv = synthetic.face_verts_of(meshes).to(dev)
    first = meshes.mesh_to_faces_packed_first_idx().to(dev)
    num = meshes.num_faces_per_mesh().to(dev)
    nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
    nb._b200_all_minus_one = True
    out = _C.rasterize_meshes(fv, first, num, nb, (H, W), blur, K, 0, 0, False, False, False)
    gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])
    lib.b200r_set_profiling(1)
    t = []
    for _ in range(15):
        torch.cuda._sleep(400000)
        out = _C.rasterize_meshes(fv, first, num, nb, (H, W), blur, K, 0, 0, False, False, False)
        _C.rasterize_meshes_backward(fv, out[0], gz, gb, gd, False, False)
        torch.cuda.synchronize()
        lib.b200r_last_phase_ms(buf)
        t.append(list(buf))
    lib.b200r_set_profiling(0)
    m = np.median(np.array(t[3:]), axis=0) * 1e3
    print("%s: binning %.1f us  fine %.1f us  backward %.1f us  (hits/slot %.3f)" % (
        name, m[0], m[1], m[2], float((out[0] >= 0).float().mean())), flush=True)
