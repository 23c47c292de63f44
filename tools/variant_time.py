"""Times the north-star forward pass with experimental builds of the library (development aid).

    python tools/variant_time.py build     # here: nvcc, one .so per variant under /root/repo/gpurun_variants/
    python tools/variant_time.py           # on the GPU box: load each variant with ctypes and time it
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_variants")
VARIANTS = {"oldwalk": ["-DB200R_EXP_OLDWALK"]}
if len(sys.argv) > 2:  # python tools/variant_time.py build name=-DFLAG,-DFLAG2 ...
    VARIANTS = {a.split("=", 1)[0]: [f for f in a.split("=", 1)[1].split(",") if f] for a in sys.argv[2:]}

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from pytorch3d_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        if flags is None:
            continue  # (round1: built by hand from the round-1 sources, `git archive f87840d pytorch3d_b200/csrc`)
        lib = os.path.join(OUT, "lib_%s.so" % name)
        cmd = [b._nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
               "-fPIC", "-shared", "-o", lib] + flags + [os.path.join(b.CSRC, f) for f in b.SOURCES]
        subprocess.check_call(cmd)
        print("built", lib)
    sys.exit(0)

import torch  # noqa: E402
from pytorch3d_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = synthetic.torus_batch(8, 187, 187, seed=0)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
N, H, W, K, F = 8, 512, 512, 8, fv.shape[0]
p2f = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
z, d, b = torch.empty((N, H, W, K), device=dev), torch.empty((N, H, W, K), device=dev), torch.empty((N, H, W, K, 3), device=dev)
for name in VARIANTS:
    path = os.path.join(OUT, "lib_%s.so" % name)
    if not os.path.exists(path):
        continue
    lib = ctypes.CDLL(path)
    for fn, (res, args) in _lib.SIGNATURES.items():
        try:
            f = getattr(lib, fn)
        except AttributeError:
            continue
        f.restype, f.argtypes = res, args
    ws_bytes = lib.b200r_rasterize_meshes_workspace_bytes(F, N, H, W, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def fwd():
        rc = lib.b200r_rasterize_meshes_forward(fv.data_ptr(), F, first.data_ptr(), num.data_ptr(), None, N, H, W, 0.0, K,
                                                0, 0, 0, 0, 0, p2f.data_ptr(), z.data_ptr(), b.data_ptr(), d.data_ptr(),
                                                ws.data_ptr(), ws_bytes, 0, stream)
        assert rc == 0

    for _ in range(5):
        fwd()
    lib.b200r_set_profiling(1)
    buf = (ctypes.c_float * 3)()
    acc = [0.0, 0.0]
    for _ in range(20):
        torch.cuda._sleep(400000)
        fwd()
        lib.b200r_last_phase_ms(buf)
        acc[0] += buf[0]
        acc[1] += buf[1]
    lib.b200r_set_profiling(0)
    print("%-20s binning %.4f ms  fine %.4f ms  hits %d" % (name, acc[0] / 20, acc[1] / 20, int((p2f >= 0).sum())), flush=True)
