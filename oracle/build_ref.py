"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference rasterizer into oracle/_ref/.

Compiles, from the sources where they lie under /root/reference (never copied):
  pytorch3d/csrc/rasterize_meshes/rasterize_meshes_cpu.cpp
  pytorch3d/csrc/rasterize_points/rasterize_points_cpu.cpp
  pytorch3d/csrc/compositing/alpha_composite_cpu.cpp  (+ alpha_composite.cu)
  pytorch3d/csrc/rasterize_meshes/rasterize_meshes.cu      (sm_100a)
  pytorch3d/csrc/rasterize_coarse/rasterize_coarse.cu      (sm_100a)
  pytorch3d/csrc/rasterize_points/rasterize_points.cu      (sm_100a)
plus oracle/ref_shim.cpp (our pybind registration), into

  oracle/_ref/ref_raster_cpu.so    CPU only  (imports anywhere; used by CPU tests + cpu_baseline)
  oracle/_ref/ref_raster_cuda.so   CPU+CUDA  (the reference's own kernels recompiled for sm_100a;
                                              the bit-exactness oracle on the B200 box)

Compiler flags follow the reference's setup.py:52,75-90 (c++17, no fast-math, default -fmad).
oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.

Usage:  python oracle/build_ref.py [--cpu-only] [--force]
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("PYTORCH3D_REFERENCE", "/root/reference")
CSRC = os.path.join(REF, "pytorch3d", "csrc")

CPU_SOURCES = [
    os.path.join(CSRC, "rasterize_meshes", "rasterize_meshes_cpu.cpp"),
    os.path.join(CSRC, "rasterize_points", "rasterize_points_cpu.cpp"),
    os.path.join(CSRC, "compositing", "alpha_composite_cpu.cpp"),
    os.path.join(CSRC, "compositing", "weighted_sum_cpu.cpp"),
    os.path.join(CSRC, "compositing", "norm_weighted_sum_cpu.cpp"),
]
CUDA_SOURCES = [
    os.path.join(CSRC, "rasterize_meshes", "rasterize_meshes.cu"),
    os.path.join(CSRC, "rasterize_coarse", "rasterize_coarse.cu"),
    os.path.join(CSRC, "rasterize_points", "rasterize_points.cu"),
    os.path.join(CSRC, "compositing", "alpha_composite.cu"),
    os.path.join(CSRC, "compositing", "weighted_sum.cu"),
    os.path.join(CSRC, "compositing", "norm_weighted_sum.cu"),
    os.path.join(CSRC, "interp_face_attrs", "interp_face_attrs.cu"),
]
SHIM = os.path.join(HERE, "ref_shim.cpp")


def reference_present():
    return all(os.path.exists(p) for p in CPU_SOURCES + CUDA_SOURCES)


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths(device_type="cuda")
    lib = ce.library_paths(device_type="cuda")
    return torch, inc, lib


def _run(cmd):
    t0 = time.time()
    print("[build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    print("[build_ref]   %.1fs" % (time.time() - t0), flush=True)


def _common(name, inc):
    flags = ["-DTORCH_EXTENSION_NAME=%s" % name, "-DTORCH_API_INCLUDE_EXTENSION_H",
             "-I" + CSRC, "-I" + sysconfig.get_paths()["include"]]
    for p in inc:
        flags += ["-isystem", p]
    return flags


def build(cpu_only=False, force=False):
    if not reference_present():
        print("[build_ref] reference sources not found under %s -- nothing to do" % REF)
        return False
    os.makedirs(OUT, exist_ok=True)
    torch, inc, lib = _torch_paths()
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    ldflags = []
    for p in lib:
        ldflags += ["-L" + p, "-Wl,-rpath," + p]
    ldflags += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]

    def stale(target, srcs):
        if force or not os.path.exists(target):
            return True
        return any(os.path.getmtime(s) > os.path.getmtime(target) for s in srcs)

    # ---------------- CPU-only module ----------------
    name = "ref_raster_cpu"
    target = os.path.join(OUT, name + ".so")
    if stale(target, CPU_SOURCES + [SHIM]):
        objs = []
        for i, src in enumerate(CPU_SOURCES + [SHIM]):
            obj = os.path.join(OUT, "%s_%d.o" % (name, i))
            _run(["g++", "-O2", "-fPIC", "-std=c++17", abi, "-c", src, "-o", obj] + _common(name, inc))
            objs.append(obj)
        _run(["g++", "-shared", "-o", target] + objs + ldflags)
    if cpu_only:
        return True

    # ---------------- CPU + CUDA module ----------------
    name = "ref_raster_cuda"
    target = os.path.join(OUT, name + ".so")
    if stale(target, CPU_SOURCES + CUDA_SOURCES + [SHIM]):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        objs = []
        for i, src in enumerate(CPU_SOURCES + [SHIM]):
            obj = os.path.join(OUT, "%s_%d.o" % (name, i))
            _run(["g++", "-O2", "-fPIC", "-std=c++17", abi, "-DWITH_CUDA", "-c", src, "-o", obj]
                 + _common(name, inc))
            objs.append(obj)
        for i, src in enumerate(CUDA_SOURCES):
            obj = os.path.join(OUT, "%s_cu%d.o" % (name, i))
            _run([nvcc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", abi, "-DWITH_CUDA",
                  "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                  "-D__CUDA_NO_HALF2_OPERATORS__", "-DTHRUST_IGNORE_CUB_VERSION_CHECK",
                  "--expt-relaxed-constexpr",
                  "-gencode", "arch=compute_100a,code=sm_100a", "-c", src, "-o", obj] + _common(name, inc))
            objs.append(obj)
        _run(["g++", "-shared", "-o", target] + objs + ldflags
             + ["-L/usr/local/cuda/lib64", "-lcudart", "-lc10_cuda", "-ltorch_cuda"])
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(cpu_only=a.cpu_only, force=a.force)
    sys.exit(0 if ok else 1)
