"""The C-ABI library loads on a CPU-only machine and exports every symbol include/b200_raster.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200_raster.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200r_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("b200r_rasterize_meshes_forward", "b200r_rasterize_meshes_backward", "b200r_rasterize_points_forward",
              "b200r_rasterize_points_backward", "b200r_rasterize_meshes_forward_host",
              "b200r_rasterize_points_forward_host", "b200r_last_error", "b200r_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in _declared_symbols():
        assert hasattr(lib, s), "libb200raster.so does not export %s" % s


def test_ctypes_prototypes_cover_the_header(built_lib):
    from pytorch3d_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.b200r_version().startswith(b"b200raster")
    assert lib.b200r_kernel_launch_count() == 0


def test_workspace_size_query_is_pure_host_code(built_lib):
    from pytorch3d_b200 import _lib
    lib = _lib.load()
    small = lib.b200r_rasterize_meshes_workspace_bytes(1000, 2, 64, 64, 0)
    big = lib.b200r_rasterize_meshes_workspace_bytes(100000, 8, 512, 512, 0)
    assert 0 < small < big
    explicit = lib.b200r_rasterize_meshes_workspace_bytes(1000, 2, 64, 64, 12345)
    assert explicit != small
    assert lib.b200r_rasterize_points_workspace_bytes(1000, 2, 64, 64, 0) > 0


def test_sass_contains_tma_bulk_copy(built_lib):
    """The setup kernels stage packed face_verts with cp.async.bulk (SASS: UBLKCP)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], stdout=subprocess.PIPE, text=True).stdout
    assert "UBLKCP" in sass
    assert "sm_100a" in sass or "SM100" in sass.upper() or "EF_CUDA_SM100" in sass
