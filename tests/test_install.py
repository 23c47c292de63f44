"""install() rebinds `_C` inside PyTorch3D's wrapper modules (rasterizers, compositing, interp_face_attrs); checked
here against stand-in modules (tests/test_gpu_modules.py runs it against the real package on the GPU box)."""
import sys
import types

import torch


def _fake_pytorch3d(monkeypatch):
    calls = []
    orig = types.SimpleNamespace(
        rasterize_meshes=lambda *a, **k: calls.append("ref_meshes") or "ref",
        rasterize_meshes_backward=lambda *a, **k: "ref_bwd",
        rasterize_points=lambda *a, **k: "ref_pts",
        rasterize_points_backward=lambda *a, **k: "ref_pts_bwd",
        knn_points_idx=lambda *a, **k: "untouched",
    )
    names = ["pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.mesh", "pytorch3d.renderer.mesh.rasterize_meshes",
             "pytorch3d.renderer.points", "pytorch3d.renderer.points.rasterize_points",
             "pytorch3d.renderer.compositing", "pytorch3d.ops", "pytorch3d.ops.interp_face_attrs"]
    for n in names:
        m = types.ModuleType(n)
        m.__path__ = []
        monkeypatch.setitem(sys.modules, n, m)
    sys.modules["pytorch3d.renderer.mesh.rasterize_meshes"]._C = orig
    sys.modules["pytorch3d.renderer.points.rasterize_points"]._C = orig
    sys.modules["pytorch3d.renderer.compositing"]._C = orig
    sys.modules["pytorch3d.ops.interp_face_attrs"]._C = orig
    return orig, calls


def test_install_and_uninstall(monkeypatch, built_lib):
    from pytorch3d_b200 import install as inst
    orig, calls = _fake_pytorch3d(monkeypatch)
    patched = inst.install()
    assert len(patched) == 4
    rm = sys.modules["pytorch3d.renderer.mesh.rasterize_meshes"]
    assert rm._C is not orig
    # CPU tensors keep the reference's CPU implementation; unrelated ops pass through untouched
    assert rm._C.rasterize_meshes(torch.zeros(1, 3, 3)) == "ref" and calls == ["ref_meshes"]
    assert rm._C.knn_points_idx() == "untouched"
    # CUDA tensors are routed to pytorch3d_b200._C (here: a stand-in object that claims to be on the GPU)
    routed = []
    monkeypatch.setattr(inst._b200_C, "rasterize_points", lambda *a, **k: routed.append(a) or "b200")
    fake_cuda = types.SimpleNamespace(is_cuda=True)
    assert sys.modules["pytorch3d.renderer.points.rasterize_points"]._C.rasterize_points(fake_cuda) == "b200"
    assert routed
    inst.uninstall()
    assert rm._C is orig
