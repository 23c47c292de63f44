"""Host-side logic that needs no GPU: argument checking, wrappers, containers, generators."""
import math

import numpy as np
import pytest
import torch

import pytorch3d_b200 as p3b
from pytorch3d_b200 import _C, synthetic
from pytorch3d_b200.rasterize_meshes import parse_image_size, pix_to_non_square_ndc
from pytorch3d_b200.rasterize_points import _format_radius


def test_cpu_tensors_fail_loudly(built_lib):
    m = synthetic.torus_batch(1, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        p3b.rasterize_meshes(m, 32)
    pc = synthetic.random_pointclouds(1, 10)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        p3b.rasterize_points(pc, 32, radius=0.1)


def test_bin_size_error_matches_reference():
    # reference: tests/test_rasterize_meshes.py:461-466 / rasterize_meshes.py:212-219
    m = synthetic.torus_batch(1, 8, 8)
    with pytest.raises(ValueError, match="bin_size too small"):
        p3b.rasterize_meshes(m, 128, 0.0, 2, bin_size=2)
    pc = synthetic.random_pointclouds(1, 10)
    with pytest.raises(ValueError, match="bin_size too small"):
        p3b.rasterize_points(pc, 128, 0.1, 2, bin_size=2)


def test_shape_and_limit_errors(built_lib):
    fv = torch.zeros(4, 3, 2)
    z = torch.zeros(1, dtype=torch.int64)
    with pytest.raises(RuntimeError, match=r"face_verts must have dimensions \(num_faces, 3, 3\)"):
        _C.rasterize_meshes(fv, z, z, torch.zeros(4, dtype=torch.int64), (8, 8), 0.0, 1, 0, 0, False, False, False)
    fv = torch.zeros(4, 3, 3)
    with pytest.raises(RuntimeError, match="Must have points_per_pixel <= 150"):
        _C.rasterize_meshes(fv, z, z, torch.full((4,), -1), (8, 8), 0.0, 151, 0, 0, False, False, False)
    with pytest.raises(RuntimeError, match="clipped_faces_neighbor_idx"):
        _C.rasterize_meshes(fv, z, z, torch.full((3,), -1), (8, 8), 0.0, 1, 0, 0, False, False, False)
    with pytest.raises(RuntimeError, match="Must have num_closest <= 150"):
        _C.rasterize_points(torch.zeros(4, 3), z, z, (8, 8), torch.zeros(4), 151, 0, 0)
    # the fused (verts, faces) entry point validates like the ops it replaces
    verts, faces = torch.zeros(5, 3), torch.zeros(4, 3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match=r"verts_packed must have dimensions \(num_verts, 3\)"):
        _C.rasterize_meshes_indexed(torch.zeros(5, 2), faces, z, z, (8, 8), 0.0, 1, False, False, False)
    with pytest.raises(RuntimeError, match=r"faces_packed must have dimensions \(num_faces, 3\)"):
        _C.rasterize_meshes_indexed(verts, torch.zeros(4, 4, dtype=torch.int64), z, z, (8, 8), 0.0, 1, False, False, False)
    with pytest.raises(RuntimeError, match="Must have points_per_pixel <= 150"):
        _C.rasterize_meshes_indexed(verts, faces, z, z, (8, 8), 0.0, 151, False, False, False)
    with pytest.raises(RuntimeError, match="expected scalar type Float"):
        _C.rasterize_meshes_indexed(verts.double(), faces, z, z, (8, 8), 0.0, 1, False, False, False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.rasterize_meshes_indexed(verts, faces, z, z, (8, 8), 0.0, 1, False, False, False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.rasterize_meshes_backward_indexed(torch.zeros(4, 3, 3), faces, 5, torch.zeros(1, 8, 8, 1, dtype=torch.int64),
                                             torch.zeros(1, 8, 8, 1), torch.zeros(1, 8, 8, 1, 3),
                                             torch.zeros(1, 8, 8, 1), False, False)


def test_parse_image_size():
    assert parse_image_size(64) == (64, 64)
    assert parse_image_size((32, 48)) == (32, 48)
    with pytest.raises(ValueError):
        parse_image_size((1, 2, 3))
    with pytest.raises(ValueError):
        parse_image_size((0, 5))
    with pytest.raises(ValueError):
        parse_image_size((4.0, 5))


def test_pix_to_ndc_matches_reference_formula():
    assert pix_to_non_square_ndc(0, 4, 4) == pytest.approx(-0.75)
    assert pix_to_non_square_ndc(3, 4, 4) == pytest.approx(0.75)
    # wider than tall: x range is [-2, 2]
    assert pix_to_non_square_ndc(0, 8, 4) == pytest.approx(-1.75)


def test_radius_formats():
    # reference: tests/test_rasterize_points.py:635-660 / rasterize_points.py:145-184
    pc = synthetic.random_pointclouds(2, 5)
    r = _format_radius(0.1, pc)
    assert r.shape == (10,) and torch.allclose(r, torch.full((10,), 0.1))
    r = _format_radius(torch.arange(10, dtype=torch.float32).reshape(2, 5), pc)
    assert torch.equal(r, torch.arange(10, dtype=torch.float32))
    with pytest.raises(ValueError, match="radius must be of shape"):
        _format_radius(torch.zeros(3, 5), pc)
    with pytest.raises(ValueError, match="radius must be a float, list, tuple or tensor"):
        _format_radius(1, pc)


def test_packed_containers():
    v1, f1 = synthetic.torus(4, 3)
    v2, f2 = synthetic.ico_sphere(0)
    m = p3b.PackedMeshes([v1, v2], [f1, f2])
    assert m.num_faces_per_mesh().tolist() == [24, 20]
    assert m.mesh_to_faces_packed_first_idx().tolist() == [0, 24]
    assert m.faces_packed()[24:].min() == v1.shape[0]
    fv = synthetic.face_verts_of(m)
    assert fv.shape == (44, 3, 3)
    pc = p3b.PackedPointclouds([torch.zeros(3, 3), torch.ones(5, 3)])
    assert pc.cloud_to_packed_first_idx().tolist() == [0, 3]
    assert pc.padded_to_packed_idx().tolist() == [0, 1, 2, 5, 6, 7, 8, 9]


def test_synthetic_generators():
    v, f = synthetic.ico_sphere(4)
    assert v.shape == (2562, 3) and f.shape == (5120, 3)  # utils/ico_sphere.py counts
    v, f = synthetic.torus(187, 187)
    assert f.shape[0] == 69938
    m = synthetic.torus_batch(2, 10, 10, seed=1)
    vp = m.verts_packed()
    assert vp[:, :2].abs().max() <= 0.9 + 1e-5 and vp[:, 2].min() >= 1.0 - 1e-5 and vp[:, 2].max() <= 3.0 + 1e-5
    m2 = synthetic.torus_batch(2, 10, 10, seed=1)
    assert torch.equal(vp, m2.verts_packed())
    # every torus edge is shared by exactly two faces (closed manifold)
    fcs = synthetic.torus(6, 5)[1]
    e = torch.cat([fcs[:, [0, 1]], fcs[:, [1, 2]], fcs[:, [2, 0]]]).sort(1).values
    _, counts = torch.unique(e, dim=0, return_counts=True)
    assert (counts == 2).all()


def test_rasterizer_settings_defaults_match_reference():
    s = p3b.RasterizationSettings()
    assert (s.image_size, s.blur_radius, s.faces_per_pixel, s.bin_size, s.cull_backfaces) == (256, 0.0, 1, None, False)
    ps = p3b.PointsRasterizationSettings()
    assert (ps.image_size, ps.radius, ps.points_per_pixel) == (256, 0.01, 8)
    fr = p3b.Fragments(torch.zeros(1), torch.zeros(1, requires_grad=True), torch.zeros(1), None)
    assert not fr.detach().zbuf.requires_grad


def test_wrapper_picks_fused_or_clipping_path(monkeypatch):
    """rasterize_meshes(meshes) hands (verts, faces) to the fused op unless clipping needs face_verts as a tensor;
    outputs that take no part in the loss reach the backward op as zeros, not as None.  (Native ops replaced by
    recording fakes: host logic only.)"""
    import importlib
    rm = importlib.import_module("pytorch3d_b200.rasterize_meshes")  # (the package attribute is the function)
    calls = []
    m = synthetic.torus_batch(2, 6, 6)
    m.requires_grad_(True)
    F, V = m.faces_packed().shape[0], m.verts_packed().shape[0]

    def frags(N, H, W, K):
        return (torch.zeros(N, H, W, K, dtype=torch.int64), torch.zeros(N, H, W, K), torch.zeros(N, H, W, K, 3),
                torch.zeros(N, H, W, K))

    def fake_indexed(verts, faces, first, num, size, blur, K, persp, clip, cull):
        calls.append("indexed")
        return frags(len(num), size[0], size[1], K) + (verts[faces],)

    def fake_indexed_bwd(face_verts, faces, num_verts, p2f, gz, gb, gd, persp, clip):
        calls.append(("indexed_bwd", gz is not None and float(gz.abs().sum()), float(gb.abs().sum()),
                      float(gd.abs().sum())))
        assert gz.shape == p2f.shape and gb.shape == p2f.shape + (3,) and gd.shape == p2f.shape
        return torch.ones(num_verts, 3)

    def fake_face_verts(fv, first, num, nb, size, blur, K, bs, mf, persp, clip, cull):
        calls.append("face_verts")
        return frags(len(num), size[0], size[1], K)

    monkeypatch.setattr(rm._C, "rasterize_meshes_indexed", fake_indexed)
    monkeypatch.setattr(rm._C, "rasterize_meshes_backward_indexed", fake_indexed_bwd)
    monkeypatch.setattr(rm._C, "rasterize_meshes", fake_face_verts)
    p2f, zbuf, bary, dists = rm.rasterize_meshes(m, 8, faces_per_pixel=2)
    assert calls == ["indexed"] and not p2f.requires_grad and zbuf.requires_grad
    (dists * 2.0).sum().backward()  # zbuf and bary unused: their gradients must arrive as zeros
    assert calls[1][0] == "indexed_bwd" and calls[1][1] == 0.0 and calls[1][2] == 0.0 and calls[1][3] > 0.0
    assert torch.equal(m.verts_packed().grad, torch.ones(V, 3))
    calls.clear()
    rm.rasterize_meshes(m, 8, faces_per_pixel=2, z_clip_value=1e-2)
    assert calls == ["face_verts"]
    calls.clear()
    rm.rasterize_meshes(m, 8, faces_per_pixel=2, cull_to_frustum=True)
    assert calls == ["face_verts"]
    assert F > 0
