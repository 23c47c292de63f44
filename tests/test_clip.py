"""Frustum culling / z-clipping: clip_faces against fixtures produced by the reference's clip.py,
and the end-to-end clipped rasterization against the reference's CPU render."""
import numpy as np
import pytest
import torch

from pytorch3d_b200 import clip as mclip


def _cases(golden):
    return sorted(k for k in golden if k.startswith("clip/"))


def _frustum(c):
    persp, cull, has_z = (int(v) for v in c["args"])
    zc = float(c["z_clip"][0]) if has_z > 0 else None
    return persp, cull, zc


def test_clip_faces_matches_reference_fixtures(golden):
    names = _cases(golden)
    assert len(names) == 6
    for name in names:
        c = golden[name]
        persp, cull, zc = _frustum(c)
        fr = mclip.ClipFrustum(left=-1, right=1, top=-1, bottom=1, perspective_correct=bool(persp), z_clip_value=zc,
                               cull=bool(cull))
        out = mclip.clip_faces(torch.from_numpy(c["face_verts"]), torch.from_numpy(c["first"]),
                               torch.from_numpy(c["num"]), fr)
        assert np.array_equal(out.face_verts.numpy(), c["out_face_verts"]), name
        assert np.array_equal(out.mesh_to_face_first_idx.numpy(), c["out_first"]), name
        assert np.array_equal(out.num_faces_per_mesh.numpy(), c["out_num"]), name
        assert np.array_equal(out.faces_clipped_to_unclipped_idx.numpy(), c["out_c2u"]), name
        if "out_neighbor" in c:
            assert np.array_equal(out.clipped_faces_neighbor_idx.numpy(), c["out_neighbor"]), name
            nb = out.clipped_faces_neighbor_idx
            pair = nb >= 0
            assert torch.equal(nb[nb[pair]], torch.arange(len(nb))[pair])  # the two halves name each other


def test_clip_nothing_to_do_returns_inputs():
    fv = torch.rand(10, 3, 3) + torch.tensor([0.0, 0.0, 1.0])
    first, num = torch.tensor([0]), torch.tensor([10])
    out = mclip.clip_faces(fv, first, num, mclip.ClipFrustum(z_clip_value=0.5, cull=False))
    assert out.face_verts is fv and out.faces_clipped_to_unclipped_idx is None
    p2f, bary = torch.zeros(1, 2, 2, 1, dtype=torch.int64), torch.rand(1, 2, 2, 1, 3)
    a, b = mclip.convert_clipped_rasterization_to_original_faces(p2f, bary, out)
    assert a is p2f and b is bary


def test_clip_gradients_are_finite_and_flow_to_kept_vertices():
    g = torch.Generator().manual_seed(0)
    fv = (torch.rand(50, 3, 3, generator=g) * 2 - 1)
    fv[..., 2] = torch.rand(50, 3, generator=g) * 2 - 0.5
    fv[0, :, 2] = 1.0  # a face with three equal depths (zero denominators in the unused rows)
    fv.requires_grad_(True)
    out = mclip.clip_faces(fv, torch.tensor([0]), torch.tensor([50]),
                           mclip.ClipFrustum(perspective_correct=True, z_clip_value=0.2, cull=False))
    (out.face_verts ** 2).sum().backward()
    assert torch.isfinite(fv.grad).all() and fv.grad.abs().sum() > 0


@pytest.mark.gpu
def test_clipped_rasterization_end_to_end(golden, built_lib):
    """rasterize_meshes(z_clip_value, cull_to_frustum) on the GPU against the reference's CPU render of the same
    scene: exercises clip_faces on CUDA tensors, the kernel variant with the clipped-neighbour logic, and the
    conversion back to the unclipped faces."""
    import pytorch3d_b200 as p3b
    dev = torch.device("cuda:0")
    for name in _cases(golden):
        c = golden[name]
        persp, cull, zc = _frustum(c)
        fv = torch.from_numpy(c["face_verts"])
        verts = [fv[:120].reshape(-1, 3).to(dev), torch.zeros(0, 3, device=dev), fv[120:].reshape(-1, 3).to(dev)]
        faces = [torch.arange(360, device=dev).reshape(-1, 3), torch.zeros(0, 3, dtype=torch.int64, device=dev),
                 torch.arange(540, device=dev).reshape(-1, 3)]
        meshes = p3b.PackedMeshes(verts, faces)
        out = p3b.rasterize_meshes(meshes, (24, 32), 1e-3, 4, None, None, bool(persp), False, False, zc, bool(cull))
        want = [c["e2e_pix_to_face"], c["e2e_zbuf"], c["e2e_bary"], c["e2e_dists"]]
        got = [o.cpu().numpy() for o in out]
        # the fixture comes from the reference's CPU arithmetic / CPU queue rule: identical indices except where a
        # pixel sits exactly on a razor edge (FMA vs no FMA); demand >= 99.9% identical slots and tight floats there
        same = got[0] == want[0]
        assert same.mean() >= 0.999, name
        for oi, (g_, w_) in enumerate(zip(got[1:], want[1:])):
            m = same if g_.ndim == 4 else same[..., None].repeat(3, -1)
            if oi == 2:
                # dists: a pixel whose nearest point lies on the edge shared by the two halves of a clipped quad is
                # equally far from both; which half survives (and hence the SIGN of the distance) is decided by the
                # last bit of `dist < neighbor_dist`, which FMA and non-FMA arithmetic round differently (the
                # reference's own CPU and CUDA builds disagree there too) -> compare magnitudes
                g_, w_ = np.abs(g_), np.abs(w_)
            err = np.abs(g_ - w_)[m]
            # (perspective-corrected values extrapolated far outside a face are ill-conditioned: allow 0.5% of
            #  the entries to exceed the tolerance)
            assert (err <= 1e-5 + 1e-3 * np.abs(w_[m])).mean() >= 0.995, name
