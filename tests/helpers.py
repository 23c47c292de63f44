"""Shared scene builders for the tests (seeded, tiny)."""
import numpy as np
import torch


def rand_faces(F, N, seed, scale=0.2, zlo=0.5, zhi=3.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(F, 1, 3, generator=g) * 2 - 1
    v = c + (torch.rand(F, 3, 3, generator=g) - 0.5) * scale * 2
    v[..., 2] = zlo + (zhi - zlo) * torch.rand(F, 3, generator=g)
    first, num = split(F, N)
    return v.contiguous(), first, num


def split(E, N):
    per = E // N
    first = (torch.arange(N) * per).long()
    num = torch.full((N,), per).long()
    num[-1] = E - first[-1]
    return first, num


def rand_points(P, N, seed, rlo=0.03, rhi=0.15, z_ties=False):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(P, 3, generator=g) * 2 - 1
    pts[:, 2] = torch.rand(P, generator=g) * 2 - 0.2
    if z_ties:
        pts[::5, 2] = 0.5
    rad = torch.rand(P, generator=g) * (rhi - rlo) + rlo
    first, num = split(P, N)
    return pts.contiguous(), first, num, rad.contiguous()


def upstream(shapes, seed=231):
    """Seeded upstream gradients (the reference's own seed, tests/test_rasterize_meshes.py:563)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


def assert_frag_equal(a, b, what=""):
    """idx bit-exact; floats bit-exact as well (same arithmetic) unless tol is given."""
    a = [x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x) for x in a]
    b = [x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x) for x in b]
    assert np.array_equal(a[0], b[0]), "%s: index mismatch in %d slots" % (what, int((a[0] != b[0]).sum()))
    for i, (x, y) in enumerate(zip(a[1:], b[1:])):
        assert np.array_equal(x, y), "%s: float output %d differs, max abs %g" % (
            what, i, float(np.abs(x.astype(np.float64) - y).max()))
