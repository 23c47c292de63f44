"""The setup pass turns a face's float box into an EXACT rectangle of pixels (`exact_pixel_range`,
pytorch3d_b200/csrc/raster_meshes.cu): the inverse pixel-centre map in plain float locates each end to within a
margin, and only ends that have a pixel centre inside the margin are settled by evaluating pix_to_ndc itself.
This restates that logic with numpy float32 arithmetic (with and without the FMA contraction nvcc may apply to the
inverse map) and checks it against brute force on adversarial boxes whose edges sit on, or a few ulps away from,
pixel centres -- i.e. it tests the margin assumption the device code relies on."""
import numpy as np
import pytest

f32 = np.float32


def fma32(a, b, c):
    # a*b is exact in float64 (24+24 significant bits); the sum is rounded once to float64 and once to float32 --
    # a double rounding that can differ from a true fma only in astronomically rare ties
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def pix_to_ndc(i, S, rng):
    # raster_math.cuh::pix_to_ndc: fsub(fdiv(ffma(range, i, offset), S), offset)
    off = f32(rng * f32(0.5))
    return f32(f32(fma32(rng, f32(i), off) / f32(S)) - off)


def exact_pixel_range(vmin, vmax, S, rng, contract):
    off, scale = f32(rng * f32(0.5)), f32(f32(S) / rng)
    margin = f32(f32(1e-3) + f32(f32(1e-6) * f32(S)))
    if contract:
        a = fma32(f32(vmin + off), scale, f32(-0.5))
        b = fma32(f32(vmax + off), scale, f32(-0.5))
    else:
        a = f32(f32(f32(vmin + off) * scale) - f32(0.5))
        b = f32(f32(f32(vmax + off) * scale) - f32(0.5))
    clampa = lambda v: min(max(v, f32(-1.0)), f32(S) + f32(1.0))
    clampb = lambda v: min(max(v, f32(-2.0)), f32(S))
    lo = max(0, int(np.ceil(clampa(f32(a - margin)))))
    hi = min(S - 1, int(np.floor(clampb(f32(b + margin)))))
    lo_sure = lo == max(0, int(np.ceil(clampa(f32(a + margin)))))
    hi_sure = hi == min(S - 1, int(np.floor(clampb(f32(b - margin)))))
    passes = lambda i: not (pix_to_ndc(i, S, rng) > vmax or pix_to_ndc(i, S, rng) < vmin)
    if not lo_sure:
        while lo <= hi and not passes(lo):
            lo += 1
    if not hi_sure:
        while hi >= lo and not passes(hi):
            hi -= 1
    return lo, hi


def brute(vmin, vmax, S, rng):
    idx = [i for i in range(S) if not (pix_to_ndc(i, S, rng) > vmax or pix_to_ndc(i, S, rng) < vmin)]
    if not idx:
        return None
    assert idx == list(range(idx[0], idx[-1] + 1)), "passing pixels must be contiguous"
    return idx[0], idx[-1]


def nudge(v, ulps):
    v = f32(v)
    for _ in range(abs(ulps)):
        v = np.nextafter(v, f32(np.inf) if ulps > 0 else f32(-np.inf), dtype=np.float32)
    return v


@pytest.mark.parametrize("S,other", [(16, 16), (17, 100), (100, 17), (512, 512), (2048, 1024), (64, 2048)])
def test_exact_pixel_range_matches_brute_force(S, other):
    rng = f32(2.0) if S <= other else f32(f32(S) * f32(2.0) / f32(other))  # ndc_range(S, other)
    g = np.random.default_rng(S * 7919 + other)
    cases = []
    for _ in range(300):  # edges exactly on / a few ulps around pixel centres
        i, j = sorted(g.integers(0, S, 2))
        cases.append((nudge(pix_to_ndc(i, S, rng), int(g.integers(-3, 4))),
                      nudge(pix_to_ndc(j, S, rng), int(g.integers(-3, 4)))))
    for _ in range(300):  # arbitrary boxes, also partly or wholly outside the image, also empty
        lo = f32(g.uniform(-1.3, 1.3) * float(rng) / 2)
        cases.append((lo, f32(lo + f32(abs(g.normal(0, 0.02)) * float(rng)))))
    for _ in range(50):  # sub-pixel boxes between two centres
        i = int(g.integers(0, S - 1))
        a, b = pix_to_ndc(i, S, rng), pix_to_ndc(i + 1, S, rng)
        cases.append((nudge(a, 1), nudge(b, -1)))
    for vmin, vmax in cases:
        if vmin > vmax:
            vmin, vmax = vmax, vmin
        want = brute(vmin, vmax, S, rng)
        for contract in (False, True):
            lo, hi = exact_pixel_range(vmin, vmax, S, rng, contract)
            if want is None:
                assert lo > hi, (vmin, vmax, lo, hi)
            else:
                assert (lo, hi) == want, (vmin, vmax, lo, hi, want, contract)
