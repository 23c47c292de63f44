"""Config 3 with alpha compositing: time of every piece (rasterizer, unfused chain, fused op), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, synthetic  # noqa: E402

dev = torch.device("cuda:0")
pc = synthetic.random_pointclouds(8, 100000, seed=0)
pts = pc.points_packed().to(dev)
pf, pn = pc.cloud_to_packed_first_idx().to(dev), pc.num_points_per_cloud().to(dev)
r = 0.01
rad = torch.full((pts.shape[0],), r, device=dev)
feats = torch.rand(pts.shape[0], 4, device=dev).permute(1, 0)
feats_planar = feats.contiguous()
idx, zb, d2 = _C.rasterize_points(pts, pf, pn, (512, 512), rad, 10, 0, 0)
g_img = torch.randn(8, 4, 512, 512, device=dev)
g_z, g_d = torch.randn_like(zb), torch.randn_like(d2)
w = (1 - d2 / (r * r)).permute(0, 3, 1, 2)
il = idx.long().permute(0, 3, 1, 2)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


rows = [
    ("rasterize_points forward", lambda: _C.rasterize_points(pts, pf, pn, (512, 512), rad, 10, 0, 0)),
    ("rasterize_points backward", lambda: _C.rasterize_points_backward(pts, idx, g_z, g_d)),
    ("torch: weights = 1 - d / r^2", lambda: (1 - d2 / (r * r))),
    ("torch: idx.long()", lambda: idx.long()),
    ("accum_alphacomposite (point-major features)", lambda: _C.accum_alphacomposite(feats, w, il)),
    ("accum_alphacomposite (planar features)", lambda: _C.accum_alphacomposite(feats_planar, w, il)),
    ("accum_alphacomposite_backward (point-major)", lambda: _C.accum_alphacomposite_backward(g_img, feats, w, il)),
    ("accum_alphacomposite_backward (planar)", lambda: _C.accum_alphacomposite_backward(g_img, feats_planar, w, il)),
    ("points_alpha_render (point-major)", lambda: _C.points_alpha_render(feats, idx, d2, r)),
    ("points_alpha_render (planar)", lambda: _C.points_alpha_render(feats_planar, idx, d2, r)),
    ("points_alpha_render_backward (point-major)", lambda: _C.points_alpha_render_backward(g_img, feats, idx, d2, r)),
    ("points_alpha_render_backward (planar)", lambda: _C.points_alpha_render_backward(g_img, feats_planar, idx, d2, r)),
]
for name, fn in rows:
    print("%-48s %8.1f us" % (name, t(fn)), flush=True)
