"""GPU time of one end-to-end step (public API forward + torch loss + backward) by kernel, via torch.profiler."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorch3d_b200 import rasterize_meshes  # noqa: E402

dev = torch.device("cuda:0")
meshes, (nm, F1, H, W, K, blur) = bench.build_workload("ns", 0)
verts_h, faces_h = meshes.verts_packed().pin_memory(), meshes.faces_packed().pin_memory()
first, num = meshes.mesh_to_faces_packed_first_idx().to(dev), meshes.num_faces_per_mesh().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
gz = torch.randn(nm, H, W, K, device=dev, generator=g)
gb = torch.randn(nm, H, W, K, 3, device=dev, generator=g)
gd = torch.randn(nm, H, W, K, device=dev, generator=g)
grad_h = torch.empty_like(verts_h).pin_memory()


def step():
    v = verts_h.to(dev, non_blocking=True).requires_grad_(True)
    f = faces_h.to(dev, non_blocking=True)
    m = bench._DeviceMeshes(v, f, first, num, F1)
    p2f, zbuf, bary, dists = rasterize_meshes(m, (H, W), blur_radius=blur, faces_per_pixel=K)
    loss = torch.dot(zbuf.reshape(-1), gz.reshape(-1)) + torch.dot(bary.reshape(-1), gb.reshape(-1)) + \
        torch.dot(dists.reshape(-1), gd.reshape(-1))
    loss.backward()
    grad_h.copy_(v.grad, non_blocking=True)
    return float(loss)


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print("GPU time per step: %.1f us" % (total / n))
for e in rows[:18]:
    print("%8.1f us/step  %5.1f%%  x%-3d %s" % (e.device_time_total / n, 100 * e.device_time_total / total, e.count // n,
                                              e.key[:90]))
