"""Stores the binary masks of the reference's module-level golden images as packed bits in
tests/golden/module_masks.npz (the PNGs live under /root/reference/tests/data, which does not exist on the GPU box).

Masks, as tests/test_rasterizer.py:48-53 of the reference computes them (pixel == image maximum):
    test_rasterized_sphere_MeshRasterizer.png        ico_sphere(5), FoV camera at 2.7, 512^2   (test_rasterizer.py:67-101)
    test_rasterized_sphere_zoom_MeshRasterizer.png   same, camera at 20                        (:116-133)
    test_simple_pointcloud_sphere.png                ico_sphere(1) verts + 0.2, r = 0.05, 256^2 (:445-492)
Only the images' pixel values are stored; run with:  python tests/golden/make_module_masks.py
"""
import os

import numpy as np
from PIL import Image

DATA = "/root/reference/tests/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "module_masks.npz")
store = {}
for name in ("test_rasterized_sphere_MeshRasterizer", "test_rasterized_sphere_zoom_MeshRasterizer",
             "test_simple_pointcloud_sphere"):
    with Image.open(os.path.join(DATA, name + ".png")) as im:
        a = np.array(im)
    mask = a == a.max()
    if mask.ndim == 3:
        mask = mask[..., 0]  # the reference compares channel 0 (test_rasterizer.py:492)
    store[name + "/shape"] = np.array(mask.shape, np.int64)
    store[name + "/bits"] = np.packbits(mask.astype(np.uint8))
    print(name, mask.shape, int(mask.sum()), "pixels set")
np.savez_compressed(OUT, **store)
print("wrote", OUT)
