#!/usr/bin/env python
"""Triangle count of the displacement-tessellated config-5 mesh for a few base tessellations (device tessellator)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from izpi_b200 import cuda, scenes
from izpi_b200.build import build
build()
ctx = cuda.Context(0)
px = scenes.height_map(4096, 2048)
step = max(np.abs(np.diff(px[..., 2], axis=0)).max(), np.abs(np.diff(px[..., 2], axis=1)).max())
for na, nt, frac in [(128, 64, 0.9), (160, 80, 0.9), (192, 96, 0.9), (160, 80, 0.6)]:
    verts, uvs = scenes.torus_mesh(na, nt, centre=(0, 0, 0), major=600.0, minor=240.0, amp=0.0)
    base = np.concatenate([verts.reshape(-1, 9), uvs.reshape(-1, 6)], axis=1)
    rng = frac * 2.0 / step
    t0 = time.perf_counter()
    tris, mats = ctx.apply_displacement(base, np.zeros(len(base), np.int32), px, -rng / 2, rng / 2)
    print(na, nt, frac, "range", round(rng, 1), "base", len(base), "out", len(tris), "seconds", round(time.perf_counter() - t0, 3), flush=True)
