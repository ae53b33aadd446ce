#!/usr/bin/env python
"""Tiny invocations of every device path, for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`:
closest hit (2-lane, 4-lane, thread-per-ray), renders over a slice world, a tiny BVH with the coherence sort forced on, a
multi-material mesh (per-material bins), the spectral epilogue, tile rows, scene image adopt, device BVH build."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["IZPI_SORT_RAYS"] = "2"  # sort every bounce that has at least two live paths
from izpi_b200 import cuda, render, scenes  # noqa: E402
from izpi_b200 import scene as S  # noqa: E402

ctx = cuda.Context(0)
verts, uvs = scenes.torus_mesh(60, 30)
sc = S.SceneSpec(bvh_seed=12345)
sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
org, d = scenes.random_rays(5000, lo, hi)
ctx.upload(cuda.HostScene(sc))
ids, t = ctx.trace_closest(org, d)
ids2, t2, st = ctx.trace_closest(org, d, stats=True)
assert np.array_equal(ids, ids2)
sc.bvh_builder = S.BVH_DEVICE_LBVH
ctx.upload(cuda.HostScene(sc))
ids3, t3 = ctx.trace_closest(org, d)
assert np.array_equal(ids, ids3) and t.tobytes() == t3.tobytes()
print("trace ok", (ids >= 0).mean())
ctx.upload(cuda.HostScene(scenes.cornell_box(1.0)))
img, rays = ctx.render(40, 40, 4, seed=3)
ctx.upload(cuda.HostScene(scenes.spectral_pyramid(1.0)))
img, rays = ctx.render(32, 32, 16, sampler=cuda.SAMPLER_SPECTRAL, seed=3)
ctx.render_setup(32, 32, 4, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
rows = ctx.render_tile_rows(0, 0, 31, 31)
print("spectral ok", float(np.abs(img[..., :3]).max()))
pm = scenes.cornell_pbr_mesh(1.0, n_around=48, n_tube=24, tex_size=32, n_pbr_materials=4)
ctx.upload(cuda.HostScene(pm))
r = render.New(ctx, 32, 32, 4, 50, seed=5, stats=cuda.RENDER_STATS)
r.Render()
print("bins", ctx.render_stats()["material_bins"])
header, sizes, ptrs = ctx.scene_image()
other = cuda.Context(0)
mine = other.scene_adopt(header, len(sizes))
import torch  # noqa: E402
for a, b, n in zip(mine, ptrs, sizes):
    if n:
        torch.as_tensor(render._DevicePtr(a, (n,), "|u1"), device="cuda").copy_(torch.as_tensor(render._DevicePtr(b, (n,), "|u1"), device="cuda"))
torch.cuda.synchronize()
other.scene_commit()
a, _ = ctx.render(32, 32, 2, seed=1)
b, _ = other.render(32, 32, 2, seed=1)
assert a.tobytes() == b.tobytes()
other.close()
ctx.close()
print("SANITIZE_SMALL_OK")
