#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/).

The reference (Go) cannot run in this environment, so these are ORACLE outputs, not reference outputs:
they pin today's behaviour of the restatement (already checked against the reference's own unit vectors in
tests/test_oracle_golden.py) so that the -m gpu tests can also run against committed data.

  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from izpi_b200 import scene as S  # noqa: E402
from izpi_b200 import scenes  # noqa: E402
from test_oracle_golden import box_random_case  # noqa: E402


def main():
    # 1. the reference's 1000 deterministic box cases (bvh4_simd_test.go:200-268): masks of the SSE flavour
    masks = np.array([oracle.ray_aabb4(0, *box_random_case(i)) for i in range(1000)], dtype=np.uint8)
    np.save(os.path.join(HERE, "box_masks_1000.npy"), masks)
    # 2. closest hit on a small displaced torus (2000 triangles, BVH4 seed 12345), 4096 incoherent rays
    verts, uvs = scenes.torus_mesh(40, 25)
    sc = S.SceneSpec(bvh_seed=12345)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org, d = scenes.random_rays(4096, lo, hi)
    osn = oracle.OracleScene(sc)
    ids, t, st = osn.trace(org, d, stats=True)
    nodes, perm = osn.bvh()
    np.savez_compressed(os.path.join(HERE, "torus_40x25_closest_hit.npz"), ids=ids, t=t, nodes_visited=st["nodes"], prim_tests=st["tris"],
                        nodes_sha256=hashlib.sha256(nodes.tobytes()).hexdigest(), perm_sha256=hashlib.sha256(perm.tobytes()).hexdigest())
    # 3. config 4 geometry (triangles + spheres in one BVH4): closest hit
    sc4 = scenes.spectral_pyramid()
    o4, d4 = scenes.random_rays(4096, (0, 0, 0), (100, 100, 100), seed=4)
    i4, t4 = oracle.OracleScene(sc4).trace(o4, d4)
    np.savez_compressed(os.path.join(HERE, "pyramid_closest_hit.npz"), ids=i4, t=t4)
    # 4. renders with the device-matching counter RNG (rng_mode=1)
    img1, rays1 = oracle.OracleScene(scenes.cornell_box(1.0)).render(32, 32, 4, rng_mode=1, seed=5)
    img4, rays4 = oracle.OracleScene(sc4).render(32, 32, 8, sampler=1, rng_mode=1, seed=5)
    np.savez_compressed(os.path.join(HERE, "renders_32x32.npz"), cornell=img1, cornell_rays=rays1, pyramid=img4, pyramid_rays=rays4)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
