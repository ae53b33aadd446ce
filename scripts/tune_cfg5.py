#!/usr/bin/env python
"""Parameter sweeps on the config-5 frame (scene built once): pair-straggler threshold, batch size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from izpi_b200 import cuda, render, scenes
from izpi_b200 import scene as S
from izpi_b200.build import build
build()
lbvh = "--lbvh" in sys.argv
spp = 16
ctx0 = cuda.Context(0)
sc = scenes.ibl_tessellated_mesh(ctx0, 3840 / 2160, bvh_builder=S.BVH_DEVICE_LBVH if lbvh else S.BVH_REFERENCE)[0]
hs = cuda.HostScene(sc)
ctx0.close()
for env in [{}, {"IZPI_PAIR_STRAGGLERS": "5"}, {"IZPI_PAIR_STRAGGLERS": "9"}, {"IZPI_PAIR_STRAGGLERS": "11"}, {"IZPI_BATCH_PATHS": str(1 << 25)},
            {"IZPI_BATCH_PATHS": str(1 << 23)}, {"IZPI_TRACE_LANES": "4"}]:
    for k in ("IZPI_PAIR_STRAGGLERS", "IZPI_BATCH_PATHS", "IZPI_TRACE_LANES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = cuda.Context(0)
    ctx.upload(hs)
    render.New(ctx, 3840, 2160, 1, 50, seed=3).Render()
    r = render.New(ctx, 3840, 2160, spp, 50, seed=3)
    t0 = time.perf_counter()
    r.Render()
    dt = time.perf_counter() - t0
    print(json.dumps({"env": env, "lbvh": lbvh, "seconds": dt, "msamples_per_s": 3840 * 2160 * spp / dt / 1e6}), flush=True)
    ctx.close()
