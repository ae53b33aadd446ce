#!/usr/bin/env python
"""Batch-size sweep (paths per batch) on render configs; scene built once per config."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from izpi_b200 import cuda, render, scenes
from izpi_b200 import scene as S
from izpi_b200.build import build
build()
cfgs = sys.argv[1].split(",")
for c in cfgs:
    ctx0 = cuda.Context(0)
    if c == "5":
        sc, w, h, spp, sampler = scenes.ibl_tessellated_mesh(ctx0, 3840 / 2160)[0], 3840, 2160, 64, cuda.SAMPLER_COLOUR
    elif c == "5l":
        sc, w, h, spp, sampler = scenes.ibl_tessellated_mesh(ctx0, 3840 / 2160, bvh_builder=S.BVH_DEVICE_LBVH)[0], 3840, 2160, 64, cuda.SAMPLER_COLOUR
    elif c == "4":
        sc, w, h, spp, sampler = scenes.spectral_pyramid(1.0), 1024, 1024, 128, cuda.SAMPLER_SPECTRAL
    elif c == "3":
        sc, w, h, spp, sampler = scenes.cornell_pbr_mesh(1.0), 1024, 1024, 128, cuda.SAMPLER_COLOUR
    hs = cuda.HostScene(sc)
    ctx0.close()
    for bp in (24, 25, 26, 24):
        os.environ["IZPI_BATCH_PATHS"] = str(1 << bp)
        ctx = cuda.Context(0)
        ctx.upload(hs)
        render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=3, sample_count=1).Render()
        r = render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=3)
        r.canvas()
        t0 = time.perf_counter()
        r.Render()
        dt = time.perf_counter() - t0
        print(json.dumps({"config": c, "batch_log2": bp, "spp": spp, "seconds": dt, "msamples_per_s": w * h * spp / dt / 1e6}), flush=True)
        ctx.close()
