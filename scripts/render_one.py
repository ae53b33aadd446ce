#!/usr/bin/env python
"""One frame of one BASELINE render config, for profiling (ncu) and quick timing.

  python scripts/render_one.py --config 4 --spp 4 [--lbvh] [--no-warmup] [--stats]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="4")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--lbvh", action="store_true")
    ap.add_argument("--no-warmup", action="store_true")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    args = ap.parse_args()
    from izpi_b200 import cuda, render, scenes
    from izpi_b200 import scene as S
    from izpi_b200.build import build
    build()
    ctx = cuda.Context(0)
    c = args.config
    if c == "1":
        sc, w, h, sampler = scenes.cornell_box(1.0), 400, 400, cuda.SAMPLER_COLOUR
    elif c == "3":
        sc, w, h, sampler = scenes.cornell_pbr_mesh(1.0), 1024, 1024, cuda.SAMPLER_COLOUR
    elif c == "4":
        sc, w, h, sampler = scenes.spectral_pyramid(1.0), 1024, 1024, cuda.SAMPLER_SPECTRAL
    elif c == "5":
        sc = scenes.ibl_tessellated_mesh(ctx, 3840 / 2160, bvh_builder=S.BVH_DEVICE_LBVH if args.lbvh else S.BVH_REFERENCE)[0]
        w, h, sampler = 3840, 2160, cuda.SAMPLER_COLOUR
    else:
        raise SystemExit("config must be 1, 3, 4 or 5")
    t0 = time.perf_counter()
    hs = cuda.HostScene(sc)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.upload(hs)
    t_up = time.perf_counter() - t0
    if not args.no_warmup:
        render.New(ctx, w, h, args.spp, 50, sampler_type=sampler, seed=3, sample_count=1).Render()
    out = {"config": c, "spp": args.spp, "host_scene_build_s": t_build, "upload_s": t_up}
    best = None
    for _ in range(args.repeat):
        r = render.New(ctx, w, h, args.spp, 50, sampler_type=sampler, seed=3, stats=cuda.RENDER_TIMING if args.stats else 0)
        r.canvas()
        l0 = ctx.launches
        t0 = time.perf_counter()
        r.Render()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        out.update({"seconds": dt, "best_seconds": best, "best_msamples_per_s": w * h * args.spp / best / 1e6, "msamples_per_s": w * h * args.spp / dt / 1e6, "mrays_per_s": r.num_rays / dt / 1e6, "launches": ctx.launches - l0})
        if args.stats:
            out["stats"] = ctx.render_stats()
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
