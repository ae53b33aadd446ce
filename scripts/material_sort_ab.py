#!/usr/bin/env python
"""A/B of the material-ID bins (north_star 3: paths sorted by material ID between bounces) on a multi-material variant of
config 3: the ~1M-triangle PBR mesh split into bands with N PBR materials, each with its own four 2048^2 fp64 textures.
IZPI_MATERIAL_BINS=0 files hits under their material CLASS only (round-1 behaviour); the default gives every
image-textured material its own bin, so a warp of shade<PBR> samples one texture set.

  python scripts/material_sort_ab.py --materials 8 --spp 32 [--tex 2048]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--materials", type=int, default=8)
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--tex", type=int, default=2048)
    ap.add_argument("--only", default="", help="'0' or '1': run one arm only (for ncu)")
    args = ap.parse_args()
    from izpi_b200 import cuda, render, scenes
    from izpi_b200.build import build
    build()
    ctx = cuda.Context(0)
    sc = scenes.cornell_pbr_mesh(1.0, tex_size=args.tex, n_pbr_materials=args.materials)
    hs = cuda.HostScene(sc)
    out = {"materials": args.materials, "spp": args.spp, "tex": args.tex}
    for mode in ("0", "1"):
        if args.only and args.only != mode:
            continue
        os.environ["IZPI_MATERIAL_BINS"] = mode
        ctx.upload(hs)
        render.New(ctx, 1024, 1024, 1, 50, seed=3).Render()
        best = None
        for _ in range(2):
            r = render.New(ctx, 1024, 1024, args.spp, 50, seed=3)
            t0 = time.perf_counter()
            img = r.Render()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        rt = render.New(ctx, 1024, 1024, args.spp, 50, seed=3, stats=cuda.RENDER_TIMING)
        rt.Render()
        st = ctx.render_stats()
        out["class_bins" if mode == "0" else "material_bins"] = {
            "bins": st["material_bins"], "seconds": best, "msamples_per_s": 1024 * 1024 * args.spp / best / 1e6,
            "shade_ms": st["shade_ms"], "extend_ms": st["extend_ms"], "sha256": hashlib.sha256(img.tobytes()).hexdigest()[:16]}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
