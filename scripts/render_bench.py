#!/usr/bin/env python
"""Msamples/s of the wavefront renderer on the BASELINE render configs (1, 3, 4) at a stated spp.

Throughput is measured through the public call (Context.render = izpi_host_render: setup, tiles,
finish incl. the D2H of the canvas).  `--cpu` times the oracle (C++ restatement of the Go CPU path,
LCG streams, all host threads) on a bounded window of the same image beside it.
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
    ap.add_argument("--configs", default="1,4,3")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (0 = per-config default below)")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--small", action="store_true", help="config 3 with a 60k-triangle mesh and 256^2 textures")
    args = ap.parse_args()
    from izpi_b200 import cuda, scenes
    from izpi_b200.build import build
    build()
    ctx = cuda.Context(0)
    out = []
    for c in [int(x) for x in args.configs.split(",")]:
        if c == 1:
            sc, w, h, spp, sampler, name = scenes.cornell_box(1.0), 400, 400, 64, cuda.SAMPLER_COLOUR, "cornell box 400x400 (config 1)"
        elif c == 4:
            sc, w, h, spp, sampler, name = scenes.spectral_pyramid(1.0), 1024, 1024, 16, cuda.SAMPLER_SPECTRAL, "spectral glass pyramid 1024x1024 (config 4)"
        elif c == 3:
            if args.small:
                sc = scenes.cornell_pbr_mesh(1.0, 300, 100, 256)
            else:
                sc = scenes.cornell_pbr_mesh(1.0)
            w, h, spp, sampler, name = 1024, 1024, 8, cuda.SAMPLER_COLOUR, "cornell + ~1M-triangle PBR mesh 1024x1024 (config 3)"
        else:
            continue
        if args.spp:
            spp = args.spp
        t0 = time.perf_counter()
        hs = cuda.HostScene(sc)
        t_build = time.perf_counter() - t0
        ctx.upload(hs)
        ctx.render(w, h, 1, sampler=sampler, seed=1)  # warm-up (allocations, module load)
        l0 = ctx.launches
        t0 = time.perf_counter()
        img, rays = ctx.render(w, h, spp, sampler=sampler, seed=2)
        dt = time.perf_counter() - t0
        rec = {"config": c, "scene": name, "width": w, "height": h, "spp": spp, "msamples_per_s": w * h * spp / dt / 1e6,
               "mrays_per_s": rays / dt / 1e6, "rays_per_sample": rays / (w * h * spp), "seconds": dt, "launches": ctx.launches - l0,
               "host_scene_build_s": t_build, "mean_rgb": [float(x) for x in img[1:, :, :3].mean(axis=(0, 1))]}
        if args.cpu:
            import oracle
            osn = oracle.OracleScene(sc)
            win = (0, h // 2, w - 1, h // 2 + 7)  # 8 rows through the middle of the image
            cspp = max(1, min(spp, 8))
            t0 = time.perf_counter()
            _, crays = osn.render(w, h, cspp, sampler=sampler, rng_mode=0, seed=3, window=win, epilogue=False)
            cdt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"msamples_per_s": w * 8 * cspp / cdt / 1e6, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"rows {win[1]}..{win[3]} at {cspp} spp, LCG streams"}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
