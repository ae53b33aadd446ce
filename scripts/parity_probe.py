#!/usr/bin/env python
"""Where do the device and the oracle differ on a same-path render, and by how much?

For every parity scene: one path per pixel (spp = 1, the oracle replaying the device's counter RNG), the histogram of
per-pixel relative differences, the oracle's own sensitivity to +-2 ulp of libm jitter, and for the worst pixels the depth
cap at which the two sides first disagree (depth cap d returns beta * (0,0,1) for a path that is still alive after d
bounces, colour.go:34-36, so the blue channel is the path's throughput after d bounces).
Writes gpurun_out/parity_probe.json."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(s > 0, d / s, 0.0)
    return np.nan_to_num(r, nan=np.inf).max(axis=-1)


def main():
    import oracle
    from izpi_b200 import cuda, scenes
    from izpi_b200 import scene as S
    from izpi_b200.build import build
    build()
    oracle.build()
    ctx = cuda.Context(0)
    cases = []
    sc = scenes.cornell_box(1.0)
    cases.append(("cornell", sc, 96, 96, cuda.SAMPLER_COLOUR))
    cases.append(("spectral_pyramid", scenes.spectral_pyramid(1.0), 96, 96, cuda.SAMPLER_SPECTRAL))
    cases.append(("pbr_mesh", scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64), 96, 96, cuda.SAMPLER_COLOUR))
    cases.append(("ibl", scenes.ibl_displaced_mesh(16 / 9, 120, 60, (256, 128)), 128, 72, cuda.SAMPLER_COLOUR))
    out = {}
    for name, spec, w, h, sampler in cases:
        ctx.upload(cuda.HostScene(spec))
        osn = oracle.OracleScene(spec)
        rec = {}
        for spp in (1, 8):
            img, rays = ctx.render(w, h, spp, sampler=sampler, seed=5)
            ref, rrays = osn.render(w, h, spp, sampler=sampler, rng_mode=1, seed=5, epilogue=True)
            r = rel(img[1:, :, :3], ref[1:, :, :3])
            jit = np.zeros_like(r, dtype=bool)
            for j in (1, 2, 3):
                c, _ = osn.render(w, h, spp, sampler=sampler, rng_mode=1, seed=5, epilogue=True, libm_jitter=j)
                jit |= rel(c[1:, :, :3], ref[1:, :, :3]) > 1e-7
            rec[f"spp{spp}"] = {"pixels": int(r.size), "rays_dev": int(rays), "rays_ref": int(rrays),
                                "frac_gt": {str(t): float((r > t).mean()) for t in (1e-13, 1e-11, 1e-9, 1e-7, 1e-5, 1e-3, 1e-1)},
                                "jitter_unstable_frac": float(jit.mean()), "diverging_and_unstable": int(((r > 1e-7) & jit).sum()),
                                "diverging_but_stable": int(((r > 1e-7) & ~jit).sum())}
            if spp == 1:
                ys, xs = np.nonzero(r > 1e-7)
                worst = []
                for y, x in list(zip(ys, xs))[:6]:
                    row = int(y) + 1  # canvas row (row 0 was cut), image y = h - row
                    iy = h - row
                    trail = []
                    if sampler == cuda.SAMPLER_COLOUR:
                        # one-pixel window does not exist on the device path: render the tile row's pixels and pick ours
                        for d in (1, 2, 3, 4, 5, 6, 8, 10, 14, 20, 30, 50):
                            a, _ = ctx.render(w, h, 1, max_depth=d, sampler=sampler, seed=5)
                            b, _ = osn.render(w, h, 1, max_depth=d, sampler=sampler, rng_mode=1, seed=5, window=(int(x), iy, int(x), iy))
                            trail.append({"depth": d, "dev": a[row, x, :3].tolist(), "ref": b[row, x, :3].tolist()})
                            if rel(a[row, x, :3][None], b[row, x, :3][None])[0] > 1e-7:
                                break
                    worst.append({"x": int(x), "y": iy, "dev": img[row, x, :3].tolist(), "ref": ref[row, x, :3].tolist(), "rel": float(r[y, x]),
                                  "trail": trail})
                rec["worst"] = worst
        out[name] = rec
        print(name, json.dumps({k: v for k, v in rec.items() if k != "worst"}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
