#!/usr/bin/env python
"""One process, all GPUs of the box: the device-group path of the C ABI (izpi_ctx_create with n_devices > 1) on the BASELINE
frames.  Scene flattened once and copied device-to-device, tiles claimed dynamically by one host thread per GPU, every GPU
writes its own tile runs into the caller's canvas.

  python scripts/group_render.py --gpus 8 --configs 5,4 [--quick]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--configs", default="5,4")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from izpi_b200 import cuda, render, scenes
    from izpi_b200.build import build
    build()
    one = cuda.Context(0)
    grp = cuda.Context(list(range(args.gpus)))
    out = {"gpus": args.gpus, "frames": []}
    for c in args.configs.split(","):
        if c == "5":
            sc, w, h, spp, sampler = scenes.ibl_tessellated_mesh(one, 3840 / 2160)[0], 3840, 2160, 1024, cuda.SAMPLER_COLOUR
        elif c == "4":
            sc, w, h, spp, sampler = scenes.spectral_pyramid(1.0), 1024, 1024, 1024, cuda.SAMPLER_SPECTRAL
        elif c == "3":
            sc, w, h, spp, sampler = scenes.cornell_pbr_mesh(1.0), 1024, 1024, 256, cuda.SAMPLER_COLOUR
        else:
            sc, w, h, spp, sampler = scenes.cornell_box(1.0), 400, 400, 64, cuda.SAMPLER_COLOUR
        if args.quick:
            spp = max(1, spp // 16)
        hs = cuda.HostScene(sc)
        t0 = time.perf_counter()
        grp.upload(hs)  # H2D to the first device + cudaMemcpyPeerAsync to the others
        t_up = time.perf_counter() - t0
        render.New(grp, w, h, spp, 50, sampler_type=sampler, seed=3, sample_count=1).Render()
        r = render.New(grp, w, h, spp, 50, sampler_type=sampler, seed=3)
        r.canvas()
        t0 = time.perf_counter()
        img = r.Render()
        dt = time.perf_counter() - t0
        rec = {"config": c, "spp": spp, "seconds": dt, "msamples_per_s": w * h * spp / dt / 1e6, "upload_and_replicate_s": t_up,
               "frame_ms": r.timings, "rays": r.num_rays, "canvas_sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()}
        out["frames"].append(rec)
        print(json.dumps(rec), flush=True)
        del hs, sc
    # closest hit through the group: one 16.7M-ray host batch cut into contiguous slices
    sc, lo, hi = scenes.closest_hit_scene()
    hs = cuda.HostScene(sc)
    one.upload(hs)
    grp.upload(hs)
    org, d = scenes.random_rays(1 << 24, lo, hi)
    import torch
    h_org, h_dir = torch.from_numpy(org).pin_memory(), torch.from_numpy(d).pin_memory()
    ids = torch.empty(1 << 24, dtype=torch.int32).pin_memory()
    t = torch.empty(1 << 24, dtype=torch.float64).pin_memory()
    for _ in range(2):
        grp.trace_closest(h_org.numpy(), h_dir.numpy(), out_ids=ids.numpy(), out_t=t.numpy())
    t0 = time.perf_counter()
    for _ in range(3):
        grp.trace_closest(h_org.numpy(), h_dir.numpy(), out_ids=ids.numpy(), out_t=t.numpy())
    dt = (time.perf_counter() - t0) / 3
    i1, t1 = one.trace_closest(org[: 1 << 20], d[: 1 << 20])
    out["trace_e2e"] = {"mrays_per_s": (1 << 24) / dt / 1e6, "rays": 1 << 24, "same_as_one_device": bool(np.array_equal(i1, ids.numpy()[: 1 << 20]) and t1.tobytes() == t.numpy()[: 1 << 20].tobytes())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
