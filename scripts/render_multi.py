#!/usr/bin/env python
"""Multi-GPU render check, launched with torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/render_multi.py

Every rank renders its round-robin share of the tiles, the fp64 canvases are SUM-reduced over NCCL, and
rank 0 checks the result is BIT-IDENTICAL to a single-GPU render of the same scene and seed (disjoint
pixels: the sum is exact), then prints the timing.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from izpi_b200 import cuda, render, scenes
    from izpi_b200.build import build
    build()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cuda.Context(local)
    results = []
    for name, spec, w, h, spp, sampler in (
            ("cornell", scenes.cornell_box(1.0), 400, 400, 64, cuda.SAMPLER_COLOUR),
            ("spectral_pyramid", scenes.spectral_pyramid(1.0), 512, 512, 32, cuda.SAMPLER_SPECTRAL)):
        ctx.upload(cuda.HostScene(spec))
        render.New(ctx, w, h, 1, 50, sampler_type=sampler, seed=5).Render()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=5)
        img = r.Render()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            single, rays = ctx.render(w, h, spp, sampler=sampler, seed=5)  # the whole image on this GPU alone
            same = bool(np.array_equal(img, single, equal_nan=True))
            results.append({"scene": name, "world": world, "bit_identical_to_single_gpu": same, "rays_match": r.num_rays == rays,
                            "msamples_per_s": w * h * spp / dt / 1e6})
            assert same, f"{name}: multi-GPU canvas differs from the single-GPU canvas"
            assert r.num_rays == rays
    if rank == 0:
        print(json.dumps(results))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
