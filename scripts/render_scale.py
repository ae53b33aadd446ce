#!/usr/bin/env python
"""Config 5 (4K image-based-lit, ~10M-triangle mesh) tile-sharded over the GPUs of one box (torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      scripts/render_scale.py --spp 16

Prints one JSON line on rank 0: Msamples/s end to end through render.New(...).Render() (setup, tiles,
NCCL reduce of the 265 MB fp64 canvas, epilogue, D2H), max over ranks, plus the scene build times.
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
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--around", type=int, default=3162)
    ap.add_argument("--tube", type=int, default=1581)
    ap.add_argument("--checksum", action="store_true")
    ap.add_argument("--tessellated", action="store_true", help="build the mesh with the displacement tessellator (izpi_displace) "
                    "instead of the analytically displaced torus")
    ap.add_argument("--lbvh", action="store_true", help="BVH4 built on the device (izpi_bvh4_build) instead of the restated NewBVH4")
    args = ap.parse_args()
    import numpy as np
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
    t0 = time.perf_counter()
    if args.tessellated:
        from izpi_b200 import scene as S
        sc, n_tris = scenes.ibl_tessellated_mesh(ctx, args.width / args.height, bvh_builder=S.BVH_DEVICE_LBVH if args.lbvh else S.BVH_REFERENCE)
    else:
        sc, n_tris = scenes.ibl_displaced_mesh(args.width / args.height, args.around, args.tube), 2 * args.around * args.tube
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    hs = cuda.HostScene(sc, threads=max(1, (os.cpu_count() or 8) // world))
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.upload(hs)
    t_up = time.perf_counter() - t0
    render.New(ctx, args.width, args.height, 1, 50, seed=1).Render()  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = render.New(ctx, args.width, args.height, args.spp, 50, seed=7)
    img = r.Render()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    if rank == 0:
        n = args.width * args.height * args.spp
        line = {"scene": "config 5: IBL + ~10M-triangle displaced mesh", "triangles": int(n_tris), "mesh": "displacement-tessellated (izpi_displace)" if args.tessellated else "analytic torus",
                "bvh": "device LBVH" if (args.lbvh and args.tessellated) else "reference (NewBVH4 restated)", "width": args.width,
                "height": args.height, "spp": args.spp, "n_gpus": world, "msamples_per_s": n / dt / 1e6, "seconds": dt,
                "mrays_per_s": r.num_rays / dt / 1e6, "rays_per_sample": r.num_rays / n, "scene_gen_s": t_gen,
                "host_bvh_build_s": t_build, "upload_s": t_up, "mean_rgb": [float(x) for x in img[1:, :, :3].mean(axis=(0, 1))]}
        if args.checksum:
            import hashlib
            line["canvas_sha256"] = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
