"""Quick kernel-iteration loop: config-2 closest hit, rays resident in HBM, best-of-N CUDA-event time of one launch, plus a
bit-exactness check of a sample against the oracle.   python scripts/trace_speed.py [--rays N] [--reps R]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from izpi_b200 import cuda, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tag", default="")
    ap.add_argument("--lbvh", action="store_true", help="BVH4 built on the device (izpi_bvh4_build) instead of the reference's tree")
    a = ap.parse_args()
    sc, lo, hi = scenes.closest_hit_scene()
    if a.lbvh:
        from izpi_b200 import scene as S
        sc.bvh_builder = S.BVH_DEVICE_LBVH
    ctx = cuda.Context(0)
    ctx.upload(cuda.HostScene(sc))
    n = a.rays
    org, d = scenes.random_rays(n, lo, hi)
    d_org, d_dir = torch.from_numpy(org).cuda(), torch.from_numpy(d).cuda()
    d_ids = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()
    times = []
    with torch.cuda.stream(stream):
        for rep in range(a.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids.data_ptr(), d_t.data_ptr(), stream=stream.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                times.append(e0.elapsed_time(e1))
    m = 1 << 16
    oi, ot = oracle.OracleScene(sc).trace(org[:m], d[:m], threads=os.cpu_count())
    ok = bool(np.array_equal(d_ids[:m].cpu().numpy(), oi) and d_t[:m].cpu().numpy().tobytes() == ot.tobytes())
    print(json.dumps({"tag": a.tag, "ms_best": min(times), "ms_mean": float(np.mean(times)), "mrays_per_s": n / min(times) / 1e3, "parity_64k": ok}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
