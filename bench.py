#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200.

Workload (config.workload): BASELINE configs[1], the closest-hit microbench -- synthetic
1,000,000-triangle displaced torus in the reference's BVH4 (LCG seed 12345), 16,777,216 incoherent
random rays per step, exact fp64 mode.  One step = one pass of the hot path over one ray batch.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement on the host cores)

Prints ONE JSON line (rank 0).  `value` = whole-job Mrays/s with rays resident in HBM;
`e2e` = the same through izpi_trace_closest with pinned HOST buffers (copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 1 << 24
WORKLOAD = "closest-hit: 1M-triangle displaced torus (BVH4, LCG seed 12345), 16,777,216 incoherent rays/step, fp64 exact"


def ncu_traffic(n_rays):
    """DRAM bytes per launch of the headline kernel from the committed `ncu --set full` capture (profiles/), scaled to the
    launch size when it differs from the captured one.  None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r01_trace_traffic.json")
    try:
        t = json.load(open(p))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (n_rays / t["rays_per_launch"]), t["kernel"]
    except Exception:
        return None, "trace_g2_kernel<false,false,48>"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _run_nvml(self):
        """Same fields through NVML (nvidia_ml_py), sampled every 20 ms: a timed region of a few steps lasts well under a
        second, which an nvidia-smi process per sample cannot resolve."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons(h)
        except Exception:
            return False
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.rows.append([str(sm), str(mx)] + [("Active" if r & b else "Not Active") for b in bits.values()])
            except Exception:
                pass
            self.stop_flag.wait(0.02)
        return True

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def build_inputs(rank):
    from izpi_b200 import scenes
    sc, lo, hi = scenes.closest_hit_scene()
    return sc, lo, hi


def cpu_baseline(sc, lo, hi, n_sample, threads):
    """The C++ restatement of the Go CPU path (oracle/), timed on the host cores on a bounded sample."""
    import oracle
    from izpi_b200 import scenes
    osn = oracle.OracleScene(sc)
    org, d = scenes.random_rays(n_sample, lo, hi)
    osn.trace(org[:4096], d[:4096], threads=threads)  # warm
    t0 = time.perf_counter()
    ids, t, st = osn.trace(org, d, threads=threads, stats=True)
    dt = time.perf_counter() - t0
    return n_sample / dt / 1e6, st, (ids, t, org, d)


def bench_render(ctx, cuda, scenes, world, rank, barrier, with_4k=True):
    """Second half of BASELINE's metric: Msamples/s of the tile renderer, end to end through
    render.New(...).Render() (setup, tiles, NCCL reduce of the fp64 canvas for N>1, epilogue, D2H).
    Strong scaling: the image is fixed, its tiles are dealt to the ranks."""
    import time as _t
    from izpi_b200 import render
    out = []
    cases = [("config 1: cornell box 400x400, 64 spp, colour", lambda: scenes.cornell_box(1.0), 400, 400, 64, cuda.SAMPLER_COLOUR),
             ("config 4 scene: spectral glass pyramid 1024x1024 at 16 spp (BASELINE: 1024 spp)", lambda: scenes.spectral_pyramid(1.0), 1024, 1024,
              16, cuda.SAMPLER_SPECTRAL)]
    if with_4k:  # the multi-GPU target of BASELINE: a 4K render, tile-sharded
        cases.append(("config 5 scene: 4K IBL + displacement-tessellated 11.5M-triangle mesh 3840x2160 at 16 spp (BASELINE: 1024 spp)",
                      lambda: scenes.ibl_tessellated_mesh(ctx, 3840 / 2160)[0], 3840, 2160, 16, cuda.SAMPLER_COLOUR))
        # same frame over the optional device-built BVH4 (izpi_bvh4_build): same closest hits, better tree; reported separately
        from izpi_b200 import scene as _S
        cases.append(("config 5 scene, BVH4 built on the device (Morton LBVH instead of the reference's random-axis median split), 16 spp",
                      lambda: scenes.ibl_tessellated_mesh(ctx, 3840 / 2160, bvh_builder=_S.BVH_DEVICE_LBVH)[0], 3840, 2160, 16, cuda.SAMPLER_COLOUR))
    for name, make, w, h, spp, sampler in cases:
        spec = make()
        ctx.upload(cuda.HostScene(spec, threads=max(1, (os.cpu_count() or 8) // world)))
        del spec
        r = render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=3)
        render.New(ctx, w, h, 1, 50, sampler_type=sampler, seed=3).Render()  # warm-up
        barrier()
        t0 = _t.perf_counter()
        r.Render()
        barrier()
        dt = _t.perf_counter() - t0
        out.append({"scene": name, "msamples_per_s": w * h * spp / dt / 1e6, "seconds": dt, "scaling": "strong",
                    "mrays_per_s": (r.num_rays / dt / 1e6) if rank == 0 else None})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from izpi_b200 import scenes
    threads = os.cpu_count() or 1
    sc, lo, hi = build_inputs(0)
    osn = oracle.OracleScene(sc)
    n_sample = 1 << 20  # bounded sample of the 16.7M-ray step (~2 s of host work per step)
    org, d = scenes.random_rays(n_sample, lo, hi)
    for _ in range(args.warmup):
        osn.trace(org[: n_sample // 8], d[: n_sample // 8], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        osn.trace(org, d, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_sample / dt / 1e6
    line = {"impl": "reference", "metric": "Mrays/s closest-hit (1M tris, incoherent)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "C++ restatement of Izpi's Go CPU path (no Go toolchain in this image), not the Go binary"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": f"{n_sample} of the step's {N_RAYS} rays per step, all host threads"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on some boxes), so
    file descriptor 1 is pointed at stderr for the whole run and the result line goes to a private copy of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="izpi_b200")
    ap.add_argument("--rays", type=int, default=N_RAYS, help="rays per step per GPU (default = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-4k", action="store_true", help="skip the 4K / 10M-triangle render (saves ~20 s of host-side scene building)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from izpi_b200 import cuda, scenes
    from izpi_b200.build import build
    build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)
    n = args.rays

    sc, lo, hi = build_inputs(rank)
    hs = cuda.HostScene(sc)
    ctx = cuda.Context(local_rank)
    ctx.upload(hs)
    # weak scaling: every rank traces its own 16.7M-ray batch (counter RNG offset by rank), no collective
    org, d = scenes.random_rays(n, lo, hi, start=rank * n)
    h_org = torch.from_numpy(org).pin_memory()
    h_dir = torch.from_numpy(d).pin_memory()
    h_ids = torch.empty(n, dtype=torch.int32).pin_memory()
    h_t = torch.empty(n, dtype=torch.float64).pin_memory()
    d_org, d_dir = h_org.cuda(), h_dir.cuda()
    d_ids = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float64, device="cuda")
    tstream = torch.cuda.Stream()  # a real (non-default) stream: the kernels are launched on it and the events see it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step_device():
        ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids.data_ptr(), d_t.data_ptr(), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # algorithmic bytes per ray from the counting kernel (same traversal as the oracle; tests assert equality)
    _, _, st = ctx.trace_closest(org[: 1 << 20], d[: 1 << 20], stats=True)
    nodes_per_ray, prims_per_ray = st["nodes"] / st["rays"], st["prims"] / st["rays"]
    bytes_per_ray = 128.0 * nodes_per_ray + 72.0 * prims_per_ray + 48.0 + 12.0

    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for a, b in evs:
        a.record()
        step_device()
        b.record()
    e1.record()
    barrier()
    launches = ctx.launches - launches0
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # end to end through the C ABI with pinned host buffers: H2D of the rays + D2H of ids/t inside the timed region
    np_org, np_dir, np_ids, np_t = h_org.numpy(), h_dir.numpy(), h_ids.numpy(), h_t.numpy()
    for _ in range(2):
        ctx.trace_closest(np_org, np_dir, out_ids=np_ids, out_t=np_t)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.trace_closest(np_org, np_dir, out_ids=np_ids, out_t=np_t)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag.set()
    sampler.join()

    # results of the two paths agree
    assert np.array_equal(np_ids, d_ids.cpu().numpy()) and np.array_equal(np_t, d_t.cpu().numpy())

    # optional fp32-primitive mode, reported separately (BASELINE north_star): same traversal, fp32 triangle test
    d_ids32 = torch.empty_like(d_ids)
    d_t32 = torch.empty_like(d_t)
    for _ in range(2):
        ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids32.data_ptr(), d_t32.data_ptr(), mode=cuda.TRACE_FP32, stream=stream)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids32.data_ptr(), d_t32.data_ptr(), mode=cuda.TRACE_FP32, stream=stream)
    f1.record()
    torch.cuda.synchronize()
    fp32_ms = f0.elapsed_time(f1) / args.steps
    step_device()
    torch.cuda.synchronize()
    fp32_info = {"value": n / (fp32_ms * 1e-3) / 1e6, "unit": "Mrays/s per GPU", "ids_equal_to_exact": float((d_ids32 == d_ids).double().mean().item()),
                 "note": "fp32 Moeller-Trumbore, same BVH4 traversal; not part of the parity claim"}

    render_info = bench_render(ctx, cuda, scenes, world, rank, barrier, with_4k=not args.no_4k)
    ctx.upload(hs)  # back to the closest-hit scene for the CPU-baseline parity check below

    # FLOP side of the traversal roofline (north_star: "the slower of bytes per ray at HBM bandwidth and intersection FLOPs at
    # fp64/fp32 peak"): measured dependent-FMA peaks of this GPU, SURVEY.md §8(d)'s operation counts per visit / per test
    fp32_peak, fp64_peak = ctx.fma_peak(False), ctx.fma_peak(True)
    fp32_flop_per_ray, fp64_flop_per_ray = 100.0 * nodes_per_ray, 50.0 * prims_per_ray

    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, kernel_ms = tt.tolist()
    if rank == 0:
        hbm, peak_src = peaks()
        value = world * n * args.steps / (total_ms * 1e-3) / 1e6
        achieved = bytes_per_ray * n / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s closest-hit (1M tris, incoherent)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n, "l2": "inputs larger than L2 (805 MB of rays per step)",
                       "mode": "exact (reference traversal order, fp32 slab test, fp64 primitives)",
                       "nodes_per_ray": nodes_per_ray, "prim_tests_per_ray": prims_per_ray,
                       "algorithmic_bytes_per_ray": bytes_per_ray},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": ncu_traffic(n)[0], "kernel": ncu_traffic(n)[1], "kernel_ms": kernel_ms, "peak_source": peak_src},
            "roofline_flops": {"fp32_tflops_peak": fp32_peak, "fp64_tflops_peak": fp64_peak, "fp32_flop_per_ray": fp32_flop_per_ray,
                               "fp64_flop_per_ray": fp64_flop_per_ray,
                               "mrays_per_s_at_flop_peak": 1e-6 / (fp32_flop_per_ray / (fp32_peak * 1e12) + fp64_flop_per_ray / (fp64_peak * 1e12)),
                               "mrays_per_s_at_hbm_peak": hbm * 1e3 / bytes_per_ray,
                               "note": "the byte side is the slower (binding) one; peaks measured by izpi_debug_fma_peak on this GPU"},
            "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n * 48,
                    "d2h_bytes_per_step": n * 12},
            "gpu_launches": int(launches),
            "render": render_info,
            "fp32_mode": fp32_info,
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (rank 0, bounded sample)
            threads = os.cpu_count() or 1
            n_cpu = 1 << 22
            v, ost, (oi, ot, oo, od) = cpu_baseline(sc, lo, hi, n_cpu, threads)
            gi, gt = ctx.trace_closest(oo, od)
            line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                    "sample": f"first {n_cpu} rays of the step, {threads} host threads; C++ restatement of the Go CPU path",
                                    "parity_on_sample": bool(np.array_equal(gi, oi) and gt.tobytes() == ot.tobytes())}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
