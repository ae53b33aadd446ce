#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200: Mrays/s closest hit and Msamples/s render, beside the CPU path.

Headline workload (config.workload): BASELINE configs[1], the closest-hit microbench -- synthetic
1,000,000-triangle displaced torus in the reference's BVH4 (LCG seed 12345), 16,777,216 incoherent
random rays per step, exact fp64 mode.  One step = one pass of the hot path over one ray batch.
The same line carries the render half of the metric in `render`: configs 1, 3, 4 and 5 at BASELINE's resolutions and
sample counts (`--quick` divides the sample counts by 16 for development sweeps), each with its own roofline entry for the
extend (closest-hit) kernel -- algorithmic bytes from the counting kernels, device time from CUDA events -- and, at N = 1,
the C++ restatement of the Go CPU path timed on a window of the same frame.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement on the host cores)

Prints ONE JSON line (rank 0).  `value` = whole-job Mrays/s with rays resident in HBM (weak scaling, no collective);
`e2e` = the same through izpi_trace_closest with pinned HOST buffers (copies inside the timed region);
`render[*]` = strong scaling: the frame is fixed, ranks claim its tiles from one shared cursor, one NCCL reduce merges.
"""
from __future__ import annotations

import argparse
import hashlib
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
TRACE_SOURCES = ["izpi_b200/csrc/device/trace.cu", "izpi_b200/csrc/device/intersect.cuh", "izpi_b200/csrc/device/intersect_g2.cuh",
                 "izpi_b200/csrc/device/intersect_g4.cuh", "izpi_b200/csrc/device/dscene.cuh", "izpi_b200/csrc/device/context.cu"]
RENDER_SOURCES = TRACE_SOURCES + ["izpi_b200/csrc/device/render.cu", "izpi_b200/csrc/device/shade.cuh", "izpi_b200/csrc/device/shade_textures.cuh"]


def kernel_source_hash(key="render"):
    """sha256 over the sources of the kernels an ncu capture describes: the closest-hit kernels (keys `trace_*`) or the whole
    renderer (keys `render_*`).  A capture is valid for ONE source state."""
    h = hashlib.sha256()
    for f in (TRACE_SOURCES if key.startswith("trace") else RENDER_SOURCES):
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def ncu_traffic(key, units):
    """DRAM bytes per launch of kernel `key` from the committed `ncu --set full` capture (profiles/r02_traffic.json,
    written by scripts/ncu_traffic.py), scaled to the launch size.  The capture records the hash of the kernel sources it
    was taken from; when the sources have changed since, the number describes another kernel and is NOT printed."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        t = json.load(open(p))[key]
    except Exception:
        return None, None, "no ncu capture committed for this kernel"
    if t.get("source_sha256") != kernel_source_hash(key):
        return None, t.get("kernel"), "stale: kernel sources changed after the ncu capture (profiles/r02_traffic.json)"
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (units / t["units_per_launch"]), t.get("kernel"), t.get("source")


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


# ---- the render half of the metric -----------------------------------------------------------------------------------
def render_cases(cuda, scenes, ctx, quick):
    """(config number, name, scene builder, width, height, BASELINE spp, sampler).  Builders run on rank 0 only."""
    from izpi_b200 import scene as S
    div = 16 if quick else 1
    return [
        (1, "config 1: cornell box (HitableSlice, 8 hitables) 400x400, colour", lambda: scenes.cornell_box(1.0), 400, 400, max(1, 64 // div), cuda.SAMPLER_COLOUR),
        (3, "config 3: cornell box + ~1M-triangle PBR mesh, 4 x 2048^2 fp64 textures, 1024x1024, colour", lambda: scenes.cornell_pbr_mesh(1.0), 1024, 1024,
         max(1, 256 // div), cuda.SAMPLER_COLOUR),
        (4, "config 4: spectral glass pyramid (dispersion, Beer-Lambert) 1024x1024, spectral", lambda: scenes.spectral_pyramid(1.0), 1024, 1024,
         max(1, 1024 // div), cuda.SAMPLER_SPECTRAL),
        (5, "config 5: 4K IBL + displacement-tessellated 11.5M-triangle mesh 3840x2160, colour, reference-shaped BVH4 (NewBVH4 restated)",
         lambda: scenes.ibl_tessellated_mesh(ctx, 3840 / 2160)[0], 3840, 2160, max(1, 1024 // div), cuda.SAMPLER_COLOUR),
        ("5-lbvh", "config 5 scene over the BVH4 built on the device (Morton LBVH instead of the reference's random-axis median split; same closest hits)",
         lambda: scenes.ibl_tessellated_mesh(ctx, 3840 / 2160, bvh_builder=S.BVH_DEVICE_LBVH)[0], 3840, 2160, max(1, 1024 // div), cuda.SAMPLER_COLOUR),
    ]


def bench_render(ctx, cuda, scenes, world, rank, barrier, args):
    """Msamples/s of the tile renderer, end to end through render.New(...).Render(): setup, tiles (claimed dynamically from
    one shared cursor for N > 1), NCCL reduce of the fp64 canvas for N > 1, epilogue, D2H into pinned host memory.
    Strong scaling: the frame is fixed.  The scene is built ONCE per box (rank 0) and replicated over NCCL."""
    import torch
    import torch.distributed as dist
    from izpi_b200 import render
    hbm, peak_src = peaks()
    out = []
    only = set(args.render_configs.split(",")) if args.render_configs else None
    for cfg_id, name, make, w, h, spp, sampler in render_cases(cuda, scenes, ctx, args.quick):
        if only is not None and str(cfg_id) not in only:
            continue
        if args.no_4k and str(cfg_id).startswith("5"):
            continue
        spec = None
        t_build = t_upload = t_make = 0.0
        if rank == 0:
            t0 = time.perf_counter()
            spec = make()
            t_make = time.perf_counter() - t0
            hs = cuda.HostScene(spec, threads=os.cpu_count() or 8)
            t_build = time.perf_counter() - t0
            t0 = time.perf_counter()
            ctx.upload(hs)
            t_upload = time.perf_counter() - t0
            n_prims = hs.n_prims
            del hs
        t0 = time.perf_counter()
        render.replicate_scene(ctx, 0)
        t_repl = time.perf_counter() - t0
        render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=3, sample_count=1).Render()  # warm-up: ONE sample of the same frame (its buffers, module load, cursor)
        r = render.New(ctx, w, h, spp, 50, sampler_type=sampler, seed=3)
        if rank == 0:
            r.canvas()
        barrier()
        l0 = ctx.launches
        t0 = time.perf_counter()
        img = r.Render()
        barrier()
        dt = time.perf_counter() - t0
        tm = [r.timings.get(k, 0.0) for k in ("setup_ms", "tiles_ms", "merge_ms", "finish_ms")]
        if world > 1:
            tt = torch.tensor([dt] + tm, device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt, tm = float(tt[0]), [float(x) for x in tt[1:]]
        entry = None
        if rank == 0:
            entry = {"config": cfg_id, "scene": name, "width": w, "height": h, "spp": spp, "msamples_per_s": w * h * spp / dt / 1e6,
                     "mrays_per_s": r.num_rays / dt / 1e6, "rays_per_sample": r.num_rays / (w * h * spp), "seconds": dt, "scaling": "strong",
                     "n_gpus": world, "primitives": n_prims, "gpu_launches_rank0": int(ctx.launches - l0),
                     "frame_ms_max_over_ranks": {"setup": tm[0], "tiles": tm[1], "merge_nccl_reduce": tm[2], "epilogue_and_d2h": tm[3]},
                     "scene_s": {"generate_rank0": t_make, "constructors_and_bvh_rank0": t_build - t_make, "upload_rank0": t_upload, "replicate_nccl": t_repl},
                     "canvas_sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
                     "mean_rgb": [float(x) for x in img[1:, :, :3].mean(axis=(0, 1))]}
        if world == 1 and not args.no_render_stats:
            # Roofline of the dominant kernel of the frame, the extend (closest-hit) stage: algorithmic bytes from the counting
            # kernels (SURVEY.md 8d formula, same as config 2), device time from CUDA events around every extend launch of a frame
            # rendered with ONE batch in flight (so that a launch's time is its own), both at a reduced sample count (the per-ray
            # statistics do not depend on it).
            # enough samples for full batches (64 M paths on the large trees): small frames are launch-bound, not kernel-bound
            s_spp = max(1, min(spp, max(4, spp // 16, -(-(1 << 26) // (w * h)))))
            rs = render.New(ctx, w, h, s_spp, 50, sampler_type=sampler, seed=3, stats=cuda.RENDER_STATS)
            rs.Render()
            st = ctx.render_stats()
            rt = render.New(ctx, w, h, s_spp, 50, sampler_type=sampler, seed=3, stats=cuda.RENDER_TIMING)
            t0 = time.perf_counter()
            rt.Render()
            t_serial = time.perf_counter() - t0
            tt_ = ctx.render_stats()
            rays = max(1, st["rays"])
            bytes_per_ray = (128.0 * st["nodes_visited"] + 72.0 * st["prim_tests"]) / rays + 60.0
            achieved = bytes_per_ray * rays / (tt_["extend_ms"] * 1e-3) / 1e9 if tt_["extend_ms"] > 0 else None
            stage_total = tt_["extend_ms"] + tt_["shade_ms"] + tt_["other_ms"]
            kernel = "extend_kernel (thread per ray)"
            if st["nodes_visited"] and n_prims >= 64:
                kernel = "extend_g2_kernel (two lanes per ray)"
            key = f"render_cfg{cfg_id}_extend"
            traffic, tk, tsrc = ncu_traffic(key, rays)
            entry["roofline"] = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                                 "frac": (achieved / hbm) if achieved else None, "traffic": traffic, "traffic_per_ray": (traffic / rays) if traffic else None,
                                 "traffic_note": "DRAM bytes of all extend launches of the measured frame (ncu bytes per ray x its rays)",
                                 "traffic_source": tsrc, "peak_source": peak_src,
                                 "algorithmic_bytes_per_ray": bytes_per_ray, "nodes_per_ray": st["nodes_visited"] / rays,
                                 "prim_tests_per_ray": st["prim_tests"] / rays, "measured_at_spp": s_spp,
                                 "extend_ms": tt_["extend_ms"], "shade_ms": tt_["shade_ms"], "raygen_resolve_ms": tt_["other_ms"],
                                 "extend_launches": int(tt_["extend_launches"]),
                                 "extend_share_of_gpu_time": tt_["extend_ms"] / stage_total if stage_total > 0 else None,
                                 "extend_mrays_per_s": rays / (tt_["extend_ms"] * 1e-3) / 1e6 if tt_["extend_ms"] > 0 else None,
                                 "serialised_frame_s": t_serial,
                                 "note": "achieved = algorithmic bytes of all extend launches / their summed CUDA-event time; the shade kernels "
                                         "are compute/latency-bound (fp64 libm), see DESIGN.md"}
        if world == 1 and not args.no_cpu_baseline and rank == 0:
            try:
                entry["cpu_baseline"] = cpu_render_baseline(spec, w, h, spp, sampler, entry["msamples_per_s"])
            except Exception as e:  # the oracle scene of config 5 needs ~6 GB of host memory
                entry["cpu_baseline"] = {"unavailable": str(e)}
        del spec
        if rank == 0:
            out.append(entry)
    return out


def cpu_render_baseline(spec, w, h, spp, sampler, gpu_msamples):
    """C++ restatement of the Go CPU path (recursive integrators, LCG streams, one thread per host core, tiles handed out in
    chunks like renderer.go:126-138) on a bounded window of the same frame: rows through the middle of the image."""
    import oracle
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    osn = oracle.OracleScene(spec)
    t_scene = time.perf_counter() - t0
    rows = 8
    win = (0, h // 2, w - 1, h // 2 + rows - 1)
    cspp = 2
    _, _ = osn.render(w, h, 1, sampler=sampler, rng_mode=0, seed=3, window=(0, h // 2, w - 1, h // 2), epilogue=False)  # warm, one row
    t0 = time.perf_counter()
    _, crays = osn.render(w, h, cspp, sampler=sampler, rng_mode=0, seed=3, window=win, epilogue=False)
    dt = time.perf_counter() - t0
    # size the sample for about 3 s of host work: more samples per pixel first (up to the frame's), then more rows
    if dt < 2.0:
        want = 3.0 / max(dt, 1e-4) * (w * rows * cspp)
        cspp = int(min(spp, max(cspp, want / (w * rows))))
        rows = int(min(h - 2, max(rows, want / (w * cspp))))
        y0 = max(1, h // 2 - rows // 2)
        win = (0, y0, w - 1, y0 + rows - 1)
        t0 = time.perf_counter()
        _, crays = osn.render(w, h, cspp, sampler=sampler, rng_mode=0, seed=3, window=win, epilogue=False)
        dt = time.perf_counter() - t0
    v = w * rows * cspp / dt / 1e6
    del osn
    return {"value": v, "unit": "Msamples/s", "mrays_per_s": crays / dt / 1e6, "cores": threads, "kind": "port",
            "sample": f"rows {win[1]}..{win[3]} of the frame at {cspp} spp ({w * rows * cspp} samples, {dt:.1f} s), LCG streams; "
                      f"oracle scene built in {t_scene:.1f} s; C++ restatement of the Go CPU path, not the Go binary",
            "gpu_over_cpu": gpu_msamples / v}


def bench_lbvh_trace(ctx, cuda, scenes, n, d_org, d_dir, org, d, stream, steps, ref_ids, ref_t):
    """Config 2 over the BVH4 built on the device (izpi_bvh4_build): same triangles, same rays, same closest hits, a tree
    with ~2.8x fewer node visits.  Its own algorithmic bytes and roofline fraction (the headline's are defined on the
    reference tree's visit counts)."""
    import torch
    from izpi_b200 import scene as S
    sc, lo, hi = scenes.closest_hit_scene()
    sc.bvh_builder = S.BVH_DEVICE_LBVH
    t0 = time.perf_counter()
    hs = cuda.HostScene(sc)
    t_build = time.perf_counter() - t0
    ctx.upload(hs)
    _, _, st = ctx.trace_closest(org[: 1 << 20], d[: 1 << 20], stats=True)
    nodes_per_ray, prims_per_ray = st["nodes"] / st["rays"], st["prims"] / st["rays"]
    bytes_per_ray = 128.0 * nodes_per_ray + 72.0 * prims_per_ray + 60.0
    d_ids = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids.data_ptr(), d_t.data_ptr(), stream=stream)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        ctx.trace_closest_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_ids.data_ptr(), d_t.data_ptr(), stream=stream)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    hbm, _ = peaks()
    achieved = bytes_per_ray * n / (ms * 1e-3) / 1e9
    traffic, tk, tsrc = ncu_traffic("trace_lbvh", n)
    same_ids = float((d_ids == ref_ids).double().mean().item())
    same_t = bool(torch.equal(d_t, ref_t))
    return {"value": n / (ms * 1e-3) / 1e6, "unit": "Mrays/s per GPU", "kernel_ms": ms, "host_scene_build_s": t_build,
            "nodes_per_ray": nodes_per_ray, "prim_tests_per_ray": prims_per_ray, "algorithmic_bytes_per_ray": bytes_per_ray,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "traffic_source": tsrc},
            "ids_equal_to_reference_tree": same_ids, "t_bit_equal_to_reference_tree": same_t,
            "note": "tree built by izpi_bvh4_build (device LBVH), reported separately: the parity claim and the headline roofline are on the reference's tree"}


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


def pin_numa(local_rank):
    """Bind this rank's host threads and future page allocations to the NUMA node of its GPU, so that the pinned ray buffers
    of the e2e leg sit next to the PCIe root of the device that reads them (round 1: eight ranks streaming from unplaced
    pinned memory fell from 31 to 17 GB/s per GPU).  Linux sysfs + sched_setaffinity; a no-op where the information is absent."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        busid = busid.lower()
        if len(busid.split(":")[0]) == 8:
            busid = busid[4:]
        node = int(open(f"/sys/bus/pci/devices/{busid}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)  # first touch of the pinned buffers below then lands on this node
        return {"numa_node": node, "cpus": len(ids)}
    except Exception as e:
        return {"numa_node": None, "why": str(e)[:80]}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="izpi_b200")
    ap.add_argument("--rays", type=int, default=N_RAYS, help="rays per step per GPU (default = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-4k", action="store_true", help="skip the 4K / 11.5M-triangle renders (config 5)")
    ap.add_argument("--no-render", action="store_true", help="closest-hit headline only")
    ap.add_argument("--no-render-stats", action="store_true", help="skip the per-config roofline passes")
    ap.add_argument("--quick", action="store_true", help="render configs at 1/16 of BASELINE's sample counts (development sweeps)")
    ap.add_argument("--render-configs", default="", help="comma list out of 1,3,4,5,5-lbvh (default: all)")
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
    numa = pin_numa(local_rank) if world > 1 else {"numa_node": None, "why": "single rank: all host cores serve one GPU"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)
    n = args.rays

    ctx = cuda.Context(local_rank)
    sc, lo, hi = build_inputs(rank)
    hs = None
    if rank == 0:  # one scene build per box; the other ranks receive the flattened arrays over NCCL
        hs = cuda.HostScene(sc)
        ctx.upload(hs)
    from izpi_b200 import render as _render
    _render.replicate_scene(ctx, 0)
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

    # what the link gives: one plain pinned-host -> device copy of the step's rays (the e2e leg cannot beat it)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(d_org)
    scratch.copy_(h_org, non_blocking=True)
    p0.record()
    for _ in range(3):
        scratch.copy_(h_org, non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    pcie_h2d_gbs = 3 * h_org.numel() * 8 / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del scratch

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
    del d_ids32, d_t32

    # FLOP side of the traversal roofline (north_star: "the slower of bytes per ray at HBM bandwidth and intersection FLOPs at
    # fp64/fp32 peak"): measured dependent-FMA peaks of this GPU, SURVEY.md §8(d)'s operation counts per visit / per test
    fp32_peak, fp64_peak = ctx.fma_peak(False), ctx.fma_peak(True)
    fp32_flop_per_ray, fp64_flop_per_ray = 100.0 * nodes_per_ray, 50.0 * prims_per_ray

    cpu_info = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (rank 0, bounded sample)
        threads = os.cpu_count() or 1
        n_cpu = 1 << 22
        v, ost, (oi, ot, oo, od) = cpu_baseline(sc, lo, hi, n_cpu, threads)
        gi, gt = ctx.trace_closest(oo, od)
        cpu_info = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                    "sample": f"first {n_cpu} rays of the step, {threads} host threads; C++ restatement of the Go CPU path",
                    "parity_on_sample": bool(np.array_equal(gi, oi) and gt.tobytes() == ot.tobytes())}
        del oi, ot, oo, od, gi, gt

    lbvh_info = None
    if world == 1:
        lbvh_info = bench_lbvh_trace(ctx, cuda, scenes, n, d_org, d_dir, org, d, stream, args.steps, d_ids, d_t)
    del hs, d_org, d_dir, d_ids, d_t, h_org, h_dir, h_ids, h_t, np_org, np_dir, np_ids, np_t, org, d
    torch.cuda.empty_cache()

    render_info = [] if args.no_render else bench_render(ctx, cuda, scenes, world, rank, barrier, args)

    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, kernel_ms = tt.tolist()
        numas = [None] * world
        dist.all_gather_object(numas, numa)
    else:
        numas = [numa]
    if rank == 0:
        hbm, peak_src = peaks()
        value = world * n * args.steps / (total_ms * 1e-3) / 1e6
        achieved = bytes_per_ray * n / (kernel_ms * 1e-3) / 1e9
        traffic, tkernel, tsrc = ncu_traffic("trace_headline", n)
        e2e_value = world * n * args.steps / (e2e_ms * 1e-3) / 1e6
        line = {
            "metric": "Mrays/s closest-hit (1M tris, incoherent)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n, "l2": "inputs larger than L2 (805 MB of rays per step)",
                       "mode": "exact (reference traversal order, fp32 slab test, fp64 primitives)",
                       "nodes_per_ray": nodes_per_ray, "prim_tests_per_ray": prims_per_ray,
                       "algorithmic_bytes_per_ray": bytes_per_ray,
                       "render_spp": "BASELINE sample counts" if not args.quick else "1/16 of BASELINE sample counts (--quick)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "traffic_source": tsrc, "kernel": tkernel or "trace_g2_kernel<false,false,48>", "kernel_ms": kernel_ms,
                         "peak_source": peak_src},
            "roofline_flops": {"fp32_tflops_peak": fp32_peak, "fp64_tflops_peak": fp64_peak, "fp32_flop_per_ray": fp32_flop_per_ray,
                               "fp64_flop_per_ray": fp64_flop_per_ray,
                               "mrays_per_s_at_flop_peak": 1e-6 / (fp32_flop_per_ray / (fp32_peak * 1e12) + fp64_flop_per_ray / (fp64_peak * 1e12)),
                               "mrays_per_s_at_hbm_peak": hbm * 1e3 / bytes_per_ray,
                               "note": "the byte side is the slower (binding) one; peaks measured by izpi_debug_fma_peak on this GPU"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 12,
                    "h2d_gbs_per_rank": e2e_value / world * 48e6 / 1e9, "plain_pinned_h2d_copy_gbs_rank0": pcie_h2d_gbs, "host_numa": numas},
            "gpu_launches": int(launches),
            "render": render_info,
            "lbvh_tree": lbvh_info,
            "fp32_mode": fp32_info,
            "clocks": sampler.summary(),
        }
        if cpu_info:
            line["cpu_baseline"] = cpu_info
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
