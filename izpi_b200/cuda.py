"""ctypes binding of libizpi_cuda.so -- the Python twin of the cgo package `internal/cuda`
(INTEGRATION.md).  Thin by design: every call maps 1:1 onto an entry point of
include/izpi_cuda.h / include/izpi_host.h; errors become exceptions carrying izpi_last_error().

There is no CPU fallback: if the library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scene import NODE_DTYPE, SceneSpec, SceneSpecC

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IZPI_LIB_PATH") or os.path.join(_HERE, "libizpi_cuda.so")  # override: kernel experiments

OK, EINVAL, ECUDA, ESTATE = 0, -1, -2, -3
TRACE_EXACT, TRACE_FP32 = 0, 1
SAMPLER_COLOUR, SAMPLER_SPECTRAL, SAMPLER_ALBEDO, SAMPLER_NORMAL = 0, 1, 2, 3


class IzpiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"izpi error {code}: {msg}")
        self.code = code


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("prim_tests", C.c_uint64), ("kernel_ms", C.c_double)]


class RenderConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("sampler", C.c_int32), ("sample_offset", C.c_int32), ("sample_count", C.c_int32), ("flags", C.c_int32),
                ("background", C.c_double * 3), ("bg_wavelengths", C.c_void_p), ("bg_values", C.c_void_p),
                ("n_bg", C.c_int32), ("reserved2", C.c_int32), ("seed", C.c_uint64)]


class RenderStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("prim_tests", C.c_uint64), ("extend_launches", C.c_uint64), ("material_bins", C.c_uint64),
                ("extend_ms", C.c_double), ("shade_ms", C.c_double), ("other_ms", C.c_double)]


RENDER_STATS, RENDER_TIMING = 1, 2

# every symbol include/izpi_cuda.h and include/izpi_host.h declare (checked by tests/test_abi.py)
EXPORTS = [
    "izpi_last_error", "izpi_version", "izpi_ctx_create", "izpi_ctx_num_devices", "izpi_ctx_destroy", "izpi_scene_upload",
    "izpi_scene_image_size", "izpi_scene_image_export", "izpi_scene_image_adopt", "izpi_scene_image_commit",
    "izpi_render_tiles_shared", "izpi_render_get_stats",
    "izpi_trace_closest", "izpi_trace_closest_device", "izpi_launch_count", "izpi_render_setup", "izpi_render_tiles", "izpi_render_tile_rows",
    "izpi_render_canvas_device", "izpi_render_finish", "izpi_debug_ray_aabb4", "izpi_debug_hit", "izpi_debug_fma_peak", "izpi_displace", "izpi_displace_fetch", "izpi_bvh4_build", "izpi_bvh4_build_fetch",
    "izpi_host_scene_create", "izpi_host_scene_destroy", "izpi_host_scene_num_nodes", "izpi_host_scene_bvh",
    "izpi_host_scene_num_lights", "izpi_host_scene_lights", "izpi_host_scene_desc", "izpi_host_scene_upload",
    "izpi_host_tiles", "izpi_host_render", "izpi_host_claim_tiles", "izpi_host_walk_grid_spiral",
    # include/izpi_proto.h
    "izpi_proto_scene_parse", "izpi_proto_scene_append_triangles", "izpi_proto_scene_to_scene", "izpi_proto_scene_spec",
    "izpi_proto_scene_name", "izpi_proto_scene_colour_representation", "izpi_proto_scene_total_triangles",
    "izpi_proto_scene_stream_triangles", "izpi_proto_scene_num_parsed_triangles", "izpi_proto_scene_background",
    "izpi_proto_scene_num_images", "izpi_proto_scene_image_filename", "izpi_proto_scene_destroy",
]

_lib = None


def lib():
    """Load the shared library (built in-tree by izpi_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IzpiError(ECUDA, f"{LIB_PATH} is missing: run `python -m izpi_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.izpi_last_error.restype = C.c_char_p
    L.izpi_version.restype = C.c_int
    L.izpi_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    L.izpi_ctx_destroy.argtypes = [C.c_void_p]
    L.izpi_ctx_destroy.restype = None
    L.izpi_scene_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.izpi_trace_closest.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                     C.c_void_p, C.c_void_p, C.POINTER(TraceStats)]
    L.izpi_trace_closest_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_launch_count.argtypes = [C.c_void_p]
    L.izpi_launch_count.restype = C.c_uint64
    L.izpi_debug_ray_aabb4.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_debug_hit.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    L.izpi_debug_fma_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.izpi_displace.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_double,
                                C.c_int, C.POINTER(C.c_int64)]
    L.izpi_displace_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_bvh4_build.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.izpi_bvh4_build_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_render_setup.argtypes = [C.c_void_p, C.POINTER(RenderConfig)]
    L.izpi_render_tiles.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.izpi_render_tiles_shared.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.izpi_render_get_stats.argtypes = [C.c_void_p, C.POINTER(RenderStats)]
    L.izpi_ctx_num_devices.argtypes = [C.c_void_p]
    L.izpi_scene_image_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    L.izpi_scene_image_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_scene_image_adopt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.izpi_scene_image_commit.argtypes = [C.c_void_p]
    L.izpi_render_tile_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.izpi_render_canvas_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.izpi_render_finish.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.izpi_host_scene_create.argtypes = [C.POINTER(SceneSpecC), C.c_int, C.POINTER(C.c_void_p)]
    L.izpi_host_scene_destroy.argtypes = [C.c_void_p]
    L.izpi_host_scene_destroy.restype = None
    L.izpi_host_scene_num_nodes.argtypes = [C.c_void_p]
    L.izpi_host_scene_num_nodes.restype = C.c_int32
    L.izpi_host_scene_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.izpi_host_scene_num_lights.argtypes = [C.c_void_p]
    L.izpi_host_scene_num_lights.restype = C.c_int32
    L.izpi_host_scene_lights.argtypes = [C.c_void_p, C.c_void_p]
    L.izpi_host_scene_desc.argtypes = [C.c_void_p, C.c_void_p]
    L.izpi_host_scene_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.izpi_host_tiles.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.izpi_host_tiles.restype = None
    L.izpi_host_walk_grid_spiral.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
    L.izpi_host_walk_grid_spiral.restype = None
    L.izpi_host_claim_tiles.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.izpi_host_render.argtypes = [C.c_void_p, C.POINTER(RenderConfig), C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.POINTER(C.c_uint64)]
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        raise IzpiError(rc, lib().izpi_last_error().decode("utf-8", "replace"))


def tiles(size_x: int, size_y: int):
    """common.Tiles (common/tiles.go:6-24)."""
    a, b = C.c_int32(), C.c_int32()
    lib().izpi_host_tiles(size_x, size_y, C.byref(a), C.byref(b))
    return a.value, b.value


class HostScene:
    """scene.Scene as built by the host runtime (izpi_host_scene_create)."""

    def __init__(self, spec, threads: int | None = None):
        """spec: a scene.SceneSpec or a proto.ProtoScene (anything with to_c() -> izpi_scene_spec)."""
        self._spec = spec
        self._c = spec.to_c()
        h = C.c_void_p()
        check(lib().izpi_host_scene_create(C.byref(self._c), threads or os.cpu_count() or 1, C.byref(h)))
        self._h = h
        self.n_prims = self._c.n_prims

    def __del__(self):
        if getattr(self, "_h", None):
            lib().izpi_host_scene_destroy(self._h)
            self._h = None

    def bvh(self):
        """(BVH4.Nodes, permutation) -- exported fields of hitable.BVH4 (bvh4.go:42-47)."""
        n = lib().izpi_host_scene_num_nodes(self._h)
        nodes = np.zeros(n, dtype=NODE_DTYPE)
        perm = np.zeros(self.n_prims if n else 0, dtype=np.int32)
        check(lib().izpi_host_scene_bvh(self._h, nodes.ctypes.data, perm.ctypes.data))
        return nodes, perm

    def lights(self):
        n = lib().izpi_host_scene_num_lights(self._h)
        ids = np.zeros(n, dtype=np.int32)
        if n:
            check(lib().izpi_host_scene_lights(self._h, ids.ctypes.data))
        return ids


class Context:
    """One GPU, or a group of GPUs driven from this process (izpi_ctx; `device` may be an int or a list of ints)."""

    def __init__(self, device=0):
        devs = [int(device)] if isinstance(device, (int, np.integer)) else [int(d) for d in device]
        h = C.c_void_p()
        ids = (C.c_int * len(devs))(*devs)
        check(lib().izpi_ctx_create(len(devs), ids, C.byref(h)))
        self._h = h
        self.device = devs[0]
        self.devices = devs
        self._scene = None

    def close(self):
        if getattr(self, "_h", None):
            lib().izpi_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def upload(self, scene: HostScene):
        check(lib().izpi_host_scene_upload(scene._h, self._h))
        self._scene = scene

    # ---- scene image: replicate an uploaded scene onto another device without rebuilding it ----------------------
    def scene_image(self):
        """(header bytes, [block sizes], [device pointers]) of the uploaded scene."""
        hb, nb = C.c_uint64(), C.c_int32()
        check(lib().izpi_scene_image_size(self._h, C.byref(hb), C.byref(nb)))
        header = np.zeros(hb.value, dtype=np.uint8)
        sizes = (C.c_uint64 * max(1, nb.value))()
        ptrs = (C.c_void_p * max(1, nb.value))()
        check(lib().izpi_scene_image_export(self._h, header.ctypes.data, sizes, ptrs))
        return header, [int(sizes[i]) for i in range(nb.value)], [int(ptrs[i] or 0) for i in range(nb.value)]

    def scene_adopt(self, header: np.ndarray, n_blocks: int):
        """Allocate the blocks of somebody else's scene image here; returns their device pointers (to be filled)."""
        header = np.ascontiguousarray(header, dtype=np.uint8)
        ptrs = (C.c_void_p * max(1, n_blocks))()
        check(lib().izpi_scene_image_adopt(self._h, header.ctypes.data, header.nbytes, ptrs))
        return [int(ptrs[i] or 0) for i in range(n_blocks)]

    def scene_commit(self):
        check(lib().izpi_scene_image_commit(self._h))
        self._scene = None

    @property
    def launches(self) -> int:
        return int(lib().izpi_launch_count(self._h))

    # ---- hitable.Hitable.Hit, batched -------------------------------------------------------
    def trace_closest(self, org, direction, tmin=0.001, tmax=np.finfo(np.float64).max, mode=TRACE_EXACT, stats=False,
                      out_ids=None, out_t=None):
        o = np.ascontiguousarray(org, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        n = o.shape[0] if o.ndim == 2 else 0
        if o.shape != d.shape or (n and o.shape[1] != 3):
            raise IzpiError(EINVAL, "org/dir must both be (n, 3)")
        ids = out_ids if out_ids is not None else np.empty(n, dtype=np.int32)
        t = out_t if out_t is not None else np.empty(n, dtype=np.float64)
        st = TraceStats()
        check(lib().izpi_trace_closest(self._h, n, o.ctypes.data, d.ctypes.data, tmin, tmax, mode, ids.ctypes.data,
                                       t.ctypes.data, C.byref(st) if stats else None))
        if stats:
            return ids, t, dict(rays=st.rays, nodes=st.nodes_visited, prims=st.prim_tests, kernel_ms=st.kernel_ms)
        return ids, t

    def trace_closest_device(self, n, d_org, d_dir, d_ids, d_t, tmin=0.001, tmax=np.finfo(np.float64).max,
                             mode=TRACE_EXACT, stream=0):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        check(lib().izpi_trace_closest_device(self._h, n, d_org, d_dir, tmin, tmax, mode, d_ids, d_t, stream or None))

    # ---- displacement.ApplyDisplacementMap ---------------------------------------------------
    def apply_displacement(self, tris15, materials, pixels, dmin, dmax, per_triangle=True):
        """tris15 (n,15) = v0 v1 v2 u0 v0 u1 v1 u2 v2; pixels (H,W,4) fp64 with the height in the blue channel.
        Returns (out_tris15 (m,15), out_materials (m,))."""
        t = np.ascontiguousarray(tris15, dtype=np.float64).reshape(-1, 15)
        mats = np.ascontiguousarray(materials, dtype=np.int32).reshape(-1)
        px = np.ascontiguousarray(pixels, dtype=np.float64)
        n_out = C.c_int64()
        check(lib().izpi_displace(self._h, len(t), t.ctypes.data, mats.ctypes.data, px.shape[1], px.shape[0], px.ctypes.data,
                                  float(dmin), float(dmax), int(per_triangle), C.byref(n_out)))
        out = np.empty((n_out.value, 15), dtype=np.float64)
        om = np.empty(n_out.value, dtype=np.int32)
        check(lib().izpi_displace_fetch(self._h, out.ctypes.data, om.ctypes.data))
        return out, om

    # ---- hitable.NewBVH4 on the device -----------------------------------------------------------
    def build_bvh4(self, boxes6):
        """boxes6 (n,6) = min.xyz max.xyz of every hitable.  Returns (BVH4.Nodes, permutation)."""
        b = np.ascontiguousarray(boxes6, dtype=np.float64).reshape(-1, 6)
        nn = C.c_int32()
        check(lib().izpi_bvh4_build(self._h, len(b), b.ctypes.data, C.byref(nn)))
        nodes = np.zeros(nn.value, dtype=NODE_DTYPE)
        perm = np.zeros(len(b), dtype=np.int32)
        check(lib().izpi_bvh4_build_fetch(self._h, nodes.ctypes.data, perm.ctypes.data))
        return nodes, perm

    def fma_peak(self, fp64: bool) -> float:
        """Measured dependent-FMA TFLOP/s of the fp32 / fp64 vector pipe (the FLOP side of the traversal roofline)."""
        v = C.c_double()
        check(lib().izpi_debug_fma_peak(self._h, int(fp64), C.byref(v)))
        return v.value

    def debug_hit(self, org, direction, tmin=0.0, tmax=np.finfo(np.float64).max):
        """world.Hit with the full HitRecord: (ids, records (n, 9) = t u v p.xyz normal.xyz)."""
        o = np.ascontiguousarray(org, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
        ids = np.zeros(len(o), dtype=np.int32)
        out = np.zeros((len(o), 9), dtype=np.float64)
        check(lib().izpi_debug_hit(self._h, len(o), o.ctypes.data, d.ctypes.data, tmin, tmax, ids.ctypes.data, out.ctypes.data))
        return ids, out

    def debug_ray_aabb4(self, org, inv, bounds, tmax):
        o = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(inv, dtype=np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 24)
        t = np.ascontiguousarray(tmax, dtype=np.float32).reshape(-1)
        m = np.zeros(len(o), dtype=np.uint8)
        check(lib().izpi_debug_ray_aabb4(self._h, len(o), o.ctypes.data, i.ctypes.data, b.ctypes.data, t.ctypes.data, m.ctypes.data))
        return m

    # ---- render.New(...).Render() ------------------------------------------------------------
    def render(self, width, height, spp, max_depth=50, sampler=SAMPLER_COLOUR, seed=1, sample_offset=0, sample_count=None,
               tile_begin=0, tile_end=-1, finish=True):
        cfg = RenderConfig(width=width, height=height, spp=spp, max_depth=max_depth, sampler=sampler,
                           sample_offset=sample_offset, sample_count=spp if sample_count is None else sample_count, seed=seed)
        canvas = np.zeros((height, width, 4), dtype=np.float64)
        rays = C.c_uint64()
        check(lib().izpi_host_render(self._h, C.byref(cfg), tile_begin, tile_end, int(finish), canvas.ctypes.data, C.byref(rays)))
        return canvas, rays.value

    # ---- worker.RenderSetup / RenderTile ------------------------------------------------------
    def render_setup(self, width, height, spp, max_depth=50, sampler=SAMPLER_COLOUR, seed=1, background=(0.0, 0.0, 0.0),
                     bg_wavelengths=None, bg_values=None):
        """RenderSetupRequest (control.proto:56-68)."""
        cfg = RenderConfig(width=width, height=height, spp=spp, max_depth=max_depth, sampler=sampler, sample_offset=0, sample_count=spp, seed=seed)
        cfg.background[:] = [float(c) for c in background]
        keep = []
        if bg_wavelengths is not None and len(bg_wavelengths):
            w = np.ascontiguousarray(bg_wavelengths, dtype=np.float64)
            v = np.ascontiguousarray(bg_values, dtype=np.float64)
            keep = [w, v]
            cfg.bg_wavelengths, cfg.bg_values, cfg.n_bg = w.ctypes.data, v.ctypes.data, len(w)
        check(lib().izpi_render_setup(self._h, C.byref(cfg)))
        del keep

    def render_tile_rows(self, x0, y0, x1, y1, strip_height=1):
        """RenderTileRequest -> the rows of the RenderTileResponse stream: (y1-y0+1, strip_height*4*(x1-x0+1))."""
        rows = np.zeros((y1 - y0 + 1, strip_height * 4 * (x1 - x0 + 1)), dtype=np.float64)
        check(lib().izpi_render_tile_rows(self._h, strip_height, x0, y0, x1, y1, rows.ctypes.data))
        return rows

    def render_stats(self) -> dict:
        st = RenderStats()
        check(lib().izpi_render_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in RenderStats._fields_}

    def canvas_device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().izpi_render_canvas_device(self._h, C.byref(p)))
        return p.value

    def render_finish(self, width, height):
        canvas = np.zeros((height, width, 4), dtype=np.float64)
        rays = C.c_uint64()
        check(lib().izpi_render_finish(self._h, canvas.ctypes.data, C.byref(rays)))
        return canvas, rays.value
