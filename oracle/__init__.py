"""ctypes binding of liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under izpi_b200/ may.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from izpi_b200.scene import NODE_DTYPE, SceneSpec, SceneSpecC

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the C++ restatement (g++ is in the image; takes a few seconds)."""
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".cpp", ".hpp", ".h"))]
    srcs.append(os.path.join(_DIR, "..", "include", "izpi_scene.h"))
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nodes", "tris", "spheres", "others", "rays")]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("sampler", C.c_int32), ("rng_mode", C.c_int32), ("flavour", C.c_int32), ("threads", C.c_int32),
                ("seed", C.c_uint64), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("epilogue", C.c_int32), ("libm_jitter", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_ray_aabb4.restype = C.c_uint8
        L.oracle_ray_aabb4.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.oracle_conservative_float32_min.restype = C.c_float
        L.oracle_conservative_float32_min.argtypes = [C.c_double]
        L.oracle_conservative_float32_max.restype = C.c_float
        L.oracle_conservative_float32_max.argtypes = [C.c_double]
        L.oracle_lcg_next.restype = C.c_double
        L.oracle_lcg_next.argtypes = [C.POINTER(C.c_uint64)]
        L.oracle_sample_wavelength.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_cie_values.argtypes = [C.c_double, C.c_void_p]
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.POINTER(SceneSpecC)]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_scene_num_nodes.restype = C.c_int32
        L.oracle_scene_num_nodes.argtypes = [C.c_void_p]
        L.oracle_scene_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_scene_num_lights.restype = C.c_int32
        L.oracle_scene_num_lights.argtypes = [C.c_void_p]
        L.oracle_scene_lights.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_prim_bbox.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_triangle_fields.restype = C.c_int
        L.oracle_triangle_fields.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_spectral_texture_value.restype = C.c_double
        L.oracle_spectral_texture_value.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.oracle_hit.restype = C.c_int32
        L.oracle_hit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.oracle_trace.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_void_p, C.c_void_p, C.POINTER(Stats), C.c_int]
        L.oracle_render.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_render_tile.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_firefly_rejection.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.oracle_xyz_to_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double]
        L.oracle_apply_displacement.restype = C.c_int64
        L.oracle_apply_displacement.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_double,
                                                C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_displacement_tessellate.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_displacement_apply_tessellation.restype = C.c_int64
        L.oracle_displacement_apply_tessellation.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_displacement_apply_displacement.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.oracle_tiles.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def ray_aabb4(flavour, org, inv, bounds, tmax) -> int:
    o = np.ascontiguousarray(org, dtype=np.float32)
    i = np.ascontiguousarray(inv, dtype=np.float32)
    b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(24)
    return int(lib().oracle_ray_aabb4(flavour, o.ctypes.data, i.ctypes.data, b.ctypes.data, np.float32(tmax)))


def tiles(sx, sy):
    a, b = C.c_int32(), C.c_int32()
    lib().oracle_tiles(sx, sy, C.byref(a), C.byref(b))
    return a.value, b.value


def apply_displacement(tris15, materials, pixels, dmin, dmax, per_triangle=True):
    """displacement.ApplyDisplacementMap (displacement.go:145) as restated by the oracle."""
    t = np.ascontiguousarray(tris15, dtype=np.float64).reshape(-1, 15)
    mats = np.ascontiguousarray(materials, dtype=np.int32).reshape(-1)
    px = np.ascontiguousarray(pixels, dtype=np.float64)
    ot, om = C.c_void_p(), C.c_void_p()
    k = lib().oracle_apply_displacement(len(t), t.ctypes.data, mats.ctypes.data, px.shape[1], px.shape[0], px.ctypes.data, float(dmin),
                                        float(dmax), int(per_triangle), C.byref(ot), C.byref(om))
    out = np.ctypeslib.as_array(C.cast(ot, C.POINTER(C.c_double)), shape=(max(k, 1) * 15,))[: k * 15].reshape(k, 15).copy()
    outm = np.ctypeslib.as_array(C.cast(om, C.POINTER(C.c_int32)), shape=(max(k, 1),))[:k].copy()
    lib().oracle_free(ot)
    lib().oracle_free(om)
    return out, outm


def displacement_tessellate(tri15):
    """tessellate() (displacement.go:36-103)."""
    t = np.ascontiguousarray(tri15, dtype=np.float64).reshape(15)
    out = np.zeros((4, 15))
    lib().oracle_displacement_tessellate(t.ctypes.data, out.ctypes.data)
    return out


def displacement_apply_tessellation(tris15, max_delta_u, max_delta_v, rgb, dmin, dmax, threshold) -> int:
    """applyTessellation() over a constant texture with explicit UV limits (displacement.go:188-218)."""
    t = np.ascontiguousarray(tris15, dtype=np.float64).reshape(-1, 15)
    c = np.ascontiguousarray(rgb, dtype=np.float64)
    return int(lib().oracle_displacement_apply_tessellation(len(t), t.ctypes.data, max_delta_u, max_delta_v, c.ctypes.data, dmin, dmax, threshold))


def displacement_apply_displacement(tris15, rgb, dmin, dmax):
    """applyDisplacement() over a constant texture (displacement.go:201-280)."""
    t = np.ascontiguousarray(tris15, dtype=np.float64).reshape(-1, 15)
    c = np.ascontiguousarray(rgb, dtype=np.float64)
    out = np.zeros_like(t)
    lib().oracle_displacement_apply_displacement(len(t), t.ctypes.data, c.ctypes.data, dmin, dmax, out.ctypes.data)
    return out


class OracleScene:
    """The reference's scene.Scene as rebuilt by the oracle from a SceneSpec."""

    def __init__(self, spec: SceneSpec):
        self._c = spec.to_c()
        self._spec = spec
        self._h = lib().oracle_scene_create(C.byref(self._c))
        if not self._h:
            raise ValueError("oracle_scene_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_scene_destroy(self._h)
            self._h = None

    def bvh(self):
        n = lib().oracle_scene_num_nodes(self._h)
        nodes = np.zeros(n, dtype=NODE_DTYPE)
        perm = np.zeros(self._c.n_prims if n else 0, dtype=np.int32)
        lib().oracle_scene_bvh(self._h, nodes.ctypes.data, perm.ctypes.data)
        return nodes, perm

    def lights(self):
        n = lib().oracle_scene_num_lights(self._h)
        ids = np.zeros(n, dtype=np.int32)
        lib().oracle_scene_lights(self._h, ids.ctypes.data)
        return ids

    def prim_bbox(self, i):
        out = np.zeros(6)
        lib().oracle_prim_bbox(self._h, i, out.ctypes.data)
        return out

    def triangle_fields(self, i):
        out = np.zeros(22)
        if lib().oracle_triangle_fields(self._h, i, out.ctypes.data) != 0:
            raise TypeError("not a triangle")
        return dict(edge1=out[0:3], edge2=out[3:6], normal=out[6:9], tangent=out[9:12], bitangent=out[12:15],
                    area=out[15], bbmin=out[16:19], bbmax=out[19:22])

    def hit(self, org, direction, tmin=0.0, tmax=np.finfo(np.float64).max, flavour=0):
        o = np.ascontiguousarray(org, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.zeros(9)
        pid = lib().oracle_hit(self._h, flavour, o.ctypes.data, d.ctypes.data, tmin, tmax, out.ctypes.data)
        if pid < 0:
            return None
        return dict(prim=pid, t=out[0], u=out[1], v=out[2], p=out[3:6].copy(), normal=out[6:9].copy())

    def trace(self, org, direction, tmin=0.001, tmax=np.finfo(np.float64).max, flavour=0, threads=None, stats=False):
        o = np.ascontiguousarray(org, dtype=np.float64)
        d = np.ascontiguousarray(direction, dtype=np.float64)
        n = o.shape[0]
        ids = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float64)
        st = Stats()
        lib().oracle_trace(self._h, flavour, n, o.ctypes.data, d.ctypes.data, tmin, tmax, ids.ctypes.data,
                           t.ctypes.data, C.byref(st) if stats else None, threads or os.cpu_count() or 1)
        if stats:
            return ids, t, {k: getattr(st, k) for k in ("nodes", "tris", "spheres", "others", "rays")}
        return ids, t

    def spectral_texture_value(self, tex, lam):
        return lib().oracle_spectral_texture_value(self._h, tex, lam)

    def render(self, width, height, spp, max_depth=50, sampler=0, rng_mode=0, flavour=0, threads=None, seed=1,
               window=None, epilogue=True, libm_jitter=0):
        x0, y0, x1, y1 = window if window is not None else (0, 0, width - 1, height - 1)
        p = RenderParams(width=width, height=height, spp=spp, max_depth=max_depth, sampler=sampler,
                         rng_mode=rng_mode, flavour=flavour, threads=threads or os.cpu_count() or 1, seed=seed,
                         x0=x0, y0=y0, x1=x1, y1=y1, epilogue=int(epilogue), libm_jitter=int(libm_jitter))
        canvas = np.zeros((height, width, 4), dtype=np.float64)
        rays = C.c_uint64()
        lib().oracle_render(self._h, C.byref(p), canvas.ctypes.data, C.byref(rays))
        return canvas, rays.value

    def render_tile(self, width, height, spp, x0, y0, x1, y1, strip_height=1, max_depth=50, sampler=0, rng_mode=0, seed=1):
        """worker.RenderTile (internal/worker/render.go:17-75): the rows a worker streams back for one tile."""
        p = RenderParams(width=width, height=height, spp=spp, max_depth=max_depth, sampler=sampler, rng_mode=rng_mode, flavour=0,
                         threads=1, seed=seed, x0=0, y0=0, x1=0, y1=0, epilogue=0, libm_jitter=0)
        rows = np.zeros((y1 - y0 + 1, strip_height * 4 * (x1 - x0 + 1)), dtype=np.float64)
        rays = C.c_uint64()
        lib().oracle_render_tile(self._h, C.byref(p), strip_height, x0, y0, x1, y1, rows.ctypes.data, C.byref(rays))
        return rows, rays.value
