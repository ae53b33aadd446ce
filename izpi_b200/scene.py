"""Host-side scene description (the data the reference's loaders hand to its constructors).

Mirrors ``include/izpi_scene.h`` with numpy structured dtypes + ctypes structs.  The builder
methods are named after the reference constructors they stand for
(internal/hitable/*.go, internal/material/*.go, internal/texture/*.go; protobuf order rules from
internal/transport/transport.go:551-566: triangles first, then spheres).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

# ---- enums (include/izpi_scene.h) -------------------------------------------------------
PRIM_TRIANGLE, PRIM_SPHERE, PRIM_XYRECT, PRIM_XZRECT, PRIM_YZRECT, PRIM_BOX = range(6)
WRAP_FLIP, WRAP_ROTATE_Y, WRAP_TRANSLATE = 1, 2, 4
TEX_CONSTANT, TEX_IMAGE = 0, 1
SPEC_GAUSSIAN, SPEC_TABULATED, SPEC_IMAGE = 0, 1, 2
MAT_LAMBERT, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_PBR = range(5)
WORLD_SLICE, WORLD_BVH4 = 0, 1
BVH_REFERENCE, BVH_DEVICE_LBVH = 0, 1

PRIM_DTYPE = np.dtype(
    [("type", "<i4"), ("material", "<i4"), ("wrap", "<i4"), ("reserved", "<i4"),
     ("p", "<f8", (15,)), ("rotate_y_deg", "<f8"), ("translate", "<f8", (3,))], align=True)
assert PRIM_DTYPE.itemsize == 168

NODE_DTYPE = np.dtype(
    [("min_x", "<f4", (4,)), ("min_y", "<f4", (4,)), ("min_z", "<f4", (4,)),
     ("max_x", "<f4", (4,)), ("max_y", "<f4", (4,)), ("max_z", "<f4", (4,)),
     ("child_index", "<i4", (4,)), ("primitive_count", "<i4", (4,))])
assert NODE_DTYPE.itemsize == 128


class TextureSpec(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("reserved", C.c_int32),
                ("color", C.c_double * 3), ("pixels", C.c_void_p)]


class SpectralTextureSpec(C.Structure):
    _fields_ = [("type", C.c_int32), ("n", C.c_int32), ("peak", C.c_double), ("centre", C.c_double),
                ("width", C.c_double), ("wavelengths", C.c_void_p), ("values", C.c_void_p)]


class MaterialSpec(C.Structure):
    _fields_ = [("type", C.c_int32), ("tex", C.c_int32), ("spectral_tex", C.c_int32),
                ("spectral_absorption_tex", C.c_int32), ("normal_tex", C.c_int32), ("roughness_tex", C.c_int32),
                ("metalness_tex", C.c_int32), ("compute_beer_lambert", C.c_int32),
                ("v", C.c_double * 3), ("s", C.c_double)]


class CameraSpec(C.Structure):
    _fields_ = [("look_from", C.c_double * 3), ("look_at", C.c_double * 3), ("vup", C.c_double * 3),
                ("vfov", C.c_double), ("aspect", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("time0", C.c_double), ("time1", C.c_double), ("exposure", C.c_double)]


class SceneSpecC(C.Structure):
    _fields_ = [("world_kind", C.c_int32), ("n_prims", C.c_int32), ("prims", C.c_void_p),
                ("n_materials", C.c_int32), ("n_textures", C.c_int32), ("materials", C.c_void_p),
                ("textures", C.c_void_p), ("n_spectral_textures", C.c_int32), ("reserved", C.c_int32),
                ("spectral_textures", C.c_void_p), ("camera", CameraSpec), ("bvh_seed", C.c_uint64),
                ("bvh_rand_zero", C.c_int32), ("bvh_builder", C.c_int32)]


def f32(x):
    """float32 provenance of every protobuf scalar (transport.proto:58-62 -> float64(...) in
    transport.go:606-622): round to fp32, then widen."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


class SceneSpec:
    """Accumulates primitives/materials/textures and produces an ``izpi_scene_spec``."""

    def __init__(self, world_kind=WORLD_BVH4, bvh_seed=12345, bvh_rand_zero=False, bvh_builder=BVH_REFERENCE):
        self.world_kind = world_kind
        self.bvh_builder = bvh_builder
        self.bvh_seed = bvh_seed
        self.bvh_rand_zero = bvh_rand_zero
        self._prim_chunks: list[np.ndarray] = []
        self.materials: list[MaterialSpec] = []
        self.textures: list[TextureSpec] = []
        self.spectral_textures: list[SpectralTextureSpec] = []
        self._keep: list = []  # arrays borrowed by the C structs
        self.camera = CameraSpec()
        self.exposure = 1.0

    # ---- textures -----------------------------------------------------------------------
    def constant_texture(self, rgb) -> int:  # texture.NewConstant
        t = TextureSpec(type=TEX_CONSTANT)
        t.color[:] = [float(c) for c in rgb]
        self.textures.append(t)
        return len(self.textures) - 1

    def image_texture(self, pixels: np.ndarray) -> int:  # texture.NewFromRawData(W, H, data)
        px = np.ascontiguousarray(pixels, dtype=np.float64)
        assert px.ndim == 3 and px.shape[2] == 4
        self._keep.append(px)
        t = TextureSpec(type=TEX_IMAGE, width=px.shape[1], height=px.shape[0], pixels=px.ctypes.data)
        self.textures.append(t)
        return len(self.textures) - 1

    def spectral_gaussian(self, peak, centre, width) -> int:  # texture.NewSpectralConstant
        self.spectral_textures.append(SpectralTextureSpec(type=SPEC_GAUSSIAN, n=0, peak=peak, centre=centre, width=width))
        return len(self.spectral_textures) - 1

    def spectral_tabulated(self, wavelengths, values) -> int:  # texture.NewSpectralConstantFromSPD
        w = np.ascontiguousarray(wavelengths, dtype=np.float64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert w.shape == v.shape
        self._keep += [w, v]
        self.spectral_textures.append(SpectralTextureSpec(type=SPEC_TABULATED, n=len(w), wavelengths=w.ctypes.data, values=v.ctypes.data))
        return len(self.spectral_textures) - 1

    def spectral_image(self, image_tex: int) -> int:  # texture.NewSpectralImageFromImage over an image texture
        assert self.textures[image_tex].type == TEX_IMAGE
        self.spectral_textures.append(SpectralTextureSpec(type=SPEC_IMAGE, n=image_tex))
        return len(self.spectral_textures) - 1

    def spectral_neutral(self, reflectance) -> int:  # texture.NewSpectralNeutral (spectral_constant.go:47-62)
        w = np.arange(380.0, 751.0, 10.0)
        return self.spectral_tabulated(w, np.full_like(w, float(reflectance)))

    # ---- materials ----------------------------------------------------------------------
    def _mat(self, **kw) -> int:
        m = MaterialSpec(type=kw.pop("type"), tex=-1, spectral_tex=-1, spectral_absorption_tex=-1, normal_tex=-1,
                         roughness_tex=-1, metalness_tex=-1, compute_beer_lambert=0)
        v = kw.pop("v", (0.0, 0.0, 0.0))
        m.v[:] = [float(c) for c in v]
        for k, val in kw.items():
            setattr(m, k, val)
        self.materials.append(m)
        return len(self.materials) - 1

    def lambertian(self, tex) -> int:  # material.NewLambertian
        return self._mat(type=MAT_LAMBERT, tex=tex)

    def spectral_lambertian(self, spectral_tex) -> int:  # material.NewSpectralLambertian
        return self._mat(type=MAT_LAMBERT, spectral_tex=spectral_tex)

    def metal(self, albedo, fuzz) -> int:  # material.NewMetal
        return self._mat(type=MAT_METAL, v=albedo, s=float(fuzz))

    def dielectric(self, ref_idx) -> int:  # material.NewDielectric
        return self._mat(type=MAT_DIELECTRIC, s=float(ref_idx))

    def colored_dielectric(self, ref_idx, absorption) -> int:  # material.NewColoredDielectric
        return self._mat(type=MAT_DIELECTRIC, s=float(ref_idx), v=absorption, compute_beer_lambert=1)

    def spectral_dielectric(self, refidx_tex, absorption_tex=-1, compute_beer_lambert=False) -> int:
        # material.NewSpectralDielectric / NewSpectralColoredDielectric
        return self._mat(type=MAT_DIELECTRIC, spectral_tex=refidx_tex, spectral_absorption_tex=absorption_tex,
                         compute_beer_lambert=int(compute_beer_lambert))

    def diffuse_light(self, tex) -> int:  # material.NewDiffuseLight
        return self._mat(type=MAT_DIFFUSE_LIGHT, tex=tex)

    def spectral_diffuse_light(self, spectral_tex) -> int:  # material.NewSpectralDiffuseLight
        return self._mat(type=MAT_DIFFUSE_LIGHT, spectral_tex=spectral_tex)

    def pbr(self, albedo, normal=-1, roughness=-1, metalness=-1, spectral_albedo=-1) -> int:
        # material.NewPBR / NewPBRWithSpectralAlbedo (transport.go:209-250: spectral scenes wrap the albedo image)
        return self._mat(type=MAT_PBR, tex=albedo, normal_tex=normal, roughness_tex=roughness, metalness_tex=metalness,
                         spectral_tex=spectral_albedo)

    # ---- primitives ---------------------------------------------------------------------
    def _new(self, n) -> np.ndarray:
        a = np.zeros(n, dtype=PRIM_DTYPE)
        self._prim_chunks.append(a)
        return a

    def triangles(self, verts, material, uvs=None):
        """verts (n,3,3) = v0,v1,v2 ; uvs (n,3,2) = (u0,v0),(u1,v1),(u2,v2) ; material int or (n,)."""
        verts = np.asarray(verts, dtype=np.float64)
        a = self._new(len(verts))
        a["type"] = PRIM_TRIANGLE
        a["material"] = material
        a["p"][:, :9] = verts.reshape(-1, 9)
        if uvs is not None:
            a["p"][:, 9:15] = np.asarray(uvs, dtype=np.float64).reshape(-1, 6)
        return a

    def sphere(self, centre, radius, material):
        a = self._new(1)
        a["type"] = PRIM_SPHERE
        a["material"] = material
        a["p"][0, :4] = [*centre, radius]
        return a

    def rect(self, kind, a0, a1, b0, b1, k, material, flip=False):
        a = self._new(1)
        a["type"] = kind
        a["material"] = material
        a["p"][0, :5] = [a0, a1, b0, b1, k]
        a["wrap"] = WRAP_FLIP if flip else 0
        return a

    def box(self, p0, p1, material, rotate_y=None, translate=None):
        a = self._new(1)
        a["type"] = PRIM_BOX
        a["material"] = material
        a["p"][0, :6] = [*p0, *p1]
        if rotate_y is not None:
            a["wrap"] |= WRAP_ROTATE_Y
            a["rotate_y_deg"] = rotate_y
        if translate is not None:
            a["wrap"] |= WRAP_TRANSLATE
            a["translate"][0] = translate
        return a

    def set_camera(self, look_from, look_at, vup, vfov, aspect, aperture=0.0, focus_dist=10.0, time0=0.0, time1=1.0,
                   exposure=1.0):
        c = self.camera
        c.look_from[:] = [float(x) for x in look_from]
        c.look_at[:] = [float(x) for x in look_at]
        c.vup[:] = [float(x) for x in vup]
        c.vfov, c.aspect, c.aperture, c.focus_dist = float(vfov), float(aspect), float(aperture), float(focus_dist)
        c.time0, c.time1, c.exposure = float(time0), float(time1), float(exposure)

    # ---- finalise -----------------------------------------------------------------------
    @property
    def prims(self) -> np.ndarray:
        if len(self._prim_chunks) != 1:
            merged = np.concatenate(self._prim_chunks) if self._prim_chunks else np.zeros(0, dtype=PRIM_DTYPE)
            self._prim_chunks = [merged]
        return self._prim_chunks[0]

    def to_c(self) -> SceneSpecC:
        prims = np.ascontiguousarray(self.prims)
        mats = (MaterialSpec * max(1, len(self.materials)))(*self.materials)
        texs = (TextureSpec * max(1, len(self.textures)))(*self.textures)
        stex = (SpectralTextureSpec * max(1, len(self.spectral_textures)))(*self.spectral_textures)
        s = SceneSpecC(world_kind=self.world_kind, n_prims=len(prims), prims=prims.ctypes.data,
                       n_materials=len(self.materials), n_textures=len(self.textures),
                       materials=C.addressof(mats), textures=C.addressof(texs),
                       n_spectral_textures=len(self.spectral_textures), spectral_textures=C.addressof(stex),
                       camera=self.camera, bvh_seed=self.bvh_seed, bvh_rand_zero=int(self.bvh_rand_zero),
                       bvh_builder=int(self.bvh_builder))
        s._keep = (prims, mats, texs, stex, self._keep)  # keep borrowed memory alive with the struct
        return s
