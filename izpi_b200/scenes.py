"""Fixture scenes for the BASELINE.json configs (synthetic, seed-free or fixed-seed).

Restated scene *definitions* (data) from the reference:
  config 1  internal/scenes/scenes.go:119-155  CornellBox (Go object graph, RGB)
  config 4  cmd/izpi/examples/cornell_box_transparent_pyramid_spectral.pbtxt
  configs 2/3/5  synthetic meshes specified in SURVEY.md §8(d)
"""
from __future__ import annotations

import numpy as np

from . import scene as S
from .scene import SceneSpec, f32

U64 = np.uint64


# ---------------------------------------------------------------------------------------
# counter RNG for synthetic inputs (splitmix64 finaliser; numpy uint64 arithmetic wraps)
def _mix64(z):
    z = np.asarray(z, dtype=U64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def counter_uniform(seed: int, index, dim: int, attempt: int = 0):
    """U[0,1) with 53-bit mantissa from (seed, ray index, dimension, attempt)."""
    with np.errstate(over="ignore"):
        k = _mix64(U64(seed) + U64(0x9E3779B97F4A7C15) * (np.asarray(index, dtype=U64) + U64(1)))
        z = _mix64(k + U64(0xD1B54A32D192ED03) * U64(dim + 1) + U64(0x8CB92BA72F3D8DD7) * U64(attempt))
    return (z >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def random_rays(n: int, box_min, box_max, seed: int = 0x12D687, start: int = 0):
    """Incoherent ray batch of config 2: origins uniform in the box inflated 10 %, directions uniform
    on the sphere, components with |d| < 1e-12 re-drawn (no 0*Inf NaNs in the slab test)."""
    lo = np.asarray(box_min, dtype=np.float64)
    hi = np.asarray(box_max, dtype=np.float64)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo) * 1.1
    idx = np.arange(start, start + n, dtype=np.uint64)
    org = np.empty((n, 3))
    for d in range(3):
        org[:, d] = c[d] - h[d] + 2.0 * h[d] * counter_uniform(seed, idx, d)
    dirs = np.empty((n, 3))
    todo = np.arange(n)
    attempt = 0
    while len(todo):
        z = 1.0 - 2.0 * counter_uniform(seed, idx[todo], 3, attempt)
        phi = 2.0 * np.pi * counter_uniform(seed, idx[todo], 4, attempt)
        r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
        v = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        dirs[todo] = v
        todo = todo[(np.abs(v) < 1e-12).any(axis=1)]
        attempt += 1
    return org, dirs


# ---------------------------------------------------------------------------------------
def torus_mesh(n_around: int = 1000, n_tube: int = 500, centre=(50.0, 50.0, 50.0), major=30.0, minor=12.0,
               amp=3.0, scale=1.0):
    """Closed displaced torus, 2*n_around*n_tube triangles, vertices rounded to fp32 (proto Vec3).

    Order = row-major quads (i around, j tube), triangles (v00,v10,v11) then (v00,v11,v01);
    UV = (i/n_around, j/n_tube).  Returns verts (n,3,3) float64 and uvs (n,3,2)."""
    i = np.arange(n_around)
    j = np.arange(n_tube)
    u = i / n_around
    v = j / n_tube
    U, Vv = np.meshgrid(u, v, indexing="ij")
    th, ph = 2 * np.pi * U, 2 * np.pi * Vv
    disp = amp * (0.4 * np.sin(2 * np.pi * (3 * U + 5 * Vv) + 0.3) + 0.3 * np.sin(2 * np.pi * (7 * U - 4 * Vv) + 1.1)
                  + 0.2 * np.sin(2 * np.pi * (13 * U + 11 * Vv) + 2.3) + 0.1 * np.sin(2 * np.pi * (29 * U - 17 * Vv) + 0.7))
    nrm = np.stack([np.cos(ph) * np.cos(th), np.sin(ph), np.cos(ph) * np.sin(th)], axis=-1)
    base = np.stack([(major + minor * np.cos(ph)) * np.cos(th), minor * np.sin(ph), (major + minor * np.cos(ph)) * np.sin(th)], axis=-1)
    P = (base + disp[..., None] * nrm) * scale + np.asarray(centre)
    P = f32(P)  # (n_around, n_tube, 3)
    I, J = np.meshgrid(i, j, indexing="ij")
    I1, J1 = (I + 1) % n_around, (J + 1) % n_tube
    v00, v10, v11, v01 = P[I, J], P[I1, J], P[I1, J1], P[I, J1]
    verts = np.empty((n_around, n_tube, 2, 3, 3))
    verts[:, :, 0, 0], verts[:, :, 0, 1], verts[:, :, 0, 2] = v00, v10, v11
    verts[:, :, 1, 0], verts[:, :, 1, 1], verts[:, :, 1, 2] = v00, v11, v01
    u0, u1 = f32(I / n_around), f32((I + 1) / n_around)
    w0, w1 = f32(J / n_tube), f32((J + 1) / n_tube)
    uvs = np.empty((n_around, n_tube, 2, 3, 2))
    uvs[:, :, 0, 0] = np.stack([u0, w0], -1); uvs[:, :, 0, 1] = np.stack([u1, w0], -1); uvs[:, :, 0, 2] = np.stack([u1, w1], -1)
    uvs[:, :, 1, 0] = np.stack([u0, w0], -1); uvs[:, :, 1, 1] = np.stack([u1, w1], -1); uvs[:, :, 1, 2] = np.stack([u0, w1], -1)
    return verts.reshape(-1, 3, 3), uvs.reshape(-1, 3, 2)


def triangle_soup(n: int, seed: int = 7):
    """Secondary stress input: centres uniform in [0,100)^3, edges uniform in [-0.3,0.3)^3."""
    idx = np.arange(n, dtype=np.uint64)
    c = np.stack([100.0 * counter_uniform(seed, idx, d) for d in range(3)], axis=1)
    e1 = np.stack([0.6 * counter_uniform(seed, idx, 3 + d) - 0.3 for d in range(3)], axis=1)
    e2 = np.stack([0.6 * counter_uniform(seed, idx, 6 + d) - 0.3 for d in range(3)], axis=1)
    verts = f32(np.stack([c, c + e1, c + e2], axis=1))
    return verts


def closest_hit_scene(n_around=1000, n_tube=500, bvh_seed=12345):
    """Config 2: the 1M-triangle mesh in a BVH4 built with LCG(12345)."""
    verts, uvs = torus_mesh(n_around, n_tube)
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=bvh_seed)
    white = sc.lambertian(sc.constant_texture(f32([0.73, 0.73, 0.73])))
    sc.triangles(verts, white, uvs)
    sc.set_camera((50, 50, -120), (50, 50, 50), (0, 1, 0), 35, 1.0)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    return sc, lo, hi


# ---------------------------------------------------------------------------------------
def cornell_box(aspect: float = 1.0) -> SceneSpec:
    """Config 1: scenes.CornellBox (scenes.go:119-155).  Go literals, so exact fp64 (no fp32 rounding)."""
    sc = SceneSpec(world_kind=S.WORLD_SLICE)
    red = sc.lambertian(sc.constant_texture((0.65, 0.05, 0.05)))
    white = sc.lambertian(sc.constant_texture((0.73, 0.73, 0.73)))
    green = sc.lambertian(sc.constant_texture((0.12, 0.45, 0.15)))
    light = sc.diffuse_light(sc.constant_texture((15, 15, 15)))
    glass = sc.dielectric(1.5)
    sc.rect(S.PRIM_YZRECT, 0, 555, 0, 555, 555, green, flip=True)
    sc.rect(S.PRIM_YZRECT, 0, 555, 0, 555, 0, red)
    sc.rect(S.PRIM_XZRECT, 213, 343, 227, 332, 554, light, flip=True)
    sc.rect(S.PRIM_XZRECT, 0, 555, 0, 555, 555, white, flip=True)
    sc.rect(S.PRIM_XZRECT, 0, 555, 0, 555, 0, white)
    sc.rect(S.PRIM_XYRECT, 0, 555, 0, 555, 555, white, flip=True)
    sc.sphere((190, 90, 190), 90, glass)
    sc.box((0, 0, 0), (165, 330, 165), white, rotate_y=15.0, translate=(265, 0, 295))
    sc.set_camera((278.0, 278.0, -800.0), (278, 278, 0), (0, 1, 0), 40.0, aspect, 0.0, 10.0, 0.0, 1.0, 1.0)
    return sc


# config 4 data
_PYRAMID_TRIS = [
    ((100, 0, 100), (0, 0, 100), (100, 100, 100), "White"), ((100, 100, 100), (0, 0, 100), (0, 100, 100), "White"),
    ((0, 0, 0), (0, 0, 100), (100, 0, 100), "White"), ((0, 0, 0), (100, 0, 100), (100, 0, 0), "White"),
    ((0, 100, 0), (100, 100, 0), (100, 100, 100), "White"), ((0, 100, 100), (0, 100, 0), (100, 100, 100), "White"),
    ((33, 99, 33), (66, 99, 33), (66, 99, 66), "white_light"), ((33, 99, 33), (66, 99, 66), (33, 99, 66), "white_light"),
    ((0, 100, 100), (0, 0, 0), (0, 100, 0), "Green"), ((0, 100, 100), (0, 0, 100), (0, 0, 0), "Green"),
    ((100, 0, 0), (100, 100, 100), (100, 100, 0), "Red"), ((100, 0, 0), (100, 0, 100), (100, 100, 100), "Red"),
]
_PYRAMID_SPHERES = [(30, 15, 30), (50, 15, 30), (70, 15, 30), (40, 15, 50), (60, 15, 50), (50, 15, 70), (40, 28, 40),
                    (60, 28, 40), (50, 28, 60), (50, 42, 50)]
_GLASS_N = [1.52, 1.51, 1.51, 1.50, 1.50, 1.49, 1.49, 1.48, 1.48, 1.47, 1.47, 1.46, 1.46, 1.45, 1.45, 1.44, 1.44, 1.43,
            1.43, 1.42]
_GLASS_L = list(range(380, 741, 20)) + [750]
# lightsources.go:230-239  cie_f1_daylight_fluorescent, 75 samples @5nm from 380
CIE_F1 = [0.0350, 0.0380, 0.0430, 0.0500, 0.0590, 0.0710, 0.0870, 0.1090, 0.1390, 0.1800, 0.2360, 0.3130, 0.4190, 0.5660,
          0.7730, 1.0000, 0.9730, 0.7380, 0.5650, 0.4610, 0.3990, 0.3620, 0.3410, 0.3310, 0.3280, 0.3300, 0.3360, 0.3450,
          0.3570, 0.3710, 0.3880, 0.4070, 0.4290, 0.4530, 0.4800, 0.5080, 0.5390, 0.5720, 0.6080, 0.6460, 0.6870, 0.7310,
          0.7780, 0.8280, 0.8820, 0.9380, 0.9980, 1.0000, 0.9860, 0.9420, 0.8740, 0.7900, 0.6960, 0.6000, 0.5060, 0.4200,
          0.3430, 0.2780, 0.2240, 0.1800, 0.1450, 0.1170, 0.0950, 0.0770, 0.0630, 0.0520, 0.0430, 0.0360, 0.0300, 0.0250,
          0.0210, 0.0180, 0.0150, 0.0130, 0.0110]


def spectral_pyramid(aspect: float = 1.0, bvh_seed: int = 12345) -> SceneSpec:
    """Config 4: Cornell box + pyramid of 10 dispersive glass spheres, spectral (protobuf path:
    every scalar is fp32 on the wire)."""
    assert len(_GLASS_L) == 20 and len(CIE_F1) == 75
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=bvh_seed)
    mats = {
        "White": sc.spectral_lambertian(sc.spectral_neutral(f32(0.73))),
        "Green": sc.spectral_lambertian(sc.spectral_gaussian(f32(0.9), f32(540), f32(40))),
        "Red": sc.spectral_lambertian(sc.spectral_gaussian(f32(0.9), f32(640), f32(40))),
        "white_light": sc.spectral_diffuse_light(sc.spectral_tabulated(380.0 + 5.0 * np.arange(75), np.asarray(CIE_F1))),
        "Transparent": sc.spectral_dielectric(sc.spectral_tabulated(f32(_GLASS_L), f32(_GLASS_N)), sc.spectral_neutral(f32(0.01))),
    }
    verts = f32(np.array([[t[0], t[1], t[2]] for t in _PYRAMID_TRIS], dtype=np.float64))
    uvs = np.zeros((len(_PYRAMID_TRIS), 3, 2))
    uvs[0] = [(0, 0), (1, 0), (1, 1)]
    sc.triangles(verts, np.array([mats[t[3]] for t in _PYRAMID_TRIS], dtype=np.int32), uvs)
    for c in _PYRAMID_SPHERES:
        sc.sphere(f32(c), float(f32(10)), mats["Transparent"])
    sc.set_camera(f32((50, 50, -120)), f32((50, 50, 50)), f32((0, 1, 0)), f32(35), aspect, f32(0), f32(10), f32(0), f32(1), f32(1.0))
    return sc


def _value_noise(n, cells, seed):
    """Band-limited value noise in [0,1): bilinear interpolation of a cells x cells lattice (periodic)."""
    idx = np.arange(cells * cells, dtype=np.uint64)
    lat = counter_uniform(seed, idx, 0).reshape(cells, cells)
    t = np.arange(n) * (cells / n)
    i0 = np.floor(t).astype(int) % cells
    i1 = (i0 + 1) % cells
    f = t - np.floor(t)
    f = f * f * (3 - 2 * f)
    rows = lat[i0][:, i0] * (1 - f)[None, :] + lat[i0][:, i1] * f[None, :]
    rows1 = lat[i1][:, i0] * (1 - f)[None, :] + lat[i1][:, i1] * f[None, :]
    return rows * (1 - f)[:, None] + rows1 * f[:, None]


def pbr_textures(size: int = 2048):
    """Config 3's four synthetic fp64 RGBA textures: albedo, roughness, metalness, normal (OpenGL)."""
    a = np.stack([_value_noise(size, 16, 11 + c) * 0.7 + 0.2 for c in range(3)], axis=-1)
    rough = _value_noise(size, 8, 21)
    metal = (_value_noise(size, 4, 31) > 0.6).astype(np.float64) * 0.9
    height = _value_noise(size, 32, 41)
    gx = (np.roll(height, -1, axis=1) - np.roll(height, 1, axis=1)) * (size / 64.0)
    gy = (np.roll(height, -1, axis=0) - np.roll(height, 1, axis=0)) * (size / 64.0)
    n = np.stack([-gx, -gy, np.ones_like(gx)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)

    def rgba(rgb):
        out = np.ones((size, size, 4))
        out[..., :3] = rgb
        return out

    return rgba(a), rgba(rough[..., None].repeat(3, -1)), rgba(metal[..., None].repeat(3, -1)), rgba(0.5 * n + 0.5)


def sky_texture(width: int = 4096, height: int = 2048):
    """Synthetic HDR environment (config 5): vertical gradient + a sun disc, fp64 RGBA equirectangular."""
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    grad = np.stack([0.35 + 0.4 * v, 0.45 + 0.45 * v, 0.6 + 0.6 * v], axis=-1)  # brighter toward the top rows
    img = np.ones((height, width, 4))
    img[..., :3] = grad[:, None, :]
    du, dv = (u[None, :] - 0.3), (v[:, None] - 0.75)
    sun = np.exp(-((du * 2.0) ** 2 + dv ** 2) / (2 * 0.012 ** 2))
    img[..., :3] += sun[..., None] * np.array([60.0, 52.0, 40.0])
    return img


def ibl_displaced_mesh(aspect: float = 3840 / 2160, n_around: int = 3162, n_tube: int = 1581, env_size=(4096, 2048),
                       bvh_seed: int = 12345) -> SceneSpec:
    """Config 5: image-based-lit scene with a ~10M-triangle displaced mesh (2*3162*1581 = 9 998 244 triangles).

    The sky is hitable.NewSkyDome's construction (sphere.go:39-48): FlipNormals(Sphere) with a DiffuseLight whose
    emit texture is an image.  Like scenes.Environment (scenes.go:233-266) the lit objects are specular (Metal,
    Dielectric): the dome is in scene.Lights and Sphere.PDFValue evaluated from INSIDE a sphere is NaN
    (sqrt(1 - r^2/d^2), sphere.go:131), so a diffuse surface under a sky dome is DeNAN'ed to black by the reference
    itself -- reproduced here, not a defect of this backend.
    The mesh is the analytically displaced torus of config 2 at higher tessellation: the reference's
    displacement tessellator (internal/displacement) is a 'next' row (DESIGN.md §1 f) and not restated yet."""
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=bvh_seed)
    metal = sc.metal(f32((0.92, 0.86, 0.78)), f32(0.03))
    glass = sc.dielectric(f32(1.5))
    sky = sc.diffuse_light(sc.image_texture(sky_texture(*env_size)))
    verts, uvs = torus_mesh(n_around, n_tube, centre=(0.0, 0.0, 0.0), major=30.0, minor=12.0, amp=3.0)
    sc.triangles(verts, metal, uvs)
    sc.sphere(f32((0.0, 22.0, 0.0)), float(f32(9.0)), glass)
    sc.sphere(f32((0.0, 0.0, 0.0)), float(f32(500.0)), sky)
    sc.prims["wrap"][-1] = S.WRAP_FLIP
    sc.set_camera(f32((70, 45, 95)), f32((0, 2, 0)), f32((0, 1, 0)), f32(38), aspect, f32(0), f32(10), f32(0), f32(1), f32(1.0))
    return sc


def height_map(width: int = 4096, height: int = 2048):
    """Smooth synthetic displacement map (fp64 RGBA, height in the blue channel as displacement.go:108-110 reads it).
    Band-limited, so adjacent texels differ by far less than the tessellator's adaptive threshold allows."""
    u = (np.arange(width) + 0.5) / width
    v = (np.arange(height) + 0.5) / height
    U, V = np.meshgrid(u, v)
    z = (0.5 + 0.25 * np.sin(2 * np.pi * 6 * U) * np.cos(2 * np.pi * 5 * V) + 0.15 * np.sin(2 * np.pi * 17 * U + 1) * np.sin(2 * np.pi * 13 * V)
         + 0.1 * np.sin(2 * np.pi * 37 * U) * np.cos(2 * np.pi * 29 * V + 2))
    px = np.ones((height, width, 4))
    px[..., 2] = z
    return px


def ibl_tessellated_mesh(ctx, aspect: float = 3840 / 2160, n_around: int = 160, n_tube: int = 80, map_size=(5632, 2816),
                         range_fraction: float = 0.9, env_size=(4096, 2048), bvh_seed: int = 12345, bvh_builder: int = S.BVH_REFERENCE):
    """Config 5 as BASELINE words it: image-based-lit scene with a DISPLACEMENT-TESSELLATED mesh.  A coarse torus
    (2*n_around*n_tube triangles, proto-style fp32 vertices) goes through displacement.ApplyDisplacementMap
    (run on the device, izpi_displace) with a smooth height map, one call per base triangle as transport.go:633-646
    does; the displacement range is `range_fraction` of the largest range for which the reference's adaptive loop
    terminates on this map (threshold 2.0 / largest adjacent-texel step).  Returns (SceneSpec, n_triangles).
    The scene is scaled so that the fixed world-space threshold of displacement.go:183 is meaningful
    (torus radii 600 / 240, displacement range a fraction of the tube radius)."""
    px = height_map(*map_size)
    step = max(np.abs(np.diff(px[..., 2], axis=0)).max(), np.abs(np.diff(px[..., 2], axis=1)).max())
    rng = range_fraction * 2.0 / step
    verts, uvs = torus_mesh(n_around, n_tube, centre=(0.0, 0.0, 0.0), major=600.0, minor=240.0, amp=0.0)
    base = np.concatenate([verts.reshape(-1, 9), uvs.reshape(-1, 6)], axis=1)
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=bvh_seed, bvh_builder=bvh_builder)
    metal = sc.metal(f32((0.92, 0.86, 0.78)), f32(0.03))
    glass = sc.dielectric(f32(1.5))
    sky = sc.diffuse_light(sc.image_texture(sky_texture(*env_size)))
    tris, mats = ctx.apply_displacement(base, np.full(len(base), metal, dtype=np.int32), px, -rng / 2, rng / 2, per_triangle=True)
    sc.triangles(tris[:, :9].reshape(-1, 3, 3), mats, tris[:, 9:].reshape(-1, 3, 2))
    sc.sphere(f32((0.0, 480.0, 0.0)), float(f32(180.0)), glass)
    sc.sphere(f32((0.0, 0.0, 0.0)), float(f32(10000.0)), sky)
    sc.prims["wrap"][-1] = S.WRAP_FLIP
    sc.set_camera(f32((1400, 900, 1900)), f32((0, 40, 0)), f32((0, 1, 0)), f32(38), aspect, f32(0), f32(10), f32(0), f32(1), f32(1.0))
    return sc, len(tris)


_CORNELL_RGB = {"White": (0.73, 0.73, 0.73), "Green": (0, 0.73, 0), "Red": (0.73, 0, 0)}


def cornell_pbr_mesh(aspect: float = 1.0, n_around: int = 1000, n_tube: int = 500, tex_size: int = 2048,
                     bvh_seed: int = 12345, n_pbr_materials: int = 1) -> SceneSpec:
    """Config 3: RGB Cornell walls (12 triangles above, materials of scenes.CornellBoxRGB
    scenes.go:934-1378) + the config-2 torus scaled x0.55 into the box, PBR with image textures.
    n_pbr_materials > 1 (not a BASELINE config) splits the mesh into bands around the torus, each with its own PBR material and
    its own four textures: the multi-material case that sorting hits by material ID is for."""
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=bvh_seed)
    mats = {k: sc.lambertian(sc.constant_texture(f32(v))) for k, v in _CORNELL_RGB.items()}
    mats["white_light"] = sc.diffuse_light(sc.constant_texture(f32((15, 15, 15))))
    alb, rough, metal, nrm = pbr_textures(tex_size)
    pbrs = []
    for k in range(max(1, n_pbr_materials)):
        a = alb if k == 0 else np.ascontiguousarray(np.roll(alb, 97 * k, axis=1)[..., [(0 + k) % 3, (1 + k) % 3, (2 + k) % 3, 3]])
        r, m, n_ = (rough, metal, nrm) if k == 0 else (np.roll(rough, 31 * k, axis=0).copy(), np.roll(metal, 53 * k, axis=1).copy(), nrm.copy())
        pbrs.append(sc.pbr(sc.image_texture(a), sc.image_texture(n_), sc.image_texture(r), sc.image_texture(m)))
    verts = f32(np.array([[t[0], t[1], t[2]] for t in _PYRAMID_TRIS], dtype=np.float64))
    uvs = np.zeros((len(_PYRAMID_TRIS), 3, 2))
    uvs[0] = [(0, 0), (1, 0), (1, 1)]
    sc.triangles(verts, np.array([mats[t[3]] for t in _PYRAMID_TRIS], dtype=np.int32), uvs)
    tv, tuv = torus_mesh(n_around, n_tube, centre=(50.0, 40.0, 50.0), major=30.0, minor=12.0, amp=3.0, scale=0.55)
    if len(pbrs) == 1:
        sc.triangles(tv, pbrs[0], tuv)
    else:  # bands of 8 quads around the torus, materials in rotation
        quad_row = np.arange(len(tv)) // (2 * n_tube)
        sc.triangles(tv, np.asarray(pbrs, dtype=np.int32)[(quad_row // 8) % len(pbrs)], tuv)
    sc.set_camera(f32((50, 50, -140)), f32((50, 50, 0)), f32((0, 1, 0)), f32(40), aspect, f32(0), f32(10), f32(0), f32(1), f32(1.0))
    return sc
