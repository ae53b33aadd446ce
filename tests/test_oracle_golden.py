"""The oracle against the reference's OWN golden vectors.

Sources (relative to the reference repo):
  internal/hitable/bvh4_simd_test.go:54-158   7 box-test cases with expected masks
  internal/hitable/bvh4_simd_test.go:200-268  1000 deterministic cases, SIMD == scalar
  internal/hitable/bvh4_test.go:160-278       +Inf inverse-direction cases
  internal/hitable/bvh4_test.go:337-415       conservativeFloat32Min/Max
  internal/hitable/triangle_test.go:15-134    Triangle ctor fields + 4 Hit records
  internal/hitable/bvh4_test.go:13-83,454-517 BVH4 structure invariants (validate())
  internal/material/dielectric_test.go:28-38  Beer-Lambert exp(-0.5)
"""
import math

import numpy as np
import pytest

from izpi_b200 import scene as S
from izpi_b200.scene import SceneSpec

INF = float("inf")
MAXF32 = float(np.finfo(np.float32).max)

BOX_CASES = [
    ("All hits", (0, 0, 0), (1, 1, 1), [1, 2, 3, 4], [1, 2, 3, 4], [1, 2, 3, 4], [2, 3, 4, 5], [2, 3, 4, 5], [2, 3, 4, 5], 100, 0b1111),
    ("No hits - ray pointing away", (0, 0, 0), (-1, -1, -1), [1, 2, 3, 4], [1, 2, 3, 4], [1, 2, 3, 4], [2, 3, 4, 5], [2, 3, 4, 5], [2, 3, 4, 5], 100, 0b0000),
    ("tMax cutoff", (0, 0, 0), (1, 1, 1), [1, 2, 10, 20], [1, 2, 10, 20], [1, 2, 10, 20], [2, 3, 11, 21], [2, 3, 11, 21], [2, 3, 11, 21], 5, 0b0011),
    ("origin inside AABB", (1.5, 1.5, 1.5), (1, 1, 1), [1, 5, 5, 5], [1, 5, 5, 5], [1, 5, 5, 5], [2, 6, 6, 6], [2, 6, 6, 6], [2, 6, 6, 6], 100, 0b1111),
    ("Negative direction", (5, 5, 5), (-1, -1, -1), [1, 2, 3, 6], [1, 2, 3, 6], [1, 2, 3, 6], [2, 3, 4, 7], [2, 3, 4, 7], [2, 3, 4, 7], 10, 0b0111),
    ("Mixed directions", (0, 0, 0), (1, 1, -1), [1, 1, 1, 1], [1, 1, 1, 1], [-2, 1, -2, 1], [2, 2, 2, 2], [2, 2, 2, 2], [-1, 2, -1, 2], 100, 0b0101),
    ("Infinite-ish direction", (0, 5, 5), (1, MAXF32, MAXF32), [1, 2, 3, 4], [4, 4, 6, 6], [4, 4, 6, 6], [2, 3, 4, 5], [6, 6, 7, 7], [6, 6, 7, 7], 100, 0b0011),
]


@pytest.mark.parametrize("case", BOX_CASES, ids=[c[0] for c in BOX_CASES])
@pytest.mark.parametrize("flavour", [0, 1], ids=["sse", "scalar"])
def test_box_golden(oracle_mod, case, flavour):
    _, org, inv, mnx, mny, mnz, mxx, mxy, mxz, tmax, want = case
    got = oracle_mod.ray_aabb4(flavour, org, inv, [mnx, mny, mnz, mxx, mxy, mxz], tmax)
    assert got == want


def box_random_case(i):
    """bvh4_simd_test.go:200-236 generator (all arithmetic in float32)."""
    f = np.float32
    org = [f(i % 10 - 5), f((i + 1) % 10 - 5), f((i + 2) % 10 - 5)]
    d = [f(f((i % 7) - 3) + f(0.1)), f(f(((i + 1) % 7) - 3) + f(0.1)), f(f(((i + 2) % 7) - 3) + f(0.1))]
    inv = [f(1.0) / x for x in d]
    b = np.zeros((6, 4), dtype=np.float32)
    for j in range(4):
        base = [f((i + j) % 20 - 10), f((i + j + 1) % 20 - 10), f((i + j + 2) % 20 - 10)]
        for a in range(3):
            b[a, j] = base[a]
            b[3 + a, j] = base[a] + f(j + 1)
    return org, inv, b, f(50 + i % 50)


def test_box_1000_deterministic(oracle_mod):
    masks = []
    for i in range(1000):
        org, inv, b, tmax = box_random_case(i)
        a = oracle_mod.ray_aabb4(0, org, inv, b, tmax)
        s = oracle_mod.ray_aabb4(1, org, inv, b, tmax)
        assert a == s, i
        masks.append(a)
    assert len(set(masks)) > 1  # the generator exercises more than one outcome


@pytest.mark.parametrize("flavour", [0, 1])
def test_box_inf_invdir(oracle_mod, flavour):
    org, inv = (0, 0, 0), (INF, INF, -1.0)
    # bvh4_test.go:160-204: only box 0 contains the ray
    b = [[-1, 10, -1, 10], [-1, -1, 10, 10], [-10] * 4, [1, 12, 1, 12], [1, 1, 12, 12], [-2] * 4]
    m = oracle_mod.ray_aabb4(flavour, org, inv, b, 100.0)
    assert m & 1 and not m & 2 and not m & 4 and not m & 8
    # :207-241 all four hit
    b = [[-2] * 4, [-2] * 4, [-10] * 4, [2] * 4, [2] * 4, [-1] * 4]
    assert oracle_mod.ray_aabb4(flavour, org, inv, b, 100.0) == 0b1111
    # :244-278 all behind
    b = [[-2] * 4, [-2] * 4, [1] * 4, [2] * 4, [2] * 4, [10] * 4]
    assert oracle_mod.ray_aabb4(flavour, org, inv, b, 100.0) == 0


@pytest.mark.parametrize("v", [1234567890.123456789, -1234567890.123456789, 0.00000012345678901234, math.pi * 1000000])
def test_conservative_float32(oracle_mod, v):
    L = oracle_mod.lib()
    lo, hi = L.oracle_conservative_float32_min(v), L.oracle_conservative_float32_max(v)
    assert float(lo) <= v <= float(hi)
    # at most one ulp apart, and exact when v is representable
    assert np.nextafter(np.float32(lo), np.float32(np.inf)) >= np.float32(hi)
    assert L.oracle_conservative_float32_min(0.5) == 0.5 == L.oracle_conservative_float32_max(0.5)


def _one_tri(v0, v1, v2, uv=None, bvh=False):
    sc = SceneSpec(world_kind=S.WORLD_BVH4 if bvh else S.WORLD_SLICE, bvh_rand_zero=True)
    m = sc.dielectric(1.0)
    sc.triangles(np.array([[v0, v1, v2]], dtype=np.float64), m, uv)
    return sc


def test_triangle_ctor_golden(oracle_mod):
    sc = _one_tri((0, 0, 0), (-1, 0, 0), (0, 1, 0), uv=np.array([[(0, 0), (1, 0), (0, 1)]], dtype=np.float64))
    f = oracle_mod.OracleScene(sc).triangle_fields(0)
    np.testing.assert_array_equal(f["edge1"], [-1, 0, 0])
    np.testing.assert_array_equal(f["edge2"], [0, 1, 0])
    np.testing.assert_array_equal(f["normal"], [0, 0, -1])
    np.testing.assert_array_equal(f["tangent"], [-1, 0, 0])
    np.testing.assert_array_equal(f["bitangent"], [0, 1, 0])
    assert f["area"] == 0.5
    np.testing.assert_array_equal(f["bbmin"], [-1.0001, -0.0001, -0.0001])
    np.testing.assert_array_equal(f["bbmax"], [0.0001, 1.0001, 0.0001])


TRI_HITS = [
    ("parallel", ((1, 0, -1), (1, 1, -1), (0, 0, -1)), (0, -1, 0), (0, 1, 0), None),
    ("perpendicular hit", ((.5, -.5, -10), (0, .5, -10), (-.5, -.5, -10)), (0, 0, 1), (0, 0, -1),
     dict(t=11.0, u=0.0, v=0.0, p=(0, 0, -10), normal=(0, 0, 1))),
    ("perpendicular miss", ((.5, -.5, -10), (0, .5, -10), (-.5, -.5, -10)), (-1, 0, 1), (-1, 0, -1), None),
    ("angled hit", ((.5, -.5, -20), (0, .5, -10), (-.5, -.5, -10)), (0, 0, 1), (0, 0, -1),
     dict(t=13.5, u=0.0, v=0.0, p=(0, 0, -12.5), normal=(0.8908708063747479, -0.44543540318737396, 0.0890870806374748))),
]


@pytest.mark.parametrize("case", TRI_HITS, ids=[c[0] for c in TRI_HITS])
@pytest.mark.parametrize("bvh", [False, True], ids=["slice", "bvh4"])
def test_triangle_hit_golden(oracle_mod, case, bvh):
    _, tri, org, d, want = case
    got = oracle_mod.OracleScene(_one_tri(*tri, bvh=bvh)).hit(org, d, 0.0, np.finfo(np.float64).max)
    if want is None:
        assert got is None
        return
    assert got is not None and got["prim"] == 0
    assert got["t"] == want["t"] and got["u"] == want["u"] and got["v"] == want["v"]
    np.testing.assert_array_equal(got["p"], want["p"])
    np.testing.assert_array_equal(got["normal"], want["normal"])


def _validate(nodes, perm, n_prims):
    """BVH4.validate (bvh4.go:399-466) + every primitive referenced exactly once."""
    seen = np.zeros(n_prims, dtype=np.int64)
    todo, visited = [0], set()
    while todo:
        i = todo.pop()
        assert 0 <= i < len(nodes) and i not in visited
        visited.add(i)
        n = nodes[i]
        for s in range(4):
            c, k = int(n["child_index"][s]), int(n["primitive_count"][s])
            if c == -1:
                assert n["min_x"][s] == np.float32(MAXF32)
                continue
            if k > 0:
                assert 0 <= c and c + k <= n_prims and k <= 4
                seen[perm[c:c + k]] += 1
            else:
                todo.append(c)
    assert len(visited) == len(nodes)
    assert (seen == 1).all()


def _sphere_scene(centres, r=1.0):
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_rand_zero=True)
    m = sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5)))
    for c in centres:
        sc.sphere(c, r, m)
    return sc


@pytest.mark.parametrize("centres,want_nodes", [
    ([(0, 0, 0)], 1),
    ([(0, 0, 0), (1, 0, 0)], 1),
    ([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)], 3),
    ([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1), (2, 0, 0), (0, 2, 0), (2, 2, 0), (0, 0, 2), (2, 2, 2)], 5),
])
def test_bvh4_small_structure(oracle_mod, centres, want_nodes):
    """bvh4_test.go:13-83 scenes with randomFunc == 0; node counts follow from the build rules
    (<=4 -> one leaf node; 5 -> root + leaves of 2 and 3; 10 -> root + 2+3+2+3)."""
    nodes, perm = oracle_mod.OracleScene(_sphere_scene(centres)).bvh()
    assert len(nodes) == want_nodes
    _validate(nodes, perm, len(centres))


def test_bvh4_10000_sphere_grid(oracle_mod):
    """bvh4_test.go:454-496: 10 000 spheres on a grid, validate() must pass; totals as debugStats."""
    g = np.stack(np.meshgrid(np.arange(100), np.arange(100), indexing="ij"), -1).reshape(-1, 2)
    sc = SceneSpec(world_kind=S.WORLD_BVH4, bvh_seed=12345)
    m = sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5)))
    a = sc._new(len(g))
    a["type"] = S.PRIM_SPHERE
    a["material"] = m
    a["p"][:, 0] = g[:, 0] * 3.0
    a["p"][:, 2] = g[:, 1] * 3.0
    a["p"][:, 3] = 1.0
    nodes, perm = oracle_mod.OracleScene(sc).bvh()
    _validate(nodes, perm, len(g))
    leaf = nodes["primitive_count"][:, 0] > 0
    assert nodes["primitive_count"].sum() == len(g)
    # leaves are their own nodes using slot 0 only (bvh4.go:737-760)
    assert (nodes["child_index"][leaf][:, 1:] == -1).all()
    assert (nodes["primitive_count"][~leaf] == 0).all()
    # root bounds contain everything (bvh4_test.go:281-310)
    root = nodes[0]
    assert root["min_x"].min() <= -1.0 and root["max_x"][root["child_index"] != -1].max() >= 99 * 3 + 1.0


def test_lcg_constants(oracle_mod):
    """fastrandom.go:7-11,41-47."""
    import ctypes as C
    st = C.c_uint64(12345)
    v = oracle_mod.lib().oracle_lcg_next(C.byref(st))
    want = (1664525 * 12345 + 1013904223) % 2**32
    assert st.value == want and v == want / 2**32


def test_beer_lambert_golden(oracle_mod):
    """dielectric_test.go:28-38: gaussian(0.5, 480, 60) absorption at 480 nm over 1 unit -> exp(-0.5)."""
    sc = SceneSpec(world_kind=S.WORLD_SLICE)
    t = sc.spectral_gaussian(0.5, 480.0, 60.0)
    sc.sphere((0, 0, 0), 1, sc.spectral_dielectric(sc.spectral_neutral(1.5), t))
    alpha = oracle_mod.OracleScene(sc).spectral_texture_value(t, 480.0)
    assert alpha == 0.5
    assert abs(math.exp(-alpha * 1.0) - 0.6065) < 0.001


def test_tiles(oracle_mod):
    """common/tiles.go:3-24."""
    assert oracle_mod.tiles(400, 400) == (25, 25)
    assert oracle_mod.tiles(1024, 1024) == (32, 32)
    assert oracle_mod.tiles(3840, 2160) == (32, 24)
    assert oracle_mod.tiles(500, 500) == (25, 25)


def test_sample_wavelength_properties(oracle_mod):
    import ctypes as C
    L = oracle_mod.lib()
    lam, pdf = C.c_double(), C.c_double()
    L.oracle_sample_wavelength(0.0, C.byref(lam), C.byref(pdf))
    assert lam.value == 380.0 and pdf.value == 0.0          # spectral.go:206-212 (i == 0 branch)
    L.oracle_sample_wavelength(0.9999999, C.byref(lam), C.byref(pdf))
    assert lam.value == 750.0                                 # cieYIntegral 21.3768 > sum(cieY): fallback :217-223
    prev = 0.0
    for r in np.linspace(0.001, 0.99, 200):
        L.oracle_sample_wavelength(float(r), C.byref(lam), C.byref(pdf))
        assert 380.0 <= lam.value <= 750.0 and lam.value >= prev and pdf.value > 0
        prev = lam.value
    xyz = np.zeros(3)
    L.oracle_cie_values(555.0, xyz.ctypes.data)
    np.testing.assert_array_equal(xyz, [0.5121, 1.0, 0.0057])
    L.oracle_cie_values(557.5, xyz.ctypes.data)
    np.testing.assert_allclose(xyz, [(0.5121 + 0.5945) / 2, (1.0 + 0.995) / 2, (0.0057 + 0.0039) / 2], rtol=1e-15)


def test_dielectric_lcg_replays(oracle_mod):
    """material/dielectric_test.go:47-216 replayed: mock world, LCG(12345), both reflected (attenuation 1) and
    transmitted (Beer-Lambert) samples appear within 1000 scatter calls; path length |exit - hit| is in bounds."""
    import ctypes as C
    L = oracle_mod.lib()
    L.oracle_probe_dielectric_test.argtypes = [C.c_int, C.c_void_p]
    out = np.zeros(8)
    L.oracle_probe_dielectric_test(2, out.ctypes.data)   # TestPathLengthCalculation
    assert 0.1 <= out[6] <= 10.0 and out[6] == math.sqrt(0.0 + 0.0 + 1.0)  # |(0.5,0.5,1.5)-(0.5,0.5,0.5)|
    L.oracle_probe_dielectric_test(0, out.ctypes.data)   # TestColoredGlassScattering
    assert out[1] == 1 and out[2] == 1 and out[0] <= 1000
    d = math.sqrt(0.25 + 0.25 + 0.25)                     # hit (0,0,1) -> mock exit (0.5,0.5,1.5)
    np.testing.assert_allclose(out[3:6], [math.exp(-0.1 * d), math.exp(-0.2 * d), math.exp(-0.3 * d)], rtol=1e-15)
    L.oracle_probe_dielectric_test(1, out.ctypes.data)   # TestSpectralColoredGlassScattering
    assert out[1] == 1 and out[2] == 1 and out[0] <= 1000
    assert out[3] == math.exp(-0.5 * d)                   # gaussian(0.5, 480, 60) at 480 nm is 0.5


def test_tbn_normal_map_golden(oracle_mod):
    """mat3/mat3_test.go:25-30 through Triangle.Hit (triangle.go:250-264): a normal-map texel (0.5, 0.5, 1.0) maps to
    tangent-space (0,0,1); for the XY-plane triangle with TBN = ((-1,0,0), (0,1,0), (0,0,-1)) the result is (0,0,-1)."""
    sc = SceneSpec(world_kind=S.WORLD_SLICE)
    nm = sc.image_texture(np.array([[[0.5, 0.5, 1.0, 1.0]]]))
    pbr = sc.pbr(sc.constant_texture((0.5, 0.5, 0.5)), normal=nm)
    sc.triangles(np.array([[(0, 0, 0), (-1, 0, 0), (0, 1, 0)]], dtype=np.float64), pbr, np.array([[(0, 0), (1, 0), (0, 1)]], dtype=np.float64))
    h = oracle_mod.OracleScene(sc).hit((-0.25, 0.25, 5.0), (0, 0, -1), 0.0)
    assert h is not None and h["t"] == 5.0
    np.testing.assert_array_equal(h["normal"], [0, 0, -1])
    np.testing.assert_allclose([h["u"], h["v"]], [0.25, 0.25], rtol=1e-15)


@pytest.mark.parametrize("make,lo,hi", [
    (lambda sc: sc.spectral_gaussian(1.0, 550, 40), 1.0, 1.0),             # peak at the centre wavelength
    (lambda sc: sc.spectral_neutral(0.73), 0.73, 0.73),
    (lambda sc: sc.spectral_tabulated([380, 500, 600, 750], [0.1, 0.5, 0.8, 0.3]), 0.65, 0.65),  # lerp(500->600) at 550
])
def test_spectral_texture_values_at_550(oracle_mod, make, lo, hi):
    """transport/transport_test.go:193-274 evaluates each spectral texture kind at 550 nm (>= 0); the exact values
    follow from spectral_constant.go:65-106."""
    sc = SceneSpec(world_kind=S.WORLD_SLICE)
    t = make(sc)
    sc.sphere((0, 0, 0), 1, sc.spectral_lambertian(t))
    v = oracle_mod.OracleScene(sc).spectral_texture_value(t, 550.0)
    assert v >= 0 and lo - 1e-12 <= v <= hi + 1e-12
