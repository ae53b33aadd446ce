"""Closest-hit parity on the B200: the CUDA path (through the C ABI) against the oracle on the same
seeded inputs.  Bar (BASELINE.json): primitive IDs bit-exact, t within 1e-12 relative (we assert
bit-equal t as well, since both sides run the same fp64 operations without FMA)."""
import numpy as np
import pytest

from izpi_b200 import cuda, scenes
from izpi_b200 import scene as S

pytestmark = pytest.mark.gpu
DMAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def ctx():
    from izpi_b200.build import build
    build()
    c = cuda.Context(0)
    yield c
    c.close()


def _check(ctx, oracle_mod, spec, org, d, tmin=0.001, tmax=DMAX, expect_hits=True, compare_prim_counts=True):
    hs = cuda.HostScene(spec)
    ctx.upload(hs)
    osn = oracle_mod.OracleScene(spec)
    gi, gt, gst = ctx.trace_closest(org, d, tmin, tmax, stats=True)
    oi, ot, ost = osn.trace(org, d, tmin, tmax, stats=True)
    np.testing.assert_array_equal(gi, oi)
    hit = oi >= 0
    if expect_hits:
        assert hit.any()
    assert np.all(np.abs(gt[hit] - ot[hit]) <= 1e-12 * np.abs(ot[hit]))
    assert gt.tobytes() == ot.tobytes()
    # same traversal: identical node-visit and primitive-test counts (the algorithmic-bytes basis)
    assert gst["nodes"] == ost["nodes"]
    if compare_prim_counts:  # (the oracle counts a Box as its six rects, the device as one record)
        assert gst["prims"] == ost["tris"] + ost["spheres"] + ost["others"]
    # the non-counting kernel gives the same answers
    gi2, gt2 = ctx.trace_closest(org, d, tmin, tmax)
    assert gi2.tobytes() == gi.tobytes() and gt2.tobytes() == gt.tobytes()
    return gi, gt


def test_box_golden_on_device(ctx, oracle_mod):
    from test_oracle_golden import BOX_CASES, box_random_case
    org, inv, b, tm, want = [], [], [], [], []
    for _, o, i, mnx, mny, mnz, mxx, mxy, mxz, t, w in BOX_CASES:
        org.append(o); inv.append(i); b.append([mnx, mny, mnz, mxx, mxy, mxz]); tm.append(t); want.append(w)
    inf = float("inf")
    for bounds, w in [([[-1, 10, -1, 10], [-1, -1, 10, 10], [-10] * 4, [1, 12, 1, 12], [1, 1, 12, 12], [-2] * 4], 0b0001),
                      ([[-2] * 4, [-2] * 4, [-10] * 4, [2] * 4, [2] * 4, [-1] * 4], 0b1111),
                      ([[-2] * 4, [-2] * 4, [1] * 4, [2] * 4, [2] * 4, [10] * 4], 0)]:
        org.append((0, 0, 0)); inv.append((inf, inf, -1.0)); b.append(bounds); tm.append(100.0); want.append(w)
    for i in range(1000):
        o, iv, bb, t = box_random_case(i)
        org.append(o); inv.append(iv); b.append(bb); tm.append(t)
        want.append(oracle_mod.ray_aabb4(0, o, iv, bb, t))
    got = ctx.debug_ray_aabb4(org, inv, np.array(b, dtype=np.float32), tm)
    np.testing.assert_array_equal(got, np.array(want, dtype=np.uint8))


def test_triangle_golden_on_device(ctx, oracle_mod):
    """triangle_test.go:69-134 on the device, the whole hitrecord.HitRecord: t, u, v, p and the exact normal
    (0.8908708063747479, -0.44543540318737396, 0.0890870806374748), through the slice world and the BVH4 world."""
    from test_oracle_golden import TRI_HITS, _one_tri
    for bvh in (False, True):
        for _, tri, org, d, want in TRI_HITS:
            hs = cuda.HostScene(_one_tri(*tri, bvh=bvh))
            ctx.upload(hs)
            ids, t = ctx.trace_closest([org], [d], 0.0, DMAX)
            hid, rec = ctx.debug_hit([org], [d], 0.0, DMAX)
            if want is None:
                assert ids[0] == -1 and hid[0] == -1
            else:
                assert ids[0] == 0 and t[0] == want["t"]
                assert hid[0] == 0 and rec[0, 0] == want["t"] and rec[0, 1] == want["u"] and rec[0, 2] == want["v"]
                np.testing.assert_array_equal(rec[0, 3:6], want["p"])
                np.testing.assert_array_equal(rec[0, 6:9], want["normal"])


def test_full_hit_records_match_oracle(ctx, oracle_mod):
    """Full hit records (t, u, v, p, normal) of every primitive kind and wrapper, device vs oracle, bit for bit except where the
    record passes through atan2 / asin (sphere UVs: the device's libm is within 2 ulp of glibc's, not identical)."""
    sc = scenes.cornell_box(1.0)
    sc.world_kind = S.WORLD_BVH4
    sc.sphere((400, 60, 150), 60, sc.metal((0.8, 0.85, 0.88), 0.1))
    verts, uvs = scenes.torus_mesh(30, 20, centre=(278.0, 278.0, 278.0), major=120.0, minor=40.0, amp=10.0)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    ctx.upload(cuda.HostScene(sc))
    osn = oracle_mod.OracleScene(sc)
    org, d = scenes.random_rays(4000, (0, 0, 0), (555, 555, 555))
    ids, rec = ctx.debug_hit(org, d, 0.001, DMAX)
    n_sphere = 0
    for i in range(len(org)):
        got = osn.hit(org[i], d[i], 0.001, DMAX)
        if got is None:
            assert ids[i] == -1
            continue
        assert ids[i] == got["prim"]
        assert rec[i, 0] == got["t"]
        np.testing.assert_array_equal(rec[i, 3:6], got["p"])
        np.testing.assert_array_equal(rec[i, 6:9], got["normal"])
        if sc.prims["type"][got["prim"]] == S.PRIM_SPHERE:
            n_sphere += 1
            np.testing.assert_allclose(rec[i, 1:3], [got["u"], got["v"]], rtol=0, atol=4e-16)
        else:
            assert rec[i, 1] == got["u"] and rec[i, 2] == got["v"]
    assert n_sphere > 20 and (ids >= 0).mean() > 0.5


@pytest.mark.parametrize("shape,nrays", [((40, 25), 1 << 14), ((300, 200), 1 << 17)])
def test_torus_parity(ctx, oracle_mod, shape, nrays):
    verts, uvs = scenes.torus_mesh(*shape)
    sc = S.SceneSpec(bvh_seed=12345)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org, d = scenes.random_rays(nrays, lo, hi)
    ids, _ = _check(ctx, oracle_mod, sc, org, d)
    assert 0.2 < (ids >= 0).mean() < 0.9


def test_ragged_and_empty_batches(ctx, oracle_mod):
    verts, uvs = scenes.torus_mesh(40, 25)
    sc = S.SceneSpec(bvh_seed=1)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    hs = cuda.HostScene(sc)
    ctx.upload(hs)
    ids, t = ctx.trace_closest(np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(ids) == 0 and len(t) == 0
    for n in (1, 31, 33, 1000):  # not multiples of the 32-ray packet
        org, d = scenes.random_rays(n, lo, hi, seed=n)
        _check(ctx, oracle_mod, sc, org, d, expect_hits=False)


def test_tmax_window_and_axis_aligned_rays(ctx, oracle_mod):
    """Finite tMax windows, and rays with zero direction components (inv = +-Inf in the slab test)."""
    verts, uvs = scenes.torus_mesh(60, 40)
    sc = S.SceneSpec(bvh_seed=2)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org, d = scenes.random_rays(1 << 13, lo, hi, seed=5)
    _check(ctx, oracle_mod, sc, org, d, tmin=0.001, tmax=15.0)
    _check(ctx, oracle_mod, sc, org, d, tmin=5.0, tmax=40.0)
    axis = np.zeros_like(d)
    axis[np.arange(len(d)), np.arange(len(d)) % 3] = np.where(np.arange(len(d)) % 2, 1.0, -1.0)
    org2 = org + 0.123456789  # off the lattice of box planes: no 0*Inf NaN
    _check(ctx, oracle_mod, sc, org2, axis)


def test_soup_deep_overlap(ctx, oracle_mod):
    sc = S.SceneSpec(bvh_seed=3)
    sc.triangles(scenes.triangle_soup(50000), sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))))
    org, d = scenes.random_rays(1 << 15, (0, 0, 0), (100, 100, 100), seed=11)
    _check(ctx, oracle_mod, sc, org, d)


def test_mixed_triangles_and_spheres_in_bvh(ctx, oracle_mod):
    """Config 4's geometry: 12 triangles + 10 spheres in one BVH4 (strict vs inclusive t bounds)."""
    sc = scenes.spectral_pyramid()
    org, d = scenes.random_rays(1 << 15, (0, 0, 0), (100, 100, 100), seed=4)
    ids, _ = _check(ctx, oracle_mod, sc, org, d)
    assert (ids >= 12).any() and (ids < 12).any()
    # rays starting inside spheres (second root) and tangent-ish rays
    c = np.array(scenes._PYRAMID_SPHERES, dtype=np.float64)
    o2 = np.repeat(c, 200, axis=0) + 0.5
    _, d2 = scenes.random_rays(len(o2), (0, 0, 0), (1, 1, 1), seed=9)
    _check(ctx, oracle_mod, sc, o2, d2)


def test_slice_world_with_wrappers(ctx, oracle_mod):
    """Config 1's geometry: rects, flipped rects, sphere and Translate(RotateY(Box)) in a HitableSlice."""
    sc = scenes.cornell_box()
    org, d = scenes.random_rays(1 << 15, (0, 0, 0), (555, 555, 555), seed=6)
    ids, _ = _check(ctx, oracle_mod, sc, org, d, compare_prim_counts=False)
    assert set(np.unique(ids)) >= {0, 1, 3, 4, 5, 6, 7}
    sc.world_kind = S.WORLD_BVH4  # same objects as BVH4 leaves
    _check(ctx, oracle_mod, sc, org, d, compare_prim_counts=False)


def test_full_size_config2_properties(ctx, oracle_mod):
    """BASELINE config 2 at full size (1M triangles, 16M rays): determinism, t/ID consistency and
    oracle agreement on a strided 1/256 sample."""
    sc, lo, hi = scenes.closest_hit_scene()
    hs = cuda.HostScene(sc)
    ctx.upload(hs)
    n = 1 << 24
    org, d = scenes.random_rays(n, lo, hi)
    ids, t = ctx.trace_closest(org, d)
    ids2, t2 = ctx.trace_closest(org, d)
    assert ids.tobytes() == ids2.tobytes() and t.tobytes() == t2.tobytes()
    hit = ids >= 0
    assert 0.3 < hit.mean() < 0.7
    assert (ids[hit] < 1_000_000).all() and (t[hit] >= 0.001).all() and (t[~hit] == 0).all()
    # a ray shortened to just before its hit must miss that primitive; extended past it, it must hit it again
    sel = np.flatnonzero(hit)[:: 4096]
    ids3, t3 = ctx.trace_closest(org[sel], d[sel], 0.001, DMAX)
    assert ids3.tobytes() == ids[sel].tobytes() and t3.tobytes() == t[sel].tobytes()
    ids4, _ = ctx.trace_closest(org[sel], d[sel], 0.001, 1.0)
    assert ((ids4 == -1) | (t[sel] <= 1.0)).all()
    osn = oracle_mod.OracleScene(sc)
    s = slice(0, n, 256)
    oi, ot = osn.trace(org[s], d[s])
    assert oi.tobytes() == ids[s].tobytes() and ot.tobytes() == t[s].tobytes()


def test_fp32_mode_agrees_closely(ctx, oracle_mod):
    """Optional IZPI_TRACE_FP32 mode (fp32 triangle test): not bit-exact by design, reported separately; it must
    still find the same primitive for nearly every ray and t to fp32 accuracy."""
    verts, uvs = scenes.torus_mesh(300, 200)
    sc = S.SceneSpec(bvh_seed=12345)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org, d = scenes.random_rays(1 << 17, lo, hi)
    ctx.upload(cuda.HostScene(sc))
    ids, t = ctx.trace_closest(org, d)
    ids32, t32 = ctx.trace_closest(org, d, mode=cuda.TRACE_FP32)
    same = ids == ids32
    assert same.mean() > 0.999
    hit = same & (ids >= 0)
    assert np.allclose(t32[hit], t[hit], rtol=1e-4)
    # a slice world has no fp32 path: loud error, no silent fallback
    ctx.upload(cuda.HostScene(scenes.cornell_box()))
    with pytest.raises(cuda.IzpiError):
        ctx.trace_closest(org[:8], d[:8], mode=cuda.TRACE_FP32)


def test_lane_layouts_agree(oracle_mod, monkeypatch):
    """The 2-lanes-per-ray kernel (default when the tree's stack bound fits its slab) and the 4-lanes-per-ray kernel (deep
    trees, IZPI_TRACE_LANES=4) follow the same traversal: identical answers AND identical visit / test counts, on a ragged
    batch size that leaves pairs and groups idle at the tail."""
    verts, uvs = scenes.torus_mesh(200, 120)
    sc = S.SceneSpec(bvh_seed=12345)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    r = np.random.default_rng(1)
    for _ in range(50):
        sc.sphere(r.uniform(20, 80, 3), r.uniform(0.5, 3.0), 0)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org, d = scenes.random_rays((1 << 17) + 13, lo, hi)
    hs = cuda.HostScene(sc)
    out = {}
    for lanes in ("2", "4"):
        monkeypatch.setenv("IZPI_TRACE_LANES", lanes)
        c = cuda.Context(0)
        c.upload(hs)
        ids, t, st = c.trace_closest(org, d, stats=True)
        ids2, t2 = c.trace_closest(org, d)
        ids3, t3 = c.trace_closest(org, d, mode=cuda.TRACE_FP32)
        assert ids2.tobytes() == ids.tobytes() and t2.tobytes() == t.tobytes()
        out[lanes] = (ids, t, st["nodes"], st["prims"], ids3)
        c.close()
    assert out["2"][0].tobytes() == out["4"][0].tobytes() and out["2"][1].tobytes() == out["4"][1].tobytes()
    assert out["2"][2] == out["4"][2] and out["2"][3] == out["4"][3]
    assert out["2"][4].tobytes() == out["4"][4].tobytes()  # the fp32 variants agree with each other as well
    oi, ot, ost = oracle_mod.OracleScene(sc).trace(org, d, stats=True)
    assert out["2"][0].tobytes() == oi.tobytes() and out["2"][1].tobytes() == ot.tobytes()
    assert out["2"][2] == ost["nodes"] and out["2"][3] == ost["tris"] + ost["spheres"] + ost["others"]


def test_fma_peak_diagnostic(ctx):
    """The FLOP side of the roofline is measured, not assumed: plausible B200 vector peaks (nominal 80 / 40 TFLOP/s)."""
    f32, f64 = ctx.fma_peak(False), ctx.fma_peak(True)
    assert 20.0 < f32 < 120.0 and 10.0 < f64 < 60.0 and f32 > f64
