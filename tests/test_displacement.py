"""displacement.ApplyDisplacementMap (internal/displacement/displacement.go:145-280): oracle properties on CPU, and the
device implementation (izpi_displace) against the oracle, bit for bit, on the GPU."""
import numpy as np
import pytest

from izpi_b200 import cuda, scenes


def _heightmap(n=128, seed=3):
    px = np.ones((n, n, 4))
    px[..., 2] = scenes._value_noise(n, 8, seed)
    return px


def _quad(size=100.0):
    # two triangles of a quad on the XZ plane with the full UV square (the reference's own test data uses unit quads too)
    return np.array([[0, 0, 0, size, 0, 0, size, 0, size, 0, 0, 1, 0, 1, 1],
                     [0, 0, 0, size, 0, size, 0, 0, size, 0, 0, 1, 1, 0, 1]], dtype=np.float64)


def test_oracle_tessellate_4_children(oracle_mod):
    """displacement_test.go pins tessellate(): a triangle splits into 4 with edge midpoints; a flat map stops at the
    UV limit only (4 texels), so the count is a power of 4 per input triangle."""
    px = np.ones((64, 64, 4)) * 0.5
    out, mats = oracle_mod.apply_displacement(_quad(), [3, 7], px, 0.0, 1.0)
    # maxDeltaU = 4/63: the unit UV edge halves until <= 0.0635 -> 4 levels -> 4^4 per triangle
    assert len(out) == 2 * 4 ** 4 and set(mats[: 4 ** 4]) == {3} and set(mats[4 ** 4:]) == {7}
    # flat map value 0.5, min 0 max 1: every vertex moves by 0.5 along the triangle normal
    # edge1 x edge2 of the XZ quad points to -Y (vec3.Cross, vec3.go:104), so the displacement is -0.5 in Y
    np.testing.assert_allclose(out[:, [1, 4, 7]], -0.5, atol=1e-12)
    # UVs stay inside the parent's UV triangle and areas add up (before displacement the XZ footprint is preserved)
    a = 0.5 * np.abs((out[:, 3] - out[:, 0]) * (out[:, 8] - out[:, 2]) - (out[:, 6] - out[:, 0]) * (out[:, 5] - out[:, 2]))
    np.testing.assert_allclose(a.sum(), 100.0 * 100.0, rtol=1e-12)


def test_oracle_per_triangle_is_a_regrouping(oracle_mod):
    px = _heightmap()
    # valid map for this range: adjacent texels differ by less than threshold / |max - min| (else the reference never stops)
    step = max(np.abs(np.diff(px[..., 2], axis=0)).max(), np.abs(np.diff(px[..., 2], axis=1)).max())
    assert step * 15.0 < 2.0
    a, am = oracle_mod.apply_displacement(_quad(), [0, 1], px, 0.0, 15.0, per_triangle=True)
    b, bm = oracle_mod.apply_displacement(_quad(), [0, 1], px, 0.0, 15.0, per_triangle=False)
    assert len(a) == len(b) > 2 * 4 ** 5   # the adaptive criterion refined beyond the UV limit (4^5 at 128 texels) somewhere
    assert sorted(map(bytes, a)) == sorted(map(bytes, b))
    assert (np.diff(am) >= 0).all() and not (np.diff(bm) >= 0).all()


@pytest.fixture(scope="module")
def ctx():
    from izpi_b200.build import build
    build()
    c = cuda.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("per_triangle", [True, False])
@pytest.mark.parametrize("dmax", [1.0, 14.0])
def test_device_displacement_bit_exact(ctx, oracle_mod, per_triangle, dmax):
    px = _heightmap(128, 5)
    verts, uvs = scenes.torus_mesh(12, 6)  # 144 triangles with real UVs
    tris = np.concatenate([verts.reshape(-1, 9), uvs.reshape(-1, 6)], axis=1)
    mats = np.arange(len(tris), dtype=np.int32) % 5
    got, gm = ctx.apply_displacement(tris, mats, px, -1.0, dmax, per_triangle=per_triangle)
    want, wm = oracle_mod.apply_displacement(tris, mats, px, -1.0, dmax, per_triangle=per_triangle)
    assert got.shape == want.shape and len(got) > len(tris)
    assert got.tobytes() == want.tobytes()
    np.testing.assert_array_equal(gm, wm)


@pytest.mark.gpu
def test_device_displacement_edge_cases(ctx, oracle_mod):
    px = _heightmap(32, 9)
    out, om = ctx.apply_displacement(np.zeros((0, 15)), np.zeros(0, dtype=np.int32), px, 0, 1)
    assert out.shape == (0, 15) and om.shape == (0,)
    # degenerate UVs (all zero): f = 1/0 -> NaN tangent; the reference's matrix product then yields NaN vertices
    t = _quad()
    t[:, 9:] = 0.0
    got, _ = ctx.apply_displacement(t, [0, 0], px, 0, 1)
    want, _ = oracle_mod.apply_displacement(t, [0, 0], px, 0, 1)
    assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
    with pytest.raises(cuda.IzpiError):
        ctx.apply_displacement(_quad(), [0, 0], np.ones((1, 1, 4)), 0, 1)  # resU-1 == 0 in displacement.go:177


# ---- the reference's own unit vectors (internal/displacement/displacement_test.go), replayed on the oracle ----------------
def _tri15(v0, v1, v2, u0=0.0, v0_=0.0, u1=0.0, v1_=0.0, u2=0.0, v2_=0.0):
    return np.array([*v0, *v1, *v2, u0, v0_, u1, v1_, u2, v2_], dtype=np.float64)


_XY_TRIANGLE = _tri15((-1, 0, 0), (1, 0, 0), (0, 1, 0), u1=1.0, u2=0.5, v2_=1.0)  # displacement_test.go:21-28


def test_reference_vectors_tessellate(oracle_mod):
    """TestTessellate (displacement_test.go:13-82): the four children of the XY-plane triangle, exact."""
    got = oracle_mod.displacement_tessellate(_XY_TRIANGLE)
    want = np.array([
        _tri15((-1, 0, 0), (0, 0, 0), (-0.5, 0.5, 0), u1=0.5, u2=0.25, v2_=0.5),
        _tri15((0, 0, 0), (0.5, 0.5, 0), (-0.5, 0.5, 0), u0=0.5, u1=0.75, u2=0.25, v1_=0.5, v2_=0.5),
        _tri15((0, 0, 0), (1, 0, 0), (0.5, 0.5, 0), u0=0.5, u1=1.0, u2=0.75, v2_=0.5),
        _tri15((-0.5, 0.5, 0), (0.5, 0.5, 0), (0, 1, 0), u0=0.25, u1=0.75, u2=0.5, v0_=0.5, v1_=0.5, v2_=1.0),
    ])
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("res_u,res_v,want", [(3, 3, 4), (4, 2, 16), (2, 4, 16)])
def test_reference_vectors_apply_tessellation(oracle_mod, res_u, res_v, want):
    """TestApplyTessellation (displacement_test.go:84-157): flat map (0, 0, 0.5), min 0, max 1, threshold 0.01,
    maxDelta = 1 / (res - 1)."""
    n = oracle_mod.displacement_apply_tessellation(_XY_TRIANGLE, 1.0 / (res_u - 1), 1.0 / (res_v - 1), (0.0, 0.0, 0.5), 0.0, 1.0, 0.01)
    assert n == want


@pytest.mark.parametrize("rgb,dmin,dmax,y", [((0.0, 0.0, 1.0), 0.0, 1.0, -1.0), ((1.0, 1.0, 1.0), -0.5, 0.5, -0.5)])
def test_reference_vectors_apply_displacement(oracle_mod, rgb, dmin, dmax, y):
    """TestApplyDisplacement (displacement_test.go:159-213): the XZ-plane triangle moves along its normal (-Y) by
    min + (max - min) * blue."""
    tri = _tri15((-1, 0, 0), (1, 0, 0), (0, 0, 1), u1=1.0, u2=0.5, v2_=1.0)
    got = oracle_mod.displacement_apply_displacement(tri, rgb, dmin, dmax)
    want = _tri15((-1, y, 0), (1, y, 0), (0, y, 1), u1=1.0, u2=0.5, v2_=1.0)
    assert got.reshape(15).tobytes() == want.tobytes()


@pytest.mark.gpu
def test_device_reproduces_the_reference_vectors(ctx, oracle_mod):
    """The same vectors through izpi_displace: a 3x3 flat map gives maxDelta = 4 / 2, so one tessellation level (4 children of
    TestTessellate, displaced by the constant height) -- vertices must equal tessellate() + applyDisplacement() of the oracle."""
    px = np.zeros((3, 3, 4)); px[..., 2] = 0.5; px[..., 3] = 1.0
    out, _ = ctx.apply_displacement(_XY_TRIANGLE[None], [0], px, 0.0, 1.0)
    kids = oracle_mod.displacement_tessellate(_XY_TRIANGLE)
    want = oracle_mod.displacement_apply_displacement(kids, (0.0, 0.0, 0.5), 0.0, 1.0)
    assert out.shape == (4, 15) and out.tobytes() == want.tobytes()
    # XY-plane triangle, normal +Z: every vertex rises by 0.5
    np.testing.assert_array_equal(out[:, [2, 5, 8]], 0.5)
