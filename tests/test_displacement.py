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
