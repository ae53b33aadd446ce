"""Render parity on the B200: the wavefront path tracer (through the C ABI, izpi_host_render) against
the oracle's recursive integrators.

Two kinds of check:
  * same-path: the oracle runs with the device's counter RNG (rng_mode=1), so both sides walk the
    same paths; pixels agree to ~1e-9 except where an ulp-level libm difference (sin/cos/pow/exp on
    the device are not correctly rounded) flips a branch.  Tolerances are written below.
  * converged: the oracle runs with the reference's LCG streams (rng_mode=0), independent noise;
    per-channel relative RMSE of box-filtered images must be < 1 % (BASELINE.json north_star bar,
    applied at test-sized resolution).
"""
import numpy as np
import pytest

from izpi_b200 import cuda, scenes
from izpi_b200 import scene as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from izpi_b200.build import build
    build()
    c = cuda.Context(0)
    yield c
    c.close()


def _same_path(ctx, oracle_mod, spec, w, h, spp, sampler, seed=5, min_close=0.97, max_depth=50):
    ctx.upload(cuda.HostScene(spec))
    img, rays = ctx.render(w, h, spp, max_depth=max_depth, sampler=sampler, seed=seed)
    ref, ref_rays = oracle_mod.OracleScene(spec).render(w, h, spp, max_depth=max_depth, sampler=sampler, rng_mode=1, seed=seed)
    assert img.shape == ref.shape
    # row 0 is never written, row ny-y flip (rgb.go:41)
    assert (img[0] == 0).all() and (ref[0] == 0).all()
    assert (img[1:, :, 3] == 1).all()
    fin = np.isfinite(ref).all(axis=-1) & np.isfinite(img).all(axis=-1)
    close = np.isclose(img, ref, rtol=1e-7, atol=1e-10).all(axis=-1) & fin
    frac = close[1:].mean()
    assert frac >= min_close, f"only {frac:.3%} of pixels agree"
    # image-level agreement, robust to the few diverged paths
    m_img, m_ref = img[1:][fin[1:]][:, :3].mean(0), ref[1:][fin[1:]][:, :3].mean(0)
    assert np.allclose(m_img, m_ref, rtol=0.05), (m_img, m_ref)
    assert abs(rays - ref_rays) <= 0.01 * ref_rays
    return img, ref, frac


def test_cornell_same_path(ctx, oracle_mod):
    """Config 1 geometry/materials: Lambert, DiffuseLight, Dielectric, rect/sphere light sampling."""
    img, ref, frac = _same_path(ctx, oracle_mod, scenes.cornell_box(1.0), 48, 48, 8, cuda.SAMPLER_COLOUR)
    assert img[1:, :, :3].mean() > 0.05


def test_spectral_pyramid_same_path(ctx, oracle_mod):
    """Config 4: spectral sampler, dispersion, Beer-Lambert with the nested path-length trace,
    triangle + sphere lights, firefly rejection + XYZ->ACEScg epilogue."""
    _same_path(ctx, oracle_mod, scenes.spectral_pyramid(1.0), 40, 40, 16, cuda.SAMPLER_SPECTRAL, min_close=0.93)


def test_pbr_mesh_same_path(ctx, oracle_mod):
    """Config 3 materials at test size: PBR with albedo/normal/roughness/metalness image textures."""
    sc = scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64)
    _same_path(ctx, oracle_mod, sc, 40, 40, 8, cuda.SAMPLER_COLOUR, min_close=0.95)


def test_metal_and_coloured_glass_same_path(ctx, oracle_mod):
    sc = scenes.cornell_box(1.0)
    sc.world_kind = S.WORLD_BVH4  # SetWorld happens on the transport path only (transport.go:83-89)
    metal = sc.metal((0.8, 0.85, 0.88), 0.1)
    glass = sc.colored_dielectric(1.5, (0.02, 0.005, 0.001))
    sc.sphere((400, 60, 150), 60, metal)
    sc.sphere((120, 300, 300), 70, glass)
    _same_path(ctx, oracle_mod, sc, 40, 40, 8, cuda.SAMPLER_COLOUR, min_close=0.95)


def test_max_depth_term(ctx, oracle_mod):
    """depth >= maxDepth returns (0,0,1) (colour.go:34-36): visible with a tiny depth cap."""
    img, ref, _ = _same_path(ctx, oracle_mod, scenes.cornell_box(1.0), 32, 32, 4, cuda.SAMPLER_COLOUR, max_depth=2)
    assert img[1:, :, 2].mean() > img[1:, :, 0].mean()


def test_tile_subset_and_sample_ranges(ctx, oracle_mod):
    """Tiles and sample ranges partition the work: rendering them separately and summing equals one pass."""
    sc = scenes.cornell_box(1.0)
    ctx.upload(cuda.HostScene(sc))
    w = h = 40  # common.Tiles -> 20x20 tiles, 4 tiles
    full, rays = ctx.render(w, h, 8, seed=9)
    # tile halves (disjoint pixels): bit-identical where written
    a, ra = ctx.render(w, h, 8, seed=9, tile_begin=0, tile_end=2)
    b, rb = ctx.render(w, h, 8, seed=9, tile_begin=2, tile_end=4)
    assert ra + rb == rays
    assert ((a[..., 3] == 1) ^ (b[..., 3] == 1))[1:].all()
    np.testing.assert_array_equal(np.where(a[..., 3:] == 1, a, b), full)
    # sample ranges: sums of partial means equal the full mean up to fp64 summation order
    s0, _ = ctx.render(w, h, 8, seed=9, sample_offset=0, sample_count=4)
    s1, _ = ctx.render(w, h, 8, seed=9, sample_offset=4, sample_count=4)
    np.testing.assert_allclose((s0 + s1)[..., :3], full[..., :3], rtol=1e-12, atol=1e-15)


def _box(img, k):
    h, w = img.shape[0] // k * k, img.shape[1] // k * k
    return img[:h, :w].reshape(h // k, k, w // k, k, -1).mean(axis=(1, 3))


def test_cornell_converged_rmse(ctx, oracle_mod):
    """Converged render vs the oracle with the reference's own LCG streams (independent noise)."""
    sc = scenes.cornell_box(1.0)
    ctx.upload(cuda.HostScene(sc))
    w = h = 64
    spp = 4096
    # the device side is cheap: 4x the samples, so that the residual is the oracle's own 4096-spp noise
    img, _ = ctx.render(w, h, 4 * spp, seed=11)
    osn = oracle_mod.OracleScene(sc)
    ref, _ = osn.render(w, h, spp, rng_mode=0, seed=12)
    ref2, _ = osn.render(w, h, spp, rng_mode=0, seed=13)
    a, b, b2 = _box(img[1:], 7)[..., :3], _box(ref[1:], 7)[..., :3], _box(ref2[1:], 7)[..., :3]
    for c in range(3):
        rel_rmse = np.sqrt(np.mean((a[..., c] - b[..., c]) ** 2)) / np.mean(b[..., c])
        floor = np.sqrt(np.mean((b2[..., c] - b[..., c]) ** 2)) / np.mean(b[..., c])  # oracle vs oracle
        assert rel_rmse < 0.01, (c, rel_rmse)
        assert rel_rmse < floor * 1.1, (c, rel_rmse, floor)  # no bias beyond the oracle's own noise


def test_render_errors(ctx):
    sc = S.SceneSpec()
    sc.triangles(np.array([[(0, 0, 0), (1, 0, 0), (0, 1, 0)]], dtype=np.float64), sc.lambertian(sc.constant_texture((1, 1, 1))))
    sc.set_camera((0, 0, -5), (0, 0, 0), (0, 1, 0), 40, 1.0)
    ctx.upload(cuda.HostScene(sc))
    with pytest.raises(cuda.IzpiError) as e:  # no emitters: HitableSlice.Random would index an empty slice
        ctx.render(16, 16, 1)
    assert "no emitters" in str(e.value)
    ctx.upload(cuda.HostScene(scenes.cornell_box(1.0)))
    with pytest.raises(cuda.IzpiError):  # common.Tiles finds no divisor (the reference divides by zero)
        ctx.render(17, 16, 1)


def test_ibl_metal_mesh_same_path(ctx, oracle_mod):
    """Config 5 materials/geometry at test size: sky dome = FlipNormals(Sphere) with an image DiffuseLight,
    Metal mesh, glass sphere; all specular, as in scenes.Environment."""
    sc = scenes.ibl_displaced_mesh(16 / 9, 120, 60, (256, 128))
    _same_path(ctx, oracle_mod, sc, 64, 36, 8, cuda.SAMPLER_COLOUR, min_close=0.97)


def test_spectral_pbr_image_albedo_same_path(ctx, oracle_mod):
    """Spectral PBR (pbr.go:158-263) with a SpectralImage albedo (spectral_image.go:61-245), the material
    transport builds for PBR in SPECTRAL scenes (transport.go:209-250)."""
    sc = scenes.spectral_pyramid(1.0)
    alb, rough, metal, nrm = scenes.pbr_textures(64)
    a = sc.image_texture(alb)
    pbr = sc.pbr(a, sc.image_texture(nrm), sc.image_texture(rough), sc.image_texture(metal), spectral_albedo=sc.spectral_image(a))
    tv, tuv = scenes.torus_mesh(60, 40, centre=(50.0, 60.0, 50.0), major=30.0, minor=12.0, amp=3.0, scale=0.5)
    sc.triangles(tv, pbr, tuv)
    _same_path(ctx, oracle_mod, sc, 40, 40, 16, cuda.SAMPLER_SPECTRAL, min_close=0.93)


@pytest.mark.parametrize("sampler", [cuda.SAMPLER_ALBEDO, cuda.SAMPLER_NORMAL], ids=["albedo", "normal"])
def test_aov_samplers_same_path(ctx, oracle_mod, sampler):
    """Debug AOV samplers (sampler/albedo.go, sampler/normal.go): first-hit albedo / normal, black on a miss."""
    for spec in (scenes.cornell_box(1.0), scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64)):
        img, ref, frac = _same_path(ctx, oracle_mod, spec, 40, 40, 4, sampler, min_close=0.995)
        assert np.abs(img[1:, :, :3]).max() > 0


def test_worker_tile_rows(ctx):
    """worker.RenderTile's wire shape (worker/render.go:17-75): rows in image order (no flip), strip_height padding, and the
    y == 0 row that the local path drops.  Pixel values are the local render's (same RNG keys, same kernels)."""
    spec = scenes.cornell_box(1.0)
    # a narrower field of view than scenes.go:119-155, so that the bottom image row (y == 0) looks into the box, not under it
    spec.set_camera((278.0, 278.0, -800.0), (278, 278, 0), (0, 1, 0), 30.0, 1.0, 0.0, 10.0, 0.0, 1.0, 1.0)
    ctx.upload(cuda.HostScene(spec))
    w = h = 50  # common.Tiles -> 25 x 25 tiles
    img, _ = ctx.render(w, h, 4, sampler=cuda.SAMPLER_COLOUR, seed=9)
    ctx.render_setup(w, h, 4, sampler=cuda.SAMPLER_COLOUR, seed=9)
    rows = ctx.render_tile_rows(25, 0, 49, 24, strip_height=1)
    assert rows.shape == (25, 100)
    for r in range(1, 25):  # image row y sits at canvas row ny - y (rgb.go:41, remote.go:63-69)
        assert rows[r].tobytes() == img[h - r, 25:50].tobytes()
    assert np.isfinite(rows[0]).all() and (rows[0].reshape(-1, 4)[:, 3] == 1).all() and rows[0].reshape(-1, 4)[:, :3].max() > 0
    rows2 = ctx.render_tile_rows(0, 25, 24, 49, strip_height=3)
    assert rows2.shape == (25, 300)
    assert (rows2[:, 100:] == 0).all()  # make([]float64, stripSize): only the first row of the strip is filled
    for r in range(25):
        assert rows2[r, :100].tobytes() == img[h - (25 + r), 0:25].tobytes()
    with pytest.raises(cuda.IzpiError):
        ctx.render_tile_rows(0, 0, 24, 24, strip_height=0)
    with pytest.raises(cuda.IzpiError):
        ctx.render_tile_rows(0, 0, 50, 24)
