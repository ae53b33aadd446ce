"""Render parity on the B200: the wavefront path tracer (through the C ABI, izpi_host_render) against
the oracle's recursive integrators.

Two kinds of check:
  * same-path: the oracle runs with the device's counter RNG (rng_mode=1), so both sides walk the
    same paths.  EVERY pixel must agree to 1e-8 relative and the ray totals must be equal: the only arithmetic that is not
    bit-identical is the device's libm (sin / cos / pow / exp / atan2 / asin within 2 ulp of glibc) and the summation order
    of the iterative estimator, both at the 1e-15 level (scripts/parity_probe.py measured no pixel beyond 1e-9 on any scene;
    round 1's 93-97 % floors were never needed).  The oracle's libm_jitter probe (+-2 ulp on every libm result) confirms that
    no pixel of these frames is sensitive at that level, so a disagreement would be a real difference, not a rounding flip.
  * converged: the oracle runs with the reference's LCG streams (rng_mode=0), independent noise, 4096 spp, no filtering;
    per-channel relative RMSE of the converged images < 1 % (BASELINE.json north_star), estimated without the reference
    image's own Monte-Carlo noise (see _converged).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from izpi_b200 import cuda, render, scenes
from izpi_b200 import scene as S

pytestmark = pytest.mark.gpu

SAME_PATH_RTOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    from izpi_b200.build import build
    build()
    c = cuda.Context(0)
    yield c
    c.close()


def _same_path(ctx, oracle_mod, spec, w, h, spp, sampler, seed=5, max_depth=50, jitter_probe=False):
    ctx.upload(cuda.HostScene(spec))
    img, rays = ctx.render(w, h, spp, max_depth=max_depth, sampler=sampler, seed=seed)
    osn = oracle_mod.OracleScene(spec)
    ref, ref_rays = osn.render(w, h, spp, max_depth=max_depth, sampler=sampler, rng_mode=1, seed=seed)
    assert img.shape == ref.shape
    # row 0 is never written, row ny-y flip (rgb.go:41)
    assert (img[0] == 0).all() and (ref[0] == 0).all()
    assert (img[1:, :, 3] == 1).all()
    assert np.isfinite(ref).all() and np.isfinite(img).all()
    close = np.isclose(img, ref, rtol=SAME_PATH_RTOL, atol=1e-14).all(axis=-1)
    assert close.all(), f"{(~close).sum()} of {close.size} pixels differ beyond {SAME_PATH_RTOL}: worst {np.abs(img - ref).max()}"
    assert rays == ref_rays  # same paths, same number of Hit calls (colour.go:38)
    if jitter_probe:  # the frame is insensitive to last-bit libm differences: nothing here may hide behind "rounding"
        for j in (1, 2):
            jit, _ = osn.render(w, h, spp, max_depth=max_depth, sampler=sampler, rng_mode=1, seed=seed, libm_jitter=j)
            assert np.isclose(jit, ref, rtol=SAME_PATH_RTOL, atol=1e-14).all()
    return img, ref


def test_cornell_same_path(ctx, oracle_mod):
    """Config 1 geometry/materials: Lambert, DiffuseLight, Dielectric, rect/sphere light sampling."""
    img, ref = _same_path(ctx, oracle_mod, scenes.cornell_box(1.0), 48, 48, 8, cuda.SAMPLER_COLOUR, jitter_probe=True)
    assert img[1:, :, :3].mean() > 0.05


def test_thin_lens_same_path(ctx, oracle_mod):
    """aperture > 0: camera.randomInUnitDisc + lens offset (camera.go:61-89); no BASELINE scene uses it, the reference's
    scenes.go examples do."""
    spec = scenes.cornell_box(1.0)
    spec.set_camera((278.0, 278.0, -800.0), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 40.0, 1000.0, 0.0, 1.0, 1.0)
    img, ref = _same_path(ctx, oracle_mod, spec, 48, 48, 8, cuda.SAMPLER_COLOUR)
    spec0 = scenes.cornell_box(1.0)
    ctx.upload(cuda.HostScene(spec0))
    pin, _ = ctx.render(48, 48, 8, sampler=cuda.SAMPLER_COLOUR, seed=5)
    assert not np.array_equal(pin, img)  # the lens does something


def test_spectral_pyramid_same_path(ctx, oracle_mod):
    """Config 4: spectral sampler, dispersion, Beer-Lambert with the nested path-length trace,
    triangle + sphere lights, firefly rejection + XYZ->ACEScg epilogue."""
    _same_path(ctx, oracle_mod, scenes.spectral_pyramid(1.0), 40, 40, 16, cuda.SAMPLER_SPECTRAL, jitter_probe=True)


def test_coherence_sort_forced_same_path(ctx, oracle_mod):
    """The coherence sort of the thread-per-ray wavefront (queue sorted by origin / direction cell between bounces) only kicks in
    above 2^18 live paths; forced on for every bounce here (IZPI_SORT_RAYS=2), the frame still equals the oracle's pixel for
    pixel, and the unsorted frame bit for bit."""
    spec = scenes.spectral_pyramid(1.0)
    os.environ["IZPI_SORT_RAYS"] = "2"
    try:
        sorted_img, _ = _same_path(ctx, oracle_mod, spec, 40, 40, 16, cuda.SAMPLER_SPECTRAL)
        os.environ["IZPI_SORT_RAYS"] = "0"
        plain, _ = ctx.render(40, 40, 16, sampler=cuda.SAMPLER_SPECTRAL, seed=5)
    finally:
        del os.environ["IZPI_SORT_RAYS"]
    assert plain.tobytes() == sorted_img.tobytes()


def test_pbr_mesh_same_path(ctx, oracle_mod):
    """Config 3 materials at test size: PBR with albedo/normal/roughness/metalness image textures."""
    sc = scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64)
    _same_path(ctx, oracle_mod, sc, 40, 40, 8, cuda.SAMPLER_COLOUR)


def test_many_pbr_materials_sorted_by_material(ctx, oracle_mod):
    """Six PBR materials with their own image textures: hits are filed in one bin per material between bounces (north_star 3:
    sort by material ID); same paths, same pixels as the oracle, and as with the class-only bins."""
    sc = scenes.cornell_pbr_mesh(1.0, n_around=96, n_tube=40, tex_size=64, n_pbr_materials=6)
    img, _ = _same_path(ctx, oracle_mod, sc, 48, 48, 8, cuda.SAMPLER_COLOUR)
    r = render.New(ctx, 48, 48, 8, 50, sampler_type=cuda.SAMPLER_COLOUR, seed=5)
    r.Render()
    assert ctx.render_stats()["material_bins"] == 2 + 6  # Lambert + DiffuseLight share per class, one bin per textured PBR material
    os.environ["IZPI_MATERIAL_BINS"] = "0"
    try:
        ctx.upload(cuda.HostScene(sc))
        flat, _ = ctx.render(48, 48, 8, sampler=cuda.SAMPLER_COLOUR, seed=5)
        render.New(ctx, 48, 48, 1, 50, seed=5).Render()
        assert ctx.render_stats()["material_bins"] == 3
    finally:
        del os.environ["IZPI_MATERIAL_BINS"]
    assert flat.tobytes() == img.tobytes()


def test_metal_and_coloured_glass_same_path(ctx, oracle_mod):
    sc = scenes.cornell_box(1.0)
    sc.world_kind = S.WORLD_BVH4  # SetWorld happens on the transport path only (transport.go:83-89)
    metal = sc.metal((0.8, 0.85, 0.88), 0.1)
    glass = sc.colored_dielectric(1.5, (0.02, 0.005, 0.001))
    sc.sphere((400, 60, 150), 60, metal)
    sc.sphere((120, 300, 300), 70, glass)
    _same_path(ctx, oracle_mod, sc, 40, 40, 8, cuda.SAMPLER_COLOUR)


def test_max_depth_term(ctx, oracle_mod):
    """depth >= maxDepth returns (0,0,1) (colour.go:34-36): visible with a tiny depth cap."""
    img, ref = _same_path(ctx, oracle_mod, scenes.cornell_box(1.0), 32, 32, 4, cuda.SAMPLER_COLOUR, max_depth=2)
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


def _device_frames(ctx, w, h, spp, sampler, seeds):
    """Device frames WITHOUT the spectral epilogue (FireflyRejection is a non-linear 3x3 filter whose effect depends on the
    sample count; converged parity is about the estimator, the epilogue kernels are checked by the same-path tests)."""
    L = cuda.lib()
    tiles = render.tile_list(w, h)
    out = []
    for seed in seeds:
        cfg = cuda.RenderConfig(width=w, height=h, spp=spp, max_depth=50, sampler=sampler, sample_offset=0, sample_count=spp, seed=seed)
        cuda.check(L.izpi_render_setup(ctx._h, C.byref(cfg)))
        canvas = np.zeros((h, w, 4), dtype=np.float64)
        cuda.check(L.izpi_render_tiles(ctx._h, len(tiles), tiles.ctypes.data, canvas.ctypes.data))
        out.append(canvas)
    return out


def _converged(ctx, oracle_mod, name, spec, w, h, sampler, spp=4096, device_factor=4):
    """Converged render vs the oracle with the reference's own LCG streams at `spp` samples per pixel, full resolution, no
    filter.  The north-star bar is per-channel relative RMSE < 1 % between the CONVERGED images.  A 4096-spp reference frame
    still carries several per cent of Monte-Carlo noise per pixel, so RMSE(device, reference) mostly measures that noise.
    The oracle's frame is therefore rendered as two independent halves b1, b2 (spp/2 each; their mean is the spp-sample
    frame) and the device renders two independent frames a1, a2: E[(a1-b1)(a2-b2)] = (A-B)^2 per pixel, whatever the noise,
    so sqrt(mean((a1-b1)(a2-b2))) / mean(B) is an unbiased estimate of the converged images' relative RMSE.
    Asserted (tolerances are the test's):
      * that estimate is < 1 %, within three standard errors of the estimate itself;
      * image means agree to 0.5 % (bias at the level of the whole frame, 67 M samples);
      * the plain RMSE of device vs the spp-sample oracle frame is explained by the oracle's own noise (split-half floor)."""
    ctx.upload(cuda.HostScene(spec))
    a1, a2 = _device_frames(ctx, w, h, device_factor * spp, sampler, (11, 12))
    osn = oracle_mod.OracleScene(spec)
    b1, _ = osn.render(w, h, spp // 2, sampler=sampler, rng_mode=0, seed=21, epilogue=False)
    b2, _ = osn.render(w, h, spp // 2, sampler=sampler, rng_mode=0, seed=22, epilogue=False)
    report = {}
    for c in range(3):
        A1, A2, B1, B2 = a1[1:, :, c], a2[1:, :, c], b1[1:, :, c], b2[1:, :, c]
        B = 0.5 * (B1 + B2)
        mean_b = B.mean()
        terms = (A1 - B1) * (A2 - B2)
        est2, se2 = terms.mean(), terms.std() / np.sqrt(terms.size)
        rel_est = np.sqrt(max(est2, 0.0)) / mean_b
        rel_lo = np.sqrt(max(est2 - 3.0 * se2, 0.0)) / mean_b
        raw = np.sqrt(np.mean((0.5 * (A1 + A2) - B) ** 2)) / mean_b
        floor = np.sqrt(np.mean((B1 - B2) ** 2)) / 2.0 / mean_b  # noise of the spp-sample oracle frame
        bias = abs(0.5 * (A1 + A2).mean() - mean_b) / mean_b
        report[c] = dict(converged_rel_rmse=float(rel_est), lower_3se=float(rel_lo), raw_rel_rmse=float(raw), oracle_noise_floor=float(floor),
                         mean_bias=float(bias), mean=float(mean_b))
        assert rel_lo < 0.01, (name, c, report[c])
        assert bias < 0.005, (name, c, report[c])
        assert raw < 1.2 * floor + 0.01, (name, c, report[c])
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", f"converged_{name}.json"), "w") as f:
        json.dump(report, f, indent=1)
    return report


def test_converged_cornell(ctx, oracle_mod):
    """Config 1 at 128 x 128, 4096 spp."""
    _converged(ctx, oracle_mod, "cornell", scenes.cornell_box(1.0), 128, 128, cuda.SAMPLER_COLOUR)


def test_converged_pbr_mesh(ctx, oracle_mod):
    """Config 3's materials and geometry (Cornell walls + PBR torus with the four image textures) at 128 x 128, 4096 spp."""
    _converged(ctx, oracle_mod, "pbr_mesh", scenes.cornell_pbr_mesh(1.0, n_around=120, n_tube=60, tex_size=128), 128, 128, cuda.SAMPLER_COLOUR)


def test_converged_spectral_pyramid(ctx, oracle_mod):
    """Config 4 (dispersion, Beer-Lambert, spectral sampler; CIE XYZ before the epilogue) at 128 x 128, 4096 spp."""
    _converged(ctx, oracle_mod, "spectral_pyramid", scenes.spectral_pyramid(1.0), 128, 128, cuda.SAMPLER_SPECTRAL)


def test_converged_ibl(ctx, oracle_mod):
    """Config 5's materials and lighting (sky-dome image light, metal mesh, glass sphere) at 128 x 72, 4096 spp."""
    _converged(ctx, oracle_mod, "ibl", scenes.ibl_displaced_mesh(16 / 9, 120, 60, (256, 128)), 128, 72, cuda.SAMPLER_COLOUR)


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
    _same_path(ctx, oracle_mod, sc, 64, 36, 8, cuda.SAMPLER_COLOUR)


def test_spectral_pbr_image_albedo_same_path(ctx, oracle_mod):
    """Spectral PBR (pbr.go:158-263) with a SpectralImage albedo (spectral_image.go:61-245), the material
    transport builds for PBR in SPECTRAL scenes (transport.go:209-250)."""
    sc = scenes.spectral_pyramid(1.0)
    alb, rough, metal, nrm = scenes.pbr_textures(64)
    a = sc.image_texture(alb)
    pbr = sc.pbr(a, sc.image_texture(nrm), sc.image_texture(rough), sc.image_texture(metal), spectral_albedo=sc.spectral_image(a))
    tv, tuv = scenes.torus_mesh(60, 40, centre=(50.0, 60.0, 50.0), major=30.0, minor=12.0, amp=3.0, scale=0.5)
    sc.triangles(tv, pbr, tuv)
    _same_path(ctx, oracle_mod, sc, 40, 40, 16, cuda.SAMPLER_SPECTRAL)


@pytest.mark.parametrize("sampler", [cuda.SAMPLER_ALBEDO, cuda.SAMPLER_NORMAL], ids=["albedo", "normal"])
def test_aov_samplers_same_path(ctx, oracle_mod, sampler):
    """Debug AOV samplers (sampler/albedo.go, sampler/normal.go): first-hit albedo / normal, black on a miss."""
    for spec in (scenes.cornell_box(1.0), scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64)):
        img, ref = _same_path(ctx, oracle_mod, spec, 40, 40, 4, sampler)
        assert np.abs(img[1:, :, :3]).max() > 0


def test_worker_tile_rows(ctx, oracle_mod):
    """worker.RenderTile's wire shape (worker/render.go:17-75) against the oracle's restatement of it (same paths): rows in image
    order (no flip), strip_height padding, and the y == 0 row that the local path drops."""
    spec = scenes.cornell_box(1.0)
    # a narrower field of view than scenes.go:119-155, so that the bottom image row (y == 0) looks into the box, not under it
    spec.set_camera((278.0, 278.0, -800.0), (278, 278, 0), (0, 1, 0), 30.0, 1.0, 0.0, 10.0, 0.0, 1.0, 1.0)
    ctx.upload(cuda.HostScene(spec))
    osn = oracle_mod.OracleScene(spec)
    w = h = 50  # common.Tiles -> 25 x 25 tiles
    img, _ = ctx.render(w, h, 4, sampler=cuda.SAMPLER_COLOUR, seed=9)
    ctx.render_setup(w, h, 4, sampler=cuda.SAMPLER_COLOUR, seed=9)
    rows = ctx.render_tile_rows(25, 0, 49, 24, strip_height=1)
    want, _ = osn.render_tile(w, h, 4, 25, 0, 49, 24, strip_height=1, sampler=0, rng_mode=1, seed=9)
    assert rows.shape == want.shape == (25, 100)
    np.testing.assert_allclose(rows, want, rtol=SAME_PATH_RTOL, atol=1e-14)
    for r in range(1, 25):  # image row y sits at canvas row ny - y of the local render (rgb.go:41, remote.go:63-69)
        assert rows[r].tobytes() == img[h - r, 25:50].tobytes()
    assert (rows[0].reshape(-1, 4)[:, 3] == 1).all() and rows[0].reshape(-1, 4)[:, :3].max() > 0
    rows2 = ctx.render_tile_rows(0, 25, 24, 49, strip_height=3)
    want2, _ = osn.render_tile(w, h, 4, 0, 25, 24, 49, strip_height=3, sampler=0, rng_mode=1, seed=9)
    assert rows2.shape == want2.shape == (25, 300)
    assert (rows2[:, 100:] == 0).all() and (want2[:, 100:] == 0).all()  # make([]float64, stripSize): only the first row of the strip is filled
    np.testing.assert_allclose(rows2, want2, rtol=SAME_PATH_RTOL, atol=1e-14)
    # spectral worker: rows carry CIE XYZ means (worker/render.go:45-47), no epilogue
    spec = scenes.spectral_pyramid(1.0)
    ctx.upload(cuda.HostScene(spec))
    osn = oracle_mod.OracleScene(spec)
    ctx.render_setup(64, 64, 8, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
    rows3 = ctx.render_tile_rows(32, 32, 63, 63, strip_height=1)
    want3, _ = osn.render_tile(64, 64, 8, 32, 32, 63, 63, strip_height=1, sampler=1, rng_mode=1, seed=4)
    np.testing.assert_allclose(rows3, want3, rtol=SAME_PATH_RTOL, atol=1e-14)
    assert rows3.reshape(-1, 4)[:, :3].max() > 0
    with pytest.raises(cuda.IzpiError):
        ctx.render_tile_rows(0, 0, 24, 24, strip_height=0)
    with pytest.raises(cuda.IzpiError):
        ctx.render_tile_rows(0, 0, 64, 24)  # x1 outside the 64-pixel-wide image
