"""Host runtime (product) vs oracle: the two independent BVH4 builders must produce byte-identical
node arrays and permutations (bvh4.go:517-855), and the same light lists (transport.go:67-72)."""
import numpy as np
import pytest

from izpi_b200 import cuda, scenes
from izpi_b200 import scene as S
from izpi_b200.build import build as build_lib


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build_lib()


def _compare(spec, oracle_mod, threads=None):
    hs = cuda.HostScene(spec, threads=threads)
    osn = oracle_mod.OracleScene(spec)
    hn, hp = hs.bvh()
    on, op = osn.bvh()
    assert len(hn) == len(on)
    assert hn.tobytes() == on.tobytes()
    np.testing.assert_array_equal(hp, op)
    np.testing.assert_array_equal(hs.lights(), osn.lights())
    return hn, hp


@pytest.mark.parametrize("shape", [(8, 4), (40, 25), (125, 80)])
@pytest.mark.parametrize("seed", [12345, 7, 2**40 + 3])
def test_bvh_torus(oracle_mod, shape, seed):
    verts, uvs = scenes.torus_mesh(*shape)
    sc = S.SceneSpec(bvh_seed=seed)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    _compare(sc, oracle_mod)


def test_bvh_thread_count_independent(oracle_mod):
    verts, uvs = scenes.torus_mesh(300, 200)  # 120k triangles: exercises the parallel split (> 32768)
    sc = S.SceneSpec(bvh_seed=99)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    a, _ = _compare(sc, oracle_mod, threads=1)
    b, _ = _compare(sc, oracle_mod, threads=8)
    assert a.tobytes() == b.tobytes()


def test_bvh_duplicate_keys(oracle_mod):
    """Many equal box.min keys: exercises the equal-key paths of the sort restatement."""
    g = np.stack(np.meshgrid(np.arange(30), np.arange(30), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    verts = np.stack([g, g + [1, 0, 0], g + [0, 1, 0]], axis=1)
    for seed in (1, 2, 3):
        sc = S.SceneSpec(bvh_seed=seed)
        sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))))
        _compare(sc, oracle_mod)


def test_bvh_soup_and_spheres(oracle_mod):
    sc = S.SceneSpec(bvh_seed=5)
    m = sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5)))
    sc.triangles(scenes.triangle_soup(20000), m)
    for i in range(50):
        sc.sphere((i * 2.0, 10.0, 5.0), 0.75, m)
    _compare(sc, oracle_mod)


def test_bvh_fixture_scenes(oracle_mod):
    _compare(scenes.spectral_pyramid(), oracle_mod)
    n, _ = _compare(scenes.cornell_box(), oracle_mod)
    assert len(n) == 0  # config 1 is a plain HitableSlice (scenes.go:153)
    sc = S.SceneSpec(bvh_rand_zero=True)
    m = sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5)))
    for c in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)]:
        sc.sphere(c, 1.0, m)
    n, _ = _compare(sc, oracle_mod)
    assert len(n) == 3


def test_bvh_wrapped_prims(oracle_mod):
    """Translate / RotateY / FlipNormals bounding boxes inside a BVH4."""
    sc = scenes.cornell_box()
    sc.world_kind = S.WORLD_BVH4
    sc.bvh_seed = 3
    _compare(sc, oracle_mod)


def test_empty_scene_rejected_cleanly(oracle_mod):
    sc = S.SceneSpec()
    hs = cuda.HostScene(sc)  # bvh4.go:559-562: empty hitables -> no BVH
    n, p = hs.bvh()
    assert len(n) == 0 and len(p) == 0


def test_tiles():
    assert cuda.tiles(400, 400) == (25, 25)
    assert cuda.tiles(1024, 1024) == (32, 32)
    assert cuda.tiles(3840, 2160) == (32, 24)
    assert cuda.tiles(7, 400) == (0, 25)  # the reference divides by zero here (renderer.go:117)
