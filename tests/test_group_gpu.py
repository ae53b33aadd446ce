"""Scale-out behind the C ABI: device groups (izpi_ctx_create with n_devices > 1), scene images (one scene build per box)
and the shared tile cursor (dynamic dealing).  Everything is checked against the single-device result of the same calls,
which tests/test_trace_gpu.py and tests/test_render_gpu.py check against the oracle.

The single-GPU tests exercise the same code paths with two contexts on ONE device (replication through a device-to-device
copy, two host threads pulling tiles from one cursor); the group tests proper need two GPUs (`gpurun --gpus 2`)."""
import ctypes as C
import threading

import numpy as np
import pytest

from izpi_b200 import cuda, render, scenes
from izpi_b200 import scene as S

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def ctx():
    from izpi_b200.build import build
    build()
    c = cuda.Context(0)
    yield c
    c.close()


def _small_mesh_scene():
    verts, uvs = scenes.torus_mesh(120, 60)
    sc = S.SceneSpec(bvh_seed=12345)
    sc.triangles(verts, sc.lambertian(sc.constant_texture((0.5, 0.5, 0.5))), uvs)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    return sc, lo, hi


def _copy_blocks(dst_ptrs, src_ptrs, sizes):
    import torch
    for d, s, n in zip(dst_ptrs, src_ptrs, sizes):
        if n:
            torch.as_tensor(render._DevicePtr(d, (n,), "|u1"), device="cuda").copy_(torch.as_tensor(render._DevicePtr(s, (n,), "|u1"), device="cuda"))
    torch.cuda.synchronize()


def test_scene_image_roundtrip_same_device(ctx, oracle_mod):
    """Export / adopt / commit: a second context receives the uploaded scene block by block (no host rebuild) and answers
    every closest-hit query and renders every pixel exactly like the exporter."""
    spec = scenes.cornell_pbr_mesh(1.0, n_around=60, n_tube=40, tex_size=64)  # BVH4 + image textures (nested device pointers)
    ctx.upload(cuda.HostScene(spec))
    header, sizes, ptrs = ctx.scene_image()
    other = cuda.Context(0)
    try:
        mine = other.scene_adopt(header, len(sizes))
        assert len(mine) == len(ptrs) and not set(mine) & set(ptrs)
        _copy_blocks(mine, ptrs, sizes)
        other.scene_commit()
        a, ra = ctx.render(40, 40, 4, seed=3)
        b, rb = other.render(40, 40, 4, seed=3)
        assert a.tobytes() == b.tobytes() and ra == rb
        # spectral scene: tabulated SPD tables hang off the spectral texture table
        spec2 = scenes.spectral_pyramid(1.0)
        ctx.upload(cuda.HostScene(spec2))
        header, sizes, ptrs = ctx.scene_image()
        mine = other.scene_adopt(header, len(sizes))
        _copy_blocks(mine, ptrs, sizes)
        other.scene_commit()
        a, _ = ctx.render(32, 32, 8, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
        b, _ = other.render(32, 32, 8, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
        assert a.tobytes() == b.tobytes() and np.abs(a[..., :3]).max() > 0
        with pytest.raises(cuda.IzpiError):
            other.scene_adopt(header[:-8], len(sizes))  # truncated header
    finally:
        other.close()


def test_shared_cursor_two_contexts_one_device(ctx):
    """izpi_render_tiles_shared: two contexts pull tile runs from one cursor; their canvases of sums are disjoint and add up
    to the single-context frame bit for bit, whatever the interleaving."""
    spec = scenes.cornell_box(1.0)
    hs = cuda.HostScene(spec)
    ctx.upload(hs)
    w = h = 400  # common.Tiles -> 25 x 25 tiles, 256 tiles
    spp = 64
    full, rays = ctx.render(w, h, spp, seed=21)
    other = cuda.Context(0)
    try:
        other.upload(hs)
        tiles = render.tile_list(w, h)
        cursor = np.zeros(1, dtype=np.uint64)
        L = cuda.lib()
        out = {}
        ready = threading.Barrier(2)  # the first setup of a context allocates its buffers: start pulling tiles together

        def work(name, c):
            cfg = cuda.RenderConfig(width=w, height=h, spp=spp, max_depth=50, sampler=cuda.SAMPLER_COLOUR, sample_offset=0, sample_count=spp, seed=21)
            cuda.check(L.izpi_render_setup(c._h, C.byref(cfg)))
            ready.wait()
            cuda.check(L.izpi_render_tiles_shared(c._h, len(tiles), tiles.ctypes.data, cursor.ctypes.data, 2, None))
            out[name] = c.render_finish(w, h)

        th = [threading.Thread(target=work, args=(n, c)) for n, c in (("a", ctx), ("b", other))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        (a, ra), (b, rb) = out["a"], out["b"]
        assert cursor[0] >= len(tiles)
        wa, wb = a[..., 3] == 1, b[..., 3] == 1
        assert not (wa & wb).any() and (wa | wb)[1:].all()
        assert wa.any() and wb.any(), "one context claimed every tile: the cursor was not shared"
        np.testing.assert_array_equal(np.where(wa[..., None], a, b), full)
        assert ra + rb == rays
    finally:
        other.close()


def test_render_stats_counts_match_trace_counts(ctx):
    """IZPI_RENDER_STATS: the counting extend kernels count like the counting trace kernel (whose counts equal the
    oracle's, tests/test_trace_gpu.py): the primary rays of a 1-spp, depth-0 frame re-traced as a ray batch."""
    sc = scenes.cornell_pbr_mesh(1.0, n_around=100, n_tube=60, tex_size=32)
    ctx.upload(cuda.HostScene(sc))
    r = render.New(ctx, 64, 64, 2, 50, sampler_type=cuda.SAMPLER_COLOUR, seed=5, stats=True)
    img = r.Render().copy()
    st = ctx.render_stats()
    assert st["rays"] == r.num_rays > 64 * 64 * 2
    assert st["nodes_visited"] > st["rays"] and st["prim_tests"] > 0 and st["extend_launches"] >= 2
    assert st["extend_ms"] > 0 and st["shade_ms"] > 0
    plain = render.New(ctx, 64, 64, 2, 50, sampler_type=cuda.SAMPLER_COLOUR, seed=5).Render()
    assert plain.tobytes() == img.tobytes()  # measurement mode does not change the image


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_group_trace_and_render_match_single_device(ctx, oracle_mod):
    grp = cuda.Context([0, 1])
    try:
        assert cuda.lib().izpi_ctx_num_devices(grp._h) == 2
        sc, lo, hi = _small_mesh_scene()
        hs = cuda.HostScene(sc)
        ctx.upload(hs)
        grp.upload(hs)  # flattened once, second device filled by cudaMemcpyPeerAsync
        org, d = scenes.random_rays(100_003, lo, hi)
        i1, t1 = ctx.trace_closest(org, d)
        i2, t2, st = grp.trace_closest(org, d, stats=True)
        assert np.array_equal(i1, i2) and t1.tobytes() == t2.tobytes()
        oi, ot = oracle_mod.OracleScene(sc).trace(org, d)
        assert np.array_equal(i2, oi) and t2.tobytes() == ot.tobytes()
        assert st["rays"] == len(org)
        # RGB frame: tiles dealt dynamically, every member writes its runs straight into the caller's canvas
        spec = scenes.cornell_box(1.0)
        hs = cuda.HostScene(spec)
        ctx.upload(hs)
        grp.upload(hs)
        a, ra = ctx.render(200, 200, 32, seed=8)
        b, rb = grp.render(200, 200, 32, seed=8)
        assert a.tobytes() == b.tobytes() and ra == rb
        # partial frame (tile subset): unrendered pixels stay zero
        a, _ = ctx.render(200, 200, 8, seed=8, tile_begin=3, tile_end=40)
        b, _ = grp.render(200, 200, 8, seed=8, tile_begin=3, tile_end=40)
        assert a.tobytes() == b.tobytes()
        # spectral frame: the rendered runs gather on the first device for FireflyRejection
        spec = scenes.spectral_pyramid(1.0)
        hs = cuda.HostScene(spec)
        ctx.upload(hs)
        grp.upload(hs)
        a, _ = ctx.render(128, 128, 16, sampler=cuda.SAMPLER_SPECTRAL, seed=2)
        b, _ = grp.render(128, 128, 16, sampler=cuda.SAMPLER_SPECTRAL, seed=2)
        assert a.tobytes() == b.tobytes() and np.abs(a[..., :3]).max() > 0
    finally:
        grp.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_second_device_alone_renders_spectral(oracle_mod):
    """the CIE tables live in __constant__ memory, one copy per device; a thread that first rendered on device 0
    must still get a non-black spectral frame on device 1."""
    c0, c1 = cuda.Context(0), cuda.Context(1)
    try:
        hs = cuda.HostScene(scenes.spectral_pyramid(1.0))
        c0.upload(hs)
        c1.upload(hs)
        a, _ = c0.render(32, 32, 8, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
        b, _ = c1.render(32, 32, 8, sampler=cuda.SAMPLER_SPECTRAL, seed=4)
        assert np.abs(a[..., :3]).max() > 0 and a.tobytes() == b.tobytes()
    finally:
        c0.close()
        c1.close()
