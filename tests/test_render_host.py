"""Host-side render logic that needs no GPU: the tile grid (common.Tiles, renderer.go:116,172-188), the
round-robin shard across ranks, and the one-reduce exchange step (world_size 2 over gloo on CPU)."""
import os
import socket

import numpy as np
import pytest

from izpi_b200 import render
from izpi_b200.build import build as build_lib


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build_lib()


@pytest.mark.parametrize("w,h,tile", [(400, 400, (25, 25)), (1024, 1024, (32, 32)), (3840, 2160, (32, 24)), (40, 60, (20, 20))])
def test_tile_list_covers_image_once(w, h, tile):
    t = render.tile_list(w, h)
    assert len(t) == (w // tile[0]) * (h // tile[1])
    assert ((t[:, 2] - t[:, 0] + 1) == tile[0]).all() and ((t[:, 3] - t[:, 1] + 1) == tile[1]).all()
    cover = np.zeros((h, w), dtype=np.int32)
    for x0, y0, x1, y1 in t[:: max(1, len(t) // 400)]:
        cover[y0:y1 + 1, x0:x1 + 1] += 1
    assert cover.max() == 1
    area = ((t[:, 2] - t[:, 0] + 1).astype(np.int64) * (t[:, 3] - t[:, 1] + 1)).sum()
    assert area == w * h
    # 4K case of BASELINE config 5: 10 800 tiles
    if (w, h) == (3840, 2160):
        assert len(t) == 10800


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_partition_tiles(world):
    t = render.tile_list(1024, 1024)
    parts = [render.shard_tiles(t, world, r) for r in range(world)]
    assert sum(len(p) for p in parts) == len(t)
    allrows = np.concatenate(parts)
    assert len(np.unique(allrows, axis=0)) == len(t)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


# TestWalkGridSpiral (internal/grid/grid_test.go:9-80): the reference's own expected paths for 3x3, 4x4 and 5x5 grids
_SPIRAL_GOLDEN = {
    (3, 3): [(1, 1), (1, 0), (2, 0), (2, 1), (2, 2), (1, 2), (0, 2), (0, 1), (0, 0)],
    (4, 4): [(2, 2), (2, 1), (3, 1), (3, 2), (3, 3), (2, 3), (1, 3), (1, 2), (1, 1), (1, 0), (2, 0), (3, 0), (0, 3), (0, 2), (0, 1), (0, 0)],
    (5, 5): [(2, 2), (2, 1), (3, 1), (3, 2), (3, 3), (2, 3), (1, 3), (1, 2), (1, 1), (1, 0), (2, 0), (3, 0), (4, 0), (4, 1), (4, 2), (4, 3),
             (4, 4), (3, 4), (2, 4), (1, 4), (0, 4), (0, 3), (0, 2), (0, 1), (0, 0)],
}


@pytest.mark.parametrize("size", sorted(_SPIRAL_GOLDEN))
def test_walk_grid_spiral_golden(size):
    assert [tuple(int(v) for v in p) for p in render.walk_grid_spiral(*size)] == _SPIRAL_GOLDEN[size]


def test_walk_grid_spiral_covers_any_grid():
    """Every cell once, centre first, also for non-square grids (cells outside the grid are walked but not emitted,
    grid.go:60-128); the tile list in that order is a permutation of the row-major one."""
    for gx, gy in ((120, 90), (32, 32), (16, 16), (1, 5), (4, 1)):
        p = render.walk_grid_spiral(gx, gy)
        assert len(p) == gx * gy and len({(int(a), int(b)) for a, b in p}) == gx * gy
        assert p[0].tolist() == [gx // 2, gy // 2]
        assert (p[:, 0] >= 0).all() and (p[:, 0] < gx).all() and (p[:, 1] >= 0).all() and (p[:, 1] < gy).all()
    t = render.spiral_tiles(3840, 2160)
    assert len(t) == 10800 and len(np.unique(t, axis=0)) == 10800
    assert {tuple(r) for r in t.tolist()} == {tuple(r) for r in render.tile_list(3840, 2160).tolist()}


def test_tiles_error_when_nothing_divides():
    from izpi_b200 import cuda
    with pytest.raises(cuda.IzpiError):
        render.tile_list(17, 16)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = h = 64
    tiles = render.tile_list(w, h)
    mine = render.shard_tiles(tiles, world, rank)
    # stand-in for izpi_render_tiles: every pixel of my tiles gets a value that depends only on (x, y)
    canvas = torch.zeros((h, w, 4), dtype=torch.float64)
    for x0, y0, x1, y1 in mine:
        ys, xs = torch.meshgrid(torch.arange(int(y0), int(y1) + 1), torch.arange(int(x0), int(x1) + 1), indexing="ij")
        canvas[int(y0):int(y1) + 1, int(x0):int(x1) + 1, 0] = (xs * 0.1 + ys * 7.3).double()
        canvas[int(y0):int(y1) + 1, int(x0):int(x1) + 1, 3] = 1.0
    render.reduce_canvas(canvas, dst=0)
    if rank == 0:
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        want = (xs * 0.1 + ys * 7.3).double()
        q.put(bool(torch.equal(canvas[..., 0], want) and bool((canvas[..., 3] == 1).all())))
    dist.destroy_process_group()


def test_reduce_canvas_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


# ---- dynamic dealing: the shared tile cursor (renderer.go:126-147: workers pull work units from one channel) -------------
def _claim_all(cursor_addr, n_tiles, tile_paths, batch_paths, takers):
    import ctypes as C
    from izpi_b200 import cuda
    L = cuda.lib()
    out = []
    b, e = C.c_int32(), C.c_int32()
    while L.izpi_host_claim_tiles(C.c_void_p(cursor_addr), n_tiles, tile_paths, batch_paths, takers, C.byref(b), C.byref(e)):
        out.append((b.value, e.value))
    return out


def test_claim_policy_single_taker_and_guided():
    cur = render.TileCursor()
    try:
        a = cur.begin_frame(0)
        assert _claim_all(a, 256, 40000, 1 << 24, 1) == [(0, 256)]  # a private cursor takes the list at once
        a = cur.begin_frame(0)
        claims = _claim_all(a, 10800, 768 * 1024, 1 << 24, 16)  # config 5: 786k paths per tile, 21 tiles fill a batch
        assert claims[0] == (0, 21) and claims[-1][1] == 10800
        assert all(c[1] - c[0] <= 21 for c in claims) and all(x[1] == y[0] for x, y in zip(claims, claims[1:]))
        sizes = [c[1] - c[0] for c in claims]
        assert sizes[-2] <= 4 and min(sizes[:-1]) >= 2  # runs shrink towards the end, never below 2^20 paths
    finally:
        cur.close()


def _claim_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for frame in range(3):  # several frames: the slots of the cursor are recycled without a barrier
        w, h = 3840, 2160
        tiles = render.tile_list(w, h)
        cur = render.shared_cursor()
        assert cur is not None
        addr = cur.begin_frame(rank)
        mine = _claim_all(addr, len(tiles), 768 * 64, 1 << 20, 2 * world)
        # stand-in for the render: mark every claimed tile once
        marks = torch.zeros(len(tiles), dtype=torch.float64)
        for b, e in mine:
            marks[b:e] += 1
        render.reduce_canvas(marks, dst=0)
        if rank == 0:
            ok = ok and bool((marks == 1).all())
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shared_cursor_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_claim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
