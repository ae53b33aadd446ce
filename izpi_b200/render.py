"""Host-side render loop: the stand-in for internal/render (renderer.go:73-222) on one box of GPUs.

`New(...).Render()` keeps the reference's call shape.  Work is sharded the way the reference farms
tiles to workers (renderer.go:126-147, remote.go:17-94) but across GPUs instead of goroutines/hosts:
one process per GPU (torchrun), the ranks CLAIM runs of tiles of the common.Tiles grid from one shared
cursor (TileCursor, the channel the reference's workers pull from), every rank renders its tiles into its own
zero-initialised fp64 canvas of running sums, and ONE `reduce(SUM)` of that canvas over NCCL (NVLink)
lands the image on rank 0 -- disjoint pixels, so the sum is exact and the 8-GPU image is bit-identical to
the 1-GPU image.  (A single process can also drive all GPUs of the box: cuda.Context([0, 1, ...]) is a device
group whose merge moves only the rendered tiles; see izpi_ctx_create.)  The spectral epilogue
(FireflyRejection needs neighbours across tile edges, renderer.go:216-219) then runs on rank 0.

The tiles are queued in the reference's spiral order from the centre (grid.WalkGrid, grid.go:27; `walk_grid_spiral`);
every order produces the same canvas, this one also ends a multi-GPU frame evenly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import cuda
from .scene import SceneSpec

ColourSampler, SpectralSampler = cuda.SAMPLER_COLOUR, cuda.SAMPLER_SPECTRAL


def tile_list(size_x: int, size_y: int) -> np.ndarray:
    """workUnit bounds for every tile of the common.Tiles grid (renderer.go:116,172-188): (n, 4) uint32
    rows {x0, y0, x1, y1}, inclusive, row-major."""
    sx, sy = cuda.tiles(size_x, size_y)
    if sx == 0 or sy == 0:
        raise cuda.IzpiError(cuda.EINVAL, "no tile size divides the image dimensions (common.Tiles)")
    gx, gy = size_x // sx, size_y // sy
    ty, tx = np.meshgrid(np.arange(gy, dtype=np.uint32), np.arange(gx, dtype=np.uint32), indexing="ij")
    tx, ty = tx.ravel(), ty.ravel()
    return np.stack([tx * sx, ty * sy, tx * sx + (sx - 1), ty * sy + (sy - 1)], axis=1).astype(np.uint32)


def walk_grid_spiral(size_x: int, size_y: int) -> np.ndarray:
    """grid.WalkGrid(sizeX, sizeY, PATTERN_SPIRAL) (internal/grid/grid.go:27-128), the order in which RendererImpl.Render queues
    its work units (renderer.go:151,172): start at the centre cell, try to turn at every step (up, right, down, left, ...),
    keep going straight while the cell in the new direction has been walked already; cells outside the grid are walked but not
    emitted.  Returns (n, 2) int32 grid positions."""
    total = size_x * size_y
    if total <= 0:
        return np.zeros((0, 2), dtype=np.int32)
    out = np.zeros((total, 2), dtype=np.int32)
    cuda.lib().izpi_host_walk_grid_spiral(size_x, size_y, out.ctypes.data)  # csrc/host/host_scene.cpp
    return out


def spiral_tiles(size_x: int, size_y: int) -> np.ndarray:
    """tile_list() in the reference's queueing order (walk_grid_spiral over the tile grid): centre of the image first, border
    last.  Pixels are disjoint, so the canvas does not depend on the order; the shared cursor deals runs of THIS list, so the
    large early claims are the (usually expensive) middle of the frame and the short final claims its (usually cheap) rim."""
    sx, sy = cuda.tiles(size_x, size_y)
    if sx == 0 or sy == 0:
        raise cuda.IzpiError(cuda.EINVAL, "no tile size divides the image dimensions (common.Tiles)")
    g = walk_grid_spiral(size_x // sx, size_y // sy).astype(np.uint32)
    tx, ty = g[:, 0], g[:, 1]
    return np.ascontiguousarray(np.stack([tx * sx, ty * sy, tx * sx + (sx - 1), ty * sy + (sy - 1)], axis=1).astype(np.uint32))


def shard_tiles(tiles: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Static round-robin deal of the tile list (rank r renders tiles r, r+world, ...): the fallback when the ranks have no
    shared cursor.  The default is dynamic dealing, see TileCursor."""
    return np.ascontiguousarray(tiles[rank::world])


class TileCursor:
    """The shared work queue of a multi-process render: the index of the next unclaimed tile, in POSIX shared memory,
    advanced with atomic fetch-adds inside izpi_render_tiles_shared.  The reference's workers pull work units from one
    channel (renderer.go:126-147); one process per GPU pulls tile runs from this cursor the same way, so no rank is
    stuck with a fixed share of expensive tiles.

    Four slots, used round-robin by frame: rank 0 clears slot (k+1) % 4 when it starts frame k.  Ranks are never more
    than one frame apart (every frame ends with a collective), so a slot is cleared well before anyone claims from it
    again and no extra barrier is needed."""

    SLOTS = 4

    def __init__(self, name: str | None = None):
        from multiprocessing import shared_memory
        self.owner = name is None
        self.shm = shared_memory.SharedMemory(create=True, size=4096) if self.owner else shared_memory.SharedMemory(name=name)
        if not self.owner:  # Python < 3.13 also registers attachments with the resource tracker, which would unlink the owner's segment
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.words = np.frombuffer(self.shm.buf, dtype=np.uint64, count=self.SLOTS)
        if self.owner:
            self.words[:] = 0
        self.frame = 0

    @property
    def name(self) -> str:
        return self.shm.name

    def begin_frame(self, rank: int) -> int:
        """Address of this frame's cursor word."""
        k = self.frame
        self.frame += 1
        if rank == 0:
            self.words[(k + 1) % self.SLOTS] = 0
        return self.words.ctypes.data + 8 * (k % self.SLOTS)

    def close(self):
        self.words = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


_cursor = None


def shared_cursor():
    """One TileCursor per process group: rank 0 creates it, the name travels by broadcast.  None when not distributed or
    when shared memory is unavailable (then tiles are dealt statically)."""
    global _cursor
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if _cursor is None:
        box = [None]
        if dist.get_rank() == 0:
            try:
                _cursor = TileCursor()
                box[0] = _cursor.name
            except Exception:
                box[0] = ""
        dist.broadcast_object_list(box, src=0)
        if dist.get_rank() != 0 and box[0]:
            try:
                _cursor = TileCursor(box[0])
            except Exception:
                _cursor = False
        ok = [bool(_cursor)]
        oks = [None] * dist.get_world_size()
        dist.all_gather_object(oks, ok[0])
        if not all(oks):
            if _cursor:
                _cursor.close()
            _cursor = False
        if _cursor:
            import atexit
            atexit.register(_cursor.close)
    return _cursor or None


def replicate_scene(ctx: cuda.Context, src: int = 0):
    """One scene build per box: rank `src` has uploaded the scene; every other rank allocates the same device blocks and
    receives their contents over NCCL (NVLink), then re-bases the pointer tables (izpi_scene_image_*).  The reference
    re-runs ToScene() + NewBVH4 on every LAN worker (transport.go:53-92); on one box that is N-1 redundant builds."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    rank = dist.get_rank()
    box = [None]
    if rank == src:
        header, sizes, ptrs = ctx.scene_image()
        box[0] = (header.tobytes(), sizes)
    dist.broadcast_object_list(box, src=src)
    header_bytes, sizes = box[0]
    if rank != src:
        ptrs = ctx.scene_adopt(np.frombuffer(header_bytes, dtype=np.uint8), len(sizes))
    dev = torch.device("cuda", ctx.device) if torch.cuda.is_available() else None
    for p, n in zip(ptrs, sizes):
        if n == 0:
            continue
        t = torch.as_tensor(_DevicePtr(p, (n,), "|u1"), device=dev)
        dist.broadcast(t, src=src)
    torch.cuda.synchronize()
    if rank != src:
        ctx.scene_commit()


def reduce_canvas(canvas, dst: int = 0):
    """The one exchange step of a multi-process render: SUM-reduce the fp64 canvas of running sums to `dst`.  Every rank's
    canvas is zero outside the tiles it rendered, so the sum is exact (x + 0) and the image is bit-identical to the
    single-GPU one.  `canvas` is a torch tensor (CUDA + NCCL in production, CPU + gloo in the host-logic tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(canvas, dst=dst, op=dist.ReduceOp.SUM)
    return canvas


class _DevicePtr:
    """Zero-copy view of device memory owned by libizpi_cuda.so, for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def host_canvas(height: int, width: int) -> np.ndarray:
    """A float64 RGBA canvas (floatimage.Float64NRGBA.Pix) in page-locked host memory, so that the device-to-host copy
    of a frame (265 MB at 4K) runs at PCIe speed.  Pageable when torch or CUDA is absent."""
    try:
        import torch
        t = torch.empty((height, width, 4), dtype=torch.float64, pin_memory=torch.cuda.is_available())
        a = t.numpy()
        _keep_alive[id(a)] = t
        return a
    except ImportError:
        return np.empty((height, width, 4), dtype=np.float64)


_keep_alive = {}


class Renderer:
    """render.RendererImpl (renderer.go:26-44).  The canvas belongs to the renderer (render.New allocates it,
    renderer.go:89): Render() returns the same array every time it is called on ONE renderer, and two renderers never
    share one."""

    def __init__(self, ctx: cuda.Context, size_x, size_y, num_samples, max_depth=50, background=(0.0, 0.0, 0.0),
                 spectral_background=None, sampler_type=ColourSampler, seed=1, stats=False, dynamic=True, sample_count=None):
        self.ctx, self.size_x, self.size_y = ctx, size_x, size_y
        self.sample_count = num_samples if sample_count is None else sample_count  # < num_samples: a partial (e.g. warm-up) pass of the frame
        self.num_samples, self.max_depth, self.sampler_type, self.seed = num_samples, max_depth, sampler_type, seed
        self.background, self.spectral_background = background, spectral_background
        # work-unit order: the reference's spiral from the centre (grid.WalkGrid; renderer.go:151) or row-major ("linear").  With
        # the spiral the large early claims of a multi-GPU frame are its middle and the short last ones its rim: the 8-GPU
        # config-5 frame ends with a 2 ms wait at the merge instead of 337 ms (5.33 s against 5.73 s).
        self.walk = os.environ.get("IZPI_TILE_WALK", "spiral")
        self.stats, self.dynamic = int(stats), dynamic  # stats: 0 | cuda.RENDER_STATS | cuda.RENDER_TIMING (measurement modes)
        self.num_rays = 0
        self.timings = {}
        self._canvas = None

    def _config(self):
        cfg = cuda.RenderConfig(width=self.size_x, height=self.size_y, spp=self.num_samples, max_depth=self.max_depth,
                                sampler=self.sampler_type, sample_offset=0, sample_count=self.sample_count, seed=self.seed,
                                flags=int(self.stats))
        cfg.background[:] = [float(c) for c in self.background]
        self._keep = None
        if self.spectral_background is not None:  # control.proto:64-67
            w = np.ascontiguousarray(self.spectral_background[0], dtype=np.float64)
            v = np.ascontiguousarray(self.spectral_background[1], dtype=np.float64)
            cfg.bg_wavelengths, cfg.bg_values, cfg.n_bg = w.ctypes.data, v.ctypes.data, len(w)
            self._keep = (w, v)
        return cfg

    def canvas(self):
        if self._canvas is None:
            self._canvas = host_canvas(self.size_y, self.size_x)
        return self._canvas

    def Render(self):
        """Returns the float64 RGBA canvas on rank 0 (None on other ranks) -- (*RendererImpl).Render, renderer.go:108.
        self.timings holds the wall-clock split of the frame in ms (setup / tiles / merge / finish+D2H)."""
        import time
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        L, h = cuda.lib(), self.ctx._h
        if rank == 0:
            self.canvas()  # page-locked allocation happens once per renderer, not per frame
        t0 = time.perf_counter()
        cfg = self._config()
        cuda.check(L.izpi_render_setup(h, C.byref(cfg)))
        t1 = time.perf_counter()
        tiles = spiral_tiles(self.size_x, self.size_y) if self.walk == "spiral" else tile_list(self.size_x, self.size_y)
        cursor = shared_cursor() if (world > 1 and self.dynamic) else None
        if world > 1 and cursor is None:
            mine = shard_tiles(tiles, world, rank)
            cuda.check(L.izpi_render_tiles(h, len(mine), mine.ctypes.data, None))
        elif world > 1:
            cuda.check(L.izpi_render_tiles_shared(h, len(tiles), tiles.ctypes.data, cursor.begin_frame(rank), world, None))
        else:
            cuda.check(L.izpi_render_tiles(h, len(tiles), tiles.ctypes.data, None))
        t2 = time.perf_counter()
        rays = C.c_uint64()
        if world > 1:
            import torch
            import torch.distributed as dist
            view = torch.as_tensor(_DevicePtr(self.ctx.canvas_device_ptr(), (self.size_y, self.size_x, 4)), device="cuda")
            reduce_canvas(view, dst=0)
            cuda.check(L.izpi_render_finish(h, None, C.byref(rays)))
            t = torch.tensor([rays.value], dtype=torch.int64, device="cuda")
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)  # RenderEnd's total_rays_traced (renderer.go:201-211)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            if rank != 0:
                self.timings = {"setup_ms": (t1 - t0) * 1e3, "tiles_ms": (t2 - t1) * 1e3, "merge_ms": (t3 - t2) * 1e3, "finish_ms": 0.0}
                return None
            self.num_rays = int(t.item())
            canvas = self.canvas()
            cuda.check(L.izpi_render_finish(h, canvas.ctypes.data, None))
            t4 = time.perf_counter()
            self.timings = {"setup_ms": (t1 - t0) * 1e3, "tiles_ms": (t2 - t1) * 1e3, "merge_ms": (t3 - t2) * 1e3, "finish_ms": (t4 - t3) * 1e3}
            return canvas
        canvas = self.canvas()
        cuda.check(L.izpi_render_finish(h, canvas.ctypes.data, C.byref(rays)))
        t3 = time.perf_counter()
        self.num_rays = rays.value
        self.timings = {"setup_ms": (t1 - t0) * 1e3, "tiles_ms": (t2 - t1) * 1e3, "merge_ms": 0.0, "finish_ms": (t3 - t2) * 1e3}
        return canvas


def New(ctx: cuda.Context, size_x, size_y, num_samples, max_depth=50, background=(0.0, 0.0, 0.0), spectral_background=None,
        sampler_type=ColourSampler, seed=1, stats=False, dynamic=True, sample_count=None) -> Renderer:
    """render.New (renderer.go:73-106); the scene is the one uploaded to `ctx`."""
    return Renderer(ctx, size_x, size_y, num_samples, max_depth, background, spectral_background, sampler_type, seed, stats, dynamic, sample_count)
