"""Host-side render loop: the stand-in for internal/render (renderer.go:73-222) on one box of GPUs.

`New(...).Render()` keeps the reference's call shape.  Work is sharded the way the reference farms
tiles to workers (renderer.go:126-147, remote.go:17-94) but across GPUs instead of goroutines/hosts:
one process per GPU (torchrun), the common.Tiles grid is dealt round-robin to the ranks, every rank
renders its tiles into its own zero-initialised fp64 canvas of running sums, and ONE
`reduce(SUM)` of that canvas over NCCL (NVLink) lands the image on rank 0 -- disjoint pixels, so the
sum is exact and the 8-GPU image is bit-identical to the 1-GPU image.  The spectral epilogue
(FireflyRejection needs neighbours across tile edges, renderer.go:216-219) then runs on rank 0.

The tile walk order is linear instead of the reference's spiral (grid.WalkGrid, grid.go:27): the
spiral only serves the live preview; every order produces the same canvas.
"""
from __future__ import annotations

import ctypes as C

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


def shard_tiles(tiles: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Round-robin deal of the tile list: rank r renders tiles r, r+world, r+2*world, ..."""
    return np.ascontiguousarray(tiles[rank::world])


def reduce_canvas(canvas, dst: int = 0):
    """The one exchange step of a multi-GPU render: SUM-reduce the fp64 canvas of running sums to `dst`.
    `canvas` is a torch tensor (CUDA + NCCL in production, CPU + gloo in the host-logic tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(canvas, dst=dst, op=dist.ReduceOp.SUM)
    return canvas


class _DevicePtr:
    """Zero-copy view of device memory owned by libizpi_cuda.so, for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


_pinned = {}


def host_canvas(height: int, width: int) -> np.ndarray:
    """The float64 RGBA canvas (floatimage.Float64NRGBA.Pix) in page-locked host memory, so that the one
    device-to-host copy of a frame (265 MB at 4K) runs at PCIe speed.  Falls back to pageable memory
    when torch is absent."""
    try:
        import torch
        key = (height, width)
        if key not in _pinned:
            _pinned.clear()
            _pinned[key] = torch.empty((height, width, 4), dtype=torch.float64, pin_memory=torch.cuda.is_available())
        return _pinned[key].numpy()
    except ImportError:
        return np.empty((height, width, 4), dtype=np.float64)


class Renderer:
    """render.RendererImpl (renderer.go:26-44)."""

    def __init__(self, ctx: cuda.Context, size_x, size_y, num_samples, max_depth=50, background=(0.0, 0.0, 0.0),
                 spectral_background=None, sampler_type=ColourSampler, seed=1):
        self.ctx, self.size_x, self.size_y = ctx, size_x, size_y
        self.num_samples, self.max_depth, self.sampler_type, self.seed = num_samples, max_depth, sampler_type, seed
        self.background, self.spectral_background = background, spectral_background
        self.num_rays = 0

    def _config(self):
        cfg = cuda.RenderConfig(width=self.size_x, height=self.size_y, spp=self.num_samples, max_depth=self.max_depth,
                                sampler=self.sampler_type, sample_offset=0, sample_count=self.num_samples, seed=self.seed)
        cfg.background[:] = [float(c) for c in self.background]
        self._keep = None
        if self.spectral_background is not None:  # control.proto:64-67
            w = np.ascontiguousarray(self.spectral_background[0], dtype=np.float64)
            v = np.ascontiguousarray(self.spectral_background[1], dtype=np.float64)
            cfg.bg_wavelengths, cfg.bg_values, cfg.n_bg = w.ctypes.data, v.ctypes.data, len(w)
            self._keep = (w, v)
        return cfg

    def Render(self):
        """Returns the float64 RGBA canvas on rank 0 (None on other ranks) -- (*RendererImpl).Render, renderer.go:108."""
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        L, h = cuda.lib(), self.ctx._h
        cfg = self._config()
        cuda.check(L.izpi_render_setup(h, C.byref(cfg)))
        mine = shard_tiles(tile_list(self.size_x, self.size_y), world, rank)
        cuda.check(L.izpi_render_tiles(h, len(mine), mine.ctypes.data, None))
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
            if rank != 0:
                return None
            self.num_rays = int(t.item())
            canvas = host_canvas(self.size_y, self.size_x)
            cuda.check(L.izpi_render_finish(h, canvas.ctypes.data, None))
            return canvas
        canvas = host_canvas(self.size_y, self.size_x)
        cuda.check(L.izpi_render_finish(h, canvas.ctypes.data, C.byref(rays)))
        self.num_rays = rays.value
        return canvas


def New(ctx: cuda.Context, size_x, size_y, num_samples, max_depth=50, background=(0.0, 0.0, 0.0), spectral_background=None,
        sampler_type=ColourSampler, seed=1) -> Renderer:
    """render.New (renderer.go:73-106); the scene is the one uploaded to `ctx`."""
    return Renderer(ctx, size_x, size_y, num_samples, max_depth, background, spectral_background, sampler_type, seed)
