/*
 * izpi_host.h -- C++ host runtime above the device C ABI, exported with C linkage for harnesses.
 *
 * Stand-in for the Go packages that change in a real deployment (BASELINE.json north_star (1),(5)):
 * it does what the modified internal/hitable, internal/material, internal/texture, internal/camera
 * and internal/render would do on the host -- run the object constructors, build the BVH4
 * (hitable.NewBVH4, bvh4.go:517), flatten everything into izpi_scene_desc, schedule tiles -- and
 * then calls the izpi_cuda.h entry points.  The Go toolchain is absent from this image, so the
 * host side is C++ (the reference is compiled code); INTEGRATION.md maps each function to the Go
 * code that would replace it.
 */
#ifndef IZPI_HOST_H
#define IZPI_HOST_H

#include "izpi_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct izpi_host_scene izpi_host_scene;

/* transport.ToScene() (transport.go:53-92) / scenes.CornellBox (scenes.go:119-155): constructors,
 * lights = emitters, world = BVH4 or slice, camera.New.  `threads` parallelises the BVH build
 * (result is independent of it). */
int izpi_host_scene_create(const izpi_scene_spec* spec, int threads, izpi_host_scene** out);
void izpi_host_scene_destroy(izpi_host_scene* s);

/* BVH4.Nodes and the Primitives permutation (bvh4.go:42-47): Primitives[i] = hitables[perm[i]]. */
int32_t izpi_host_scene_num_nodes(const izpi_host_scene* s);
int izpi_host_scene_bvh(const izpi_host_scene* s, izpi_bvh4_node* nodes, int32_t* perm);
int32_t izpi_host_scene_num_lights(const izpi_host_scene* s);
int izpi_host_scene_lights(const izpi_host_scene* s, int32_t* orig_ids);

/* The flattened description (pointers stay valid until the host scene is destroyed). */
int izpi_host_scene_desc(const izpi_host_scene* s, izpi_scene_desc* out);
int izpi_host_scene_upload(const izpi_host_scene* s, izpi_ctx* ctx);

/* common.Tiles (common/tiles.go:6-24); 0 when no listed size divides the dimension. */
void izpi_host_tiles(int32_t size_x, int32_t size_y, int32_t* step_x, int32_t* step_y);

/* grid.WalkGrid(sizeX, sizeY, PATTERN_SPIRAL) (internal/grid/grid.go:27-128): the order in which RendererImpl.Render queues its
 * work units (renderer.go:151,172).  Start at the centre cell (sizeX/2, sizeY/2); at every step try the next direction of
 * (up, right, down, left) and keep the previous one while the cell that way has been walked already; cells outside the grid
 * are walked but not emitted.  xy receives size_x * size_y pairs {x, y}.  Pinned by the reference's own expected paths
 * (grid_test.go:9-80). */
void izpi_host_walk_grid_spiral(int32_t size_x, int32_t size_y, int32_t* xy);

/* The claim policy of a shared tile cursor (izpi_render_tiles_shared): guided self-scheduling.  A claimer with room for
 * one more batch takes (tiles left) / (2 * takers) tiles, at most one full batch (batch_paths / tile_paths tiles) and at
 * least 2^20 paths' worth (smaller batches are launch-bound), with one atomic fetch-add on *cursor; takers = contexts x
 * batches in flight per context.  takers <= 1 (a private cursor) claims everything at once.  Returns 0 when the list is
 * exhausted, else 1 with the claimed range [*begin, *end).  Pure host code: the cursor may sit in memory shared between
 * processes.  Mirrors the work-unit channel of renderer.go:126-147. */
int izpi_host_claim_tiles(uint64_t* cursor, int32_t n_tiles, int64_t tile_paths, int64_t batch_paths, int32_t takers,
                          int32_t* begin, int32_t* end);

/* RendererImpl.Render (renderer.go:108-222) for the local path on one context (one GPU or a device group): tile grid from
 * common.Tiles, queued in the reference's spiral order (izpi_host_walk_grid_spiral); the work units [tile_begin, tile_end)
 * of that queue are submitted, then izpi_render_finish when `finish` is set.  tile_end = -1 means all of them. */
int izpi_host_render(izpi_ctx* ctx, const izpi_render_config* cfg, int32_t tile_begin, int32_t tile_end,
                     int32_t finish, double* canvas_rgba, uint64_t* total_rays);

#ifdef __cplusplus
}
#endif
#endif /* IZPI_HOST_H */
