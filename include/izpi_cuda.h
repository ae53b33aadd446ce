/*
 * izpi_cuda.h -- C ABI of libizpi_cuda.so, the B200 (sm_100a) backend for Izpi's hot path.
 *
 * This is what a cgo package `internal/cuda` binds (INTEGRATION.md shows the stub).  The
 * reference has no FFI seam; each entry point replaces one of its Go seams:
 *
 *   izpi_scene_upload    <- the scene.Scene object graph produced by transport.ToScene()
 *                           (internal/transport/transport.go:53-92): BVH4.Nodes / BVH4.Primitives
 *                           (internal/hitable/bvh4.go:42-47) flattened by package hitable,
 *                           material/texture tables flattened by packages material/texture.
 *   izpi_trace_closest   <- hitable.Hitable.Hit(r, tMin, tMax) (internal/hitable/api.go:15) as
 *                           implemented by *BVH4 (bvh4.go:49-164) under HitableSlice.Hit
 *                           (hitable_slice.go:30-45), batched (a cgo call per ray costs ~35 ns,
 *                           bvh4_simd_arm64.go:14, so the boundary is per batch).
 *   izpi_render_setup    <- render.New(...) (internal/render/renderer.go:73-106) and
 *                           RenderSetup (internal/proto/control/control.proto:56-68).
 *   izpi_render_tiles    <- renderRectRGB / renderRectSpectral (render/rgb.go:12,
 *                           render/spectral.go:14) and worker.RenderTile
 *                           (internal/worker/render.go:17-75), for a batch of tiles.
 *   izpi_render_finish   <- the tail of RendererImpl.Render (renderer.go:213-221):
 *                           FireflyRejection + XYZToRGB for the spectral sampler, ray total.
 *
 * Conventions: every function returns 0 on success and a negative IZPI_E* code on failure;
 * izpi_last_error() returns a thread-local message (the Go side wraps it into `error`).  All
 * pointers are borrowed for the duration of the call (cgo rule); uploads copy; outputs are
 * caller-allocated host memory unless the name says `_device`.  Calls on different contexts are
 * thread-safe; calls on one context must be serialised by the caller (one goroutine per GPU,
 * runtime.LockOSThread).  There is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef IZPI_CUDA_H
#define IZPI_CUDA_H

#include <stdint.h>
#include "izpi_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IZPI_OK 0
#define IZPI_EINVAL (-1)  /* bad argument                         */
#define IZPI_ECUDA (-2)   /* CUDA runtime error / no device       */
#define IZPI_ESTATE (-3)  /* call out of order (no scene, no setup) */

typedef struct izpi_ctx izpi_ctx;

const char* izpi_last_error(void);
int izpi_version(void);

/* A context drives one GPU or a GROUP of GPUs of one box (n_devices > 1, distinct device_ids; NULL = devices 0..n-1).
 * The reference starts its N workers from ONE process and hands them work units from one channel
 * (internal/render/renderer.go:126-147); a group does the same with one host thread per GPU inside the library:
 *   izpi_scene_upload   flattens / validates once, uploads to the first device and copies every array to the other
 *                       members device-to-device (cudaMemcpyPeerAsync, NVLink when the pair are peers);
 *   izpi_trace_closest  cuts the ray batch into contiguous slices, one per member, answers land in disjoint host ranges;
 *   izpi_render_tiles   deals the tiles DYNAMICALLY: every member claims the next run of tiles from one shared cursor
 *                       whenever it has room for another batch (no static assignment, sky and mesh tiles balance out);
 *   izpi_render_finish  merges: pixels are disjoint, so every member sends only the tiles it rendered (RGB / AOV samplers:
 *                       straight into the caller's canvas, N device-to-host streams in parallel; spectral: to the first
 *                       device, whose FireflyRejection needs the neighbours) -- no full-canvas reduction.
 * Every other entry point (displacement, device BVH build, tile rows, diagnostics) runs on the first device.
 * One process per GPU (one single-device context each, NCCL between them) remains possible: izpi_render_tiles_shared. */
int izpi_ctx_create(int n_devices, const int* device_ids, izpi_ctx** out);
int izpi_ctx_num_devices(const izpi_ctx* ctx);
void izpi_ctx_destroy(izpi_ctx* ctx);

/* ---- flattened scene ------------------------------------------------------------------
 * Device primitive record, 80 bytes, 16-byte aligned, stored in WORLD ORDER: the BVH4's
 * reordered Primitives order (bvh4.go:585-590) when the world is a BVH4, else the
 * HitableSlice order.  A 4-primitive leaf is 320 contiguous bytes. */
typedef struct izpi_prim_rec {
  /* TRIANGLE: vertex0.xyz edge1.xyz edge2.xyz      (triangle.go:22-34)
   * SPHERE:   center.xyz radius                    (sphere.go:21-27)
   * X?RECT:   a0 a1 b0 b1 k                        (xyrect.go:18-25 ...)
   * BOX:      pMin.xyz pMax.xyz                    (box.go:16-21) */
  double a[9];
  int32_t orig_id; /* index in the construction-order hitables list (what Hit callers see) */
  uint32_t tag;    /* bits 0-2 IZPI_PRIM_*, bit 3 FlipNormals, bits 4-17 material, bits 18-31 xform index + 1 (0 = none) */
} izpi_prim_rec;

#define IZPI_TAG(type, flip, material, xform1) \
  ((uint32_t)(type) | ((uint32_t)((flip) ? 1 : 0) << 3) | ((uint32_t)(material) << 4) | ((uint32_t)(xform1) << 18))

/* Shading attributes of a triangle, same index as its record (triangle.go:20-50). 176 bytes.
 * vertex1/vertex2 are kept exactly because Triangle.Random (triangle.go:317-326) lerps between
 * the original vertices. */
typedef struct izpi_tri_attr {
  double normal[3], tangent[3], bitangent[3];
  double uv[6]; /* u0 v0 u1 v1 u2 v2 */
  double area;
  double vertex1[3], vertex2[3];
} izpi_tri_attr;

/* Translate(RotateY(.)) wrapper chain (translate.go:16-19, rotate_y.go:19-25). */
typedef struct izpi_xform {
  double sin_theta, cos_theta; /* RotateY; identity = 0, 1 */
  double offset[3];            /* Translate; identity = 0   */
  int32_t has_rotate, has_translate;
} izpi_xform;

/* Thin-lens camera, fields of camera.Camera after camera.New (camera/camera.go:13-58). */
typedef struct izpi_camera {
  double lens_radius, time0, time1, exposure;
  double u[3], v[3], origin[3], lower_left_corner[3], horizontal[3], vertical[3];
} izpi_camera;

typedef struct izpi_scene_desc {
  int32_t world_kind; /* IZPI_WORLD_SLICE | IZPI_WORLD_BVH4 */
  int32_t n_nodes;
  const izpi_bvh4_node* nodes; /* BVH4.Nodes verbatim (128 B each) */
  int32_t n_prims;
  int32_t n_xforms;
  const izpi_prim_rec* prims;
  const izpi_tri_attr* tri_attrs; /* [n_prims]; entries of non-triangles are ignored; may be NULL for trace-only use */
  const izpi_xform* xforms;
  int32_t n_lights;
  int32_t n_materials;
  const int32_t* lights; /* record indices of the scene.Lights members, in Lights order */
  const izpi_material_spec* materials;
  int32_t n_textures, n_spectral_textures;
  const izpi_texture_spec* textures;
  const izpi_spectral_texture_spec* spectral_textures;
  izpi_camera camera;
  int32_t dielectric_has_world; /* Dielectric.SetWorld was called (transport.go:83-89) */
  int32_t reserved;
} izpi_scene_desc;

int izpi_scene_upload(izpi_ctx* ctx, const izpi_scene_desc* desc);

/* ---- scene image: replicate an uploaded scene without rebuilding it ---------------------------------------------------
 * The reference ships the scene to every LAN worker, each of which re-runs ToScene() and NewBVH4
 * (internal/transport/transport.go:53-92); on one NVLink box the flattened arrays are copied instead.  The uploaded
 * scene is an ordered list of device blocks plus a small host header (sizes, pointer tables).  Exporter:
 * izpi_scene_image_size, then izpi_scene_image_export (header bytes; per block its size and device pointer).  Importer:
 * izpi_scene_image_adopt(header) allocates the same blocks on ITS device and returns their pointers, the caller fills
 * them by any transport (cudaMemcpyPeerAsync inside a process -- what a device group does -- or an NCCL broadcast
 * between one-process-per-GPU ranks), then izpi_scene_image_commit re-bases the pointer tables.  The header is opaque
 * and only valid between contexts of the same library build. */
int izpi_scene_image_size(izpi_ctx* src, uint64_t* header_bytes, int32_t* n_blocks);
int izpi_scene_image_export(izpi_ctx* src, void* header, uint64_t* block_bytes, void** d_blocks);
int izpi_scene_image_adopt(izpi_ctx* dst, const void* header, uint64_t header_bytes, void** d_blocks_out);
int izpi_scene_image_commit(izpi_ctx* dst);

/* ---- closest hit ------------------------------------------------------------------------ */
#define IZPI_TRACE_EXACT 0 /* reference traversal order, fp32 SSE-flavour box test, fp64 primitives: bit-exact */
#define IZPI_TRACE_FP32 1  /* optional: fp32 primitive tests, reported separately (not parity-checked bitwise) */

typedef struct izpi_trace_stats {
  uint64_t rays, nodes_visited, prim_tests; /* filled when requested (slower counting kernel) */
  double kernel_ms;                         /* device time of the traversal kernel alone */
} izpi_trace_stats;

/* org/dir: n*3 doubles (xyz interleaved).  prim_id[i] = orig_id of the closest primitive or -1;
 * t[i] = hit parameter (0 on miss).  Host buffers; copies are part of the call. */
int izpi_trace_closest(izpi_ctx* ctx, int64_t n, const double* org_xyz, const double* dir_xyz, double tmin,
                       double tmax, int mode, int32_t* prim_id, double* t, izpi_trace_stats* stats_opt);
/* Same, all four buffers already resident in this context's device memory; launched on `stream`
 * (a cudaStream_t, NULL = the context's stream) and NOT synchronised. */
int izpi_trace_closest_device(izpi_ctx* ctx, int64_t n, const double* d_org_xyz, const double* d_dir_xyz,
                              double tmin, double tmax, int mode, int32_t* d_prim_id, double* d_t, void* stream);
/* Counts the kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t izpi_launch_count(const izpi_ctx* ctx);

/* ---- displacement tessellation -------------------------------------------------------------
 * displacement.ApplyDisplacementMap(triangles, displacementMap, min, max) (internal/displacement/displacement.go:145):
 * adaptive 1->4 tessellation until a triangle spans <= 4 texels of the map and its displacement variation is under
 * the threshold, then TBN displacement of every vertex.  tris15: n x {v0.xyz v1.xyz v2.xyz u0 v0 u1 v1 u2 v2};
 * the map is a W x H fp64 RGBA image whose BLUE channel is the height (image.go:73-101, displacement.go:108-110).
 * per_triangle != 0 reproduces transport.go:633-646 (one call per input triangle, results concatenated);
 * 0 = one call on the whole list.  The result stays on the device until fetched. */
int izpi_displace(izpi_ctx* ctx, int64_t n, const double* tris15, const int32_t* materials, int32_t tex_w, int32_t tex_h,
                  const double* pixels_rgba, double min, double max, int per_triangle, int64_t* n_out);
int izpi_displace_fetch(izpi_ctx* ctx, double* out_tris15, int32_t* out_materials);

/* ---- device-side BVH4 build ------------------------------------------------------------------
 * Optional replacement for hitable.NewBVH4 (internal/hitable/bvh4.go:517-855) when setup time matters: boxes6 =
 * n x {min.xyz max.xyz}, the BoundingBox(time0, time1) of every hitable (fp64).  Builds a Morton-ordered LBVH on the
 * GPU and collapses it into the reference's node format (128-byte BVH4Node, leaf-nodes using slot 0 only, boxes rounded
 * outward to float32, leaves contiguous in the returned permutation: Primitives[i] = hitables[perm[i]]), so BVH4.Hit and
 * every kernel here consume it unchanged.  The tree is NOT the reference's tree; closest hits are the same.  Fails if
 * the tree could overflow the traversal's 64-entry stack (bvh4.go:71).  The result stays in the context until fetched. */
int izpi_bvh4_build(izpi_ctx* ctx, int32_t n, const double* boxes6, int32_t* n_nodes);
int izpi_bvh4_build_fetch(izpi_ctx* ctx, izpi_bvh4_node* nodes, int32_t* perm);

/* Diagnostic (not part of the drop-in surface): the 4-wide fp32 slab test alone, n independent
 * cases -- RayAABB4_SIMD (bvh4_simd_amd64.go:27) -- so the reference's golden masks
 * (bvh4_simd_test.go:54-268) can be replayed on the device.  org/inv: n*3 floats; bounds: n*24 floats
 * (minX[4] minY[4] minZ[4] maxX[4] maxY[4] maxZ[4]); tmax: n floats; masks: n bytes. */
int izpi_debug_ray_aabb4(izpi_ctx* ctx, int32_t n, const float* org, const float* inv, const float* bounds,
                         const float* tmax, uint8_t* masks);

/* Diagnostic: hitable.Hitable.Hit with the full hitrecord.HitRecord (hitrecord.go:6-12) for n rays against the uploaded world:
 * prim_id[i] = original index or -1, out9 = t u v p.xyz normal.xyz (zeros on a miss) -- what shading consumes, so that the
 * reference's exact records (triangle_test.go:69-134: normal 0.8908708063747479, ...) can be checked on the device. */
int izpi_debug_hit(izpi_ctx* ctx, int32_t n, const double* org, const double* dir, double tmin, double tmax, int32_t* prim_id,
                   double* out9);

/* Diagnostic: measured dependent-FMA throughput (TFLOP/s, 2 flops per FMA) of the fp32 (fp64 = 0) or fp64 (fp64 = 1) vector
 * pipe of the context's device: the FLOP side of the traversal roofline (BASELINE north_star). */
int izpi_debug_fma_peak(izpi_ctx* ctx, int fp64, double* tflops);

/* ---- tile rendering --------------------------------------------------------------------- */
#define IZPI_SAMPLER_COLOUR 0   /* sampler/colour.go   */
#define IZPI_SAMPLER_SPECTRAL 1 /* sampler/spectral.go */
#define IZPI_SAMPLER_ALBEDO 2   /* sampler/albedo.go: material albedo at the first hit, black on a miss */
#define IZPI_SAMPLER_NORMAL 3   /* sampler/normal.go: hit-record normal at the first hit */

/* izpi_render_config.flags: measurement modes; the image is unchanged.  Results: izpi_render_get_stats.
 *   IZPI_RENDER_TIMING  CUDA events around every stage of every bounce, with ONE batch in flight instead of two so that a
 *                       stage's time is its own (the frame is slower by the lost overlap);
 *   IZPI_RENDER_STATS   the same plus counting variants of the extend kernels: nodes visited / primitive tests, the
 *                       counters of izpi_trace_stats (a few per cent slower than the plain kernels). */
#define IZPI_RENDER_STATS 1
#define IZPI_RENDER_TIMING 2

typedef struct izpi_render_config {
  int32_t width, height, spp, max_depth;
  int32_t sampler;
  int32_t sample_offset; /* first global sample index rendered by this context (sample-range sharding) */
  int32_t sample_count;  /* samples of every pixel rendered by this context; the mean still divides by spp */
  int32_t flags;         /* IZPI_RENDER_* bits; 0 for production renders */
  double background[3];        /* colours.Black by default */
  const double* bg_wavelengths; /* spectral background SPD (control.proto:64-67); NULL = SpectralBlack */
  const double* bg_values;
  int32_t n_bg;
  int32_t reserved2;
  uint64_t seed;
} izpi_render_config;

int izpi_render_setup(izpi_ctx* ctx, const izpi_render_config* cfg);
/* tiles: n_tiles * 4 uint32 {x0, y0, x1, y1}, inclusive pixel bounds in the reference's workUnit
 * convention (renderer.go:183-186).  Accumulates into the device canvas; if canvas_rgba is not
 * NULL the rendered tiles' pixels are also written there (4*W*H doubles, Float64NRGBA layout,
 * reference row flip rgb.go:41). */
int izpi_render_tiles(izpi_ctx* ctx, int32_t n_tiles, const uint32_t* x0y0x1y1, double* canvas_rgba);
/* The same for a tile list SHARED with other contexts: *cursor (initially 0) is the index of the next unclaimed tile; every
 * sharer passes the same list and the same cursor and claims runs of tiles from it with atomic fetch-adds until the list is
 * exhausted, the way the reference's workers pull work units from one channel (renderer.go:126-147).  `sharers` = how many
 * contexts pull from the cursor (sizes the claims: guided self-scheduling, large runs first, short ones at the end).  The
 * cursor may live in memory shared between processes (one process per GPU).  Each context's canvas then holds the sums of
 * the tiles it claimed and zeros elsewhere, so the merge is a sum (NCCL reduce) or a gather of disjoint pixels.  A device
 * group uses this internally; cursor = NULL is izpi_render_tiles. */
int izpi_render_tiles_shared(izpi_ctx* ctx, int32_t n_tiles, const uint32_t* x0y0x1y1, uint64_t* cursor, int32_t sharers,
                             double* canvas_rgba);
/* worker.RenderTile (internal/worker/render.go:17-75; RenderTileRequest/Response, internal/proto/control/control.proto:75-89):
 * renders ONE tile and returns what the worker streams back, row by row: (y1-y0+1) rows, row r = image row y0 + r (the
 * worker does not flip; the leader places it at ny - pos_y, render/remote.go:63-69), each row strip_height*4*(x1-x0+1)
 * doubles of which the first 4*(x1-x0+1) are {R,G,B,1} (spectral: CIE X,Y,Z) pixel means and the rest zeros, as the
 * reference allocates them.  The Go worker wraps row r into RenderTileResponse{width, height: 1, pos_x: x0, pos_y: y0+r}.
 * Also accumulates into the device canvas like izpi_render_tiles; a tile must be requested once per setup. */
int izpi_render_tile_rows(izpi_ctx* ctx, uint32_t strip_height, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, double* rows);
/* Device pointer of the context's canvas accumulator (4*W*H doubles: per-pixel SUMS over the
 * samples rendered so far, alpha = 1 where written) so that the host runtime can ncclReduce it. */
int izpi_render_canvas_device(izpi_ctx* ctx, double** d_canvas);
/* Divide by spp, apply the epilogue (spectral: FireflyRejection + XYZ->ACEScg), copy to host.  For a device group this
 * is also the merge of the members' disjoint tiles (see izpi_ctx_create). */
int izpi_render_finish(izpi_ctx* ctx, double* canvas_rgba, uint64_t* total_rays);

/* Totals since the last izpi_render_setup.  `rays` is always filled (the reference's numRays, colour.go:38); the visit /
 * test counts and the stage times need IZPI_RENDER_STATS.  Algorithmic bytes of the extend stage:
 * 128 * nodes_visited + 72 * prim_tests + 60 * rays.  For a device group: counts are summed, times are the slowest member's. */
typedef struct izpi_render_stats {
  uint64_t rays, nodes_visited, prim_tests;
  uint64_t extend_launches;
  uint64_t material_bins; /* bins the hits are sorted into between bounces: one per image-textured material + one per class for the rest */
  double extend_ms; /* summed device time of the extend (closest-hit) launches */
  double shade_ms;  /* ... of the shade launches + queue advance */
  double other_ms;  /* ... of ray generation and resolve */
} izpi_render_stats;
int izpi_render_get_stats(izpi_ctx* ctx, izpi_render_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* IZPI_CUDA_H */
