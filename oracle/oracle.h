/* oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * liboracle.so is a CPU restatement of Izpi's hot path (reference: /root/reference, Go).
 * It is loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, as the checker and as the timed CPU baseline.  The product
 * (izpi_b200/, libizpi_cuda.so, libizpi_host.so) never links or loads it.
 *
 * Parity pinning: the box test, conservative float32 rounding, Triangle ctor/Hit and the
 * BVH4 structure invariants are checked against the reference's own golden vectors
 * (tests/test_oracle_golden.py <- hitable/bvh4_simd_test.go, bvh4_test.go, triangle_test.go,
 * material/dielectric_test.go).  NO reference test pins a closest hit on a mesh or a rendered
 * value, and no Go toolchain exists here to run the reference, so for those outputs this
 * oracle is "parity unpinned" beyond the unit-level vectors.
 */
#ifndef IZPI_ORACLE_H
#define IZPI_ORACLE_H
#include <stdint.h>
#include "../include/izpi_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_scene oracle_scene;

typedef struct oracle_stats {
  uint64_t nodes, tris, spheres, others, rays;
} oracle_stats;

/* unit-level entry points for the reference's golden vectors */
uint8_t oracle_ray_aabb4(int flavour /*0 sse, 1 scalar*/, const float org[3], const float inv[3],
                         const float bounds[24] /*minX[4] minY[4] minZ[4] maxX[4] maxY[4] maxZ[4]*/, float tmax);
float oracle_conservative_float32_min(double v);
float oracle_conservative_float32_max(double v);
double oracle_lcg_next(uint64_t* state);
void oracle_sample_wavelength(double random, double* lambda, double* pdf);
void oracle_cie_values(double lambda, double* xyz);

/* scene */
oracle_scene* oracle_scene_create(const izpi_scene_spec* spec);
void oracle_scene_destroy(oracle_scene* s);
int32_t oracle_scene_num_nodes(const oracle_scene* s);
/* copies Nodes (128 B each) and the leaf-order permutation (Primitives[i] = hitables[perm[i]]) */
void oracle_scene_bvh(const oracle_scene* s, izpi_bvh4_node* nodes, int32_t* perm);
int32_t oracle_scene_num_lights(const oracle_scene* s);
void oracle_scene_lights(const oracle_scene* s, int32_t* prim_ids);
/* bounding box of primitive i as the reference's BoundingBox() returns it: min.xyz max.xyz */
void oracle_prim_bbox(const oracle_scene* s, int32_t prim, double* out6);
/* Triangle ctor fields for primitive i: edge1 edge2 normal tangent bitangent (15) area (1) bbmin bbmax (6) */
int oracle_triangle_fields(const oracle_scene* s, int32_t prim, double* out22);

/* world.Hit for one ray, full record: out = t u v p.xyz n.xyz (9); returns prim id or -1 */
int32_t oracle_hit(const oracle_scene* s, int flavour, const double org[3], const double dir[3], double tmin,
                   double tmax, double* out9);
/* batched closest hit over `threads` host threads: world.Hit(ray, tmin, tmax) -> id (-1 miss), t */
void oracle_trace(const oracle_scene* s, int flavour, int64_t n, const double* org, const double* dir, double tmin,
                  double tmax, int32_t* prim_id, double* t, oracle_stats* stats, int threads);

/* Dielectric.SpectralScatter transmission attenuation for golden test dielectric_test.go:28-45 is
 * covered through oracle_render on a tiny scene; direct material probe: */
double oracle_spectral_texture_value(const oracle_scene* s, int32_t spectral_tex, double lambda);

typedef struct oracle_render_params {
  int32_t width, height, spp, max_depth;
  int32_t sampler;  /* 0 colour (sampler/colour.go), 1 spectral (sampler/spectral.go) */
  int32_t rng_mode; /* 0 reference LCG streams (one per tile), 1 device counter RNG keyed (pixel, sample) */
  int32_t flavour;  /* box test flavour */
  int32_t threads;
  uint64_t seed;
  int32_t x0, y0, x1, y1; /* pixel window to render, inclusive (reference's workUnit); full image = 0,0,W-1,H-1 */
  int32_t epilogue;       /* spectral only: run FireflyRejection + XYZToRGB (renderer.go:216-219) */
  int32_t libm_jitter;    /* 0 = off.  != 0: every libm result on the path is moved by up to +-2 ulp, keyed by this seed: the
                             sensitivity probe that classifies the pixels on which device and oracle may differ (oracle_geom.hpp) */
} oracle_render_params;

/* canvas: 4*W*H doubles, Float64NRGBA layout (y*W+x)*4+c, reference row flip (rgb.go:41) */
void oracle_render(const oracle_scene* s, const oracle_render_params* p, double* canvas, uint64_t* num_rays);
/* worker.RenderTile (internal/worker/render.go:17-75): (y1-y0+1) rows of strip_height*4*(x1-x0+1) doubles, row r = image row
 * y0 + r (no flip), first 4*(x1-x0+1) doubles = {X,Y,Z,1} pixel means, rest zero.  The window fields of p are ignored. */
void oracle_render_tile(const oracle_scene* s, const oracle_render_params* p, uint32_t strip_height, uint32_t x0, uint32_t y0,
                        uint32_t x1, uint32_t y1, double* rows, uint64_t* num_rays);
void oracle_firefly_rejection(double* canvas, int32_t width, int32_t height);      /* firefly_rejection.go:12 */
void oracle_xyz_to_rgb(const double* in, double* out, int32_t width, int32_t height, double exposure); /* rgb_image.go:28 */
void oracle_tiles(int32_t size_x, int32_t size_y, int32_t* step_x, int32_t* step_y);  /* common/tiles.go:6 */

#ifdef __cplusplus
}
#endif
#endif
