/*
 * izpi_scene.h -- plain-C description of an Izpi scene, as it exists on the host
 * BEFORE any acceleration structure is built.
 *
 * This is the input data format of the hot path.  It carries exactly the information
 * the reference's scene loader hands to its object constructors:
 *
 *   - protobuf scenes:   internal/proto/transport/transport.proto:58-281
 *                        -> internal/transport/transport.go:53-92 (ToScene): triangles
 *                        first, then spheres; world = HitableSlice{BVH4(all)}; lights =
 *                        every hitable whose IsEmitter() is true (transport.go:67-72).
 *   - Go object graphs:  internal/scenes/scenes.go:119-155 (CornellBox): rects, boxes,
 *                        FlipNormals/Translate/RotateY wrappers in a plain HitableSlice.
 *
 * Both the CPU oracle (oracle/) and the product host code (izpi_b200/csrc/host/) consume
 * this struct; each runs its OWN constructor math, BVH4 build and flattening on it.
 * All pointers are borrowed for the duration of the call that receives them.
 */
#ifndef IZPI_SCENE_H
#define IZPI_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives (internal/hitable) -------------------------------------------------- */
enum {
  IZPI_PRIM_TRIANGLE = 0, /* hitable/triangle.go:60  NewTriangleWithUV            */
  IZPI_PRIM_SPHERE = 1,   /* hitable/sphere.go:51    NewSphere (center0==center1) */
  IZPI_PRIM_XYRECT = 2,   /* hitable/xyrect.go:27    NewXYRect(x0,x1,y0,y1,k)     */
  IZPI_PRIM_XZRECT = 3,   /* hitable/xzrect.go:28    NewXZRect(x0,x1,z0,z1,k)     */
  IZPI_PRIM_YZRECT = 4,   /* hitable/yzrect.go:27    NewYZRect(y0,y1,z0,z1,k)     */
  IZPI_PRIM_BOX = 5       /* hitable/box.go:23       NewBox(p0,p1)                */
};

/* wrapper flags; nesting order is FlipNormals(Translate(RotateY(base))) */
enum {
  IZPI_WRAP_FLIP = 1,      /* hitable/flip_normals.go:22 */
  IZPI_WRAP_ROTATE_Y = 2,  /* hitable/rotate_y.go:27     */
  IZPI_WRAP_TRANSLATE = 4  /* hitable/translate.go:22    */
};

typedef struct izpi_prim_spec {
  int32_t type;     /* IZPI_PRIM_*                                   */
  int32_t material; /* index into izpi_scene_spec.materials          */
  int32_t wrap;     /* IZPI_WRAP_* bits                              */
  int32_t reserved;
  /* TRIANGLE: v0.xyz v1.xyz v2.xyz u0 v0 u1 v1 u2 v2   (15)
   * SPHERE:   c.xyz r                                   (4)
   * XYRECT:   x0 x1 y0 y1 k ; XZRECT: x0 x1 z0 z1 k ; YZRECT: y0 y1 z0 z1 k   (5)
   * BOX:      p0.xyz p1.xyz                             (6) */
  double p[15];
  double rotate_y_deg; /* used when IZPI_WRAP_ROTATE_Y */
  double translate[3]; /* used when IZPI_WRAP_TRANSLATE */
} izpi_prim_spec; /* 168 bytes */

/* ---- textures (internal/texture) ---------------------------------------------------- */
enum {
  IZPI_TEX_CONSTANT = 0, /* texture/constant.go:14 */
  IZPI_TEX_IMAGE = 1     /* texture/image.go:24 NewFromRawData: W*H*4 float64 RGBA, row-major */
};

typedef struct izpi_texture_spec {
  int32_t type;
  int32_t width, height; /* IMAGE */
  int32_t reserved;
  double color[3];       /* CONSTANT */
  const double* pixels;  /* IMAGE: width*height*4 doubles, index (y*W+x)*4+c */
} izpi_texture_spec;

enum {
  IZPI_SPEC_GAUSSIAN = 0,  /* texture/spectral_constant.go:26 NewSpectralConstant(peak, centre, width) */
  IZPI_SPEC_TABULATED = 1, /* spectral_constant.go:37 NewSpectralConstantFromSPD / :47 NewSpectralNeutral */
  IZPI_SPEC_IMAGE = 2      /* spectral_image.go:61 NewSpectralImageFromImage: RGB image -> 75 spectral buckets */
};

typedef struct izpi_spectral_texture_spec {
  int32_t type;
  int32_t n;                  /* TABULATED: number of samples; IMAGE: index of the RGB image texture */
  double peak, centre, width; /* GAUSSIAN */
  const double* wavelengths;  /* TABULATED [n] */
  const double* values;       /* TABULATED [n] */
} izpi_spectral_texture_spec;

/* ---- materials (internal/material) -------------------------------------------------- */
enum {
  IZPI_MAT_LAMBERT = 0,       /* material/lambertian.go:30,37 */
  IZPI_MAT_METAL = 1,         /* material/metal.go:26         */
  IZPI_MAT_DIELECTRIC = 2,    /* material/dielectric.go:33-60 */
  IZPI_MAT_DIFFUSE_LIGHT = 3, /* material/diffuselight.go:26,33 */
  IZPI_MAT_PBR = 4            /* material/pbr.go:33           */
};

typedef struct izpi_material_spec {
  int32_t type;
  int32_t tex;          /* RGB texture: Lambert albedo / DiffuseLight emit / PBR albedo; -1 = none */
  int32_t spectral_tex; /* spectral: Lambert albedo / DiffuseLight emit / Dielectric refidx; -1 = none */
  int32_t spectral_absorption_tex; /* Dielectric spectral absorption coeff; -1 = none */
  int32_t normal_tex, roughness_tex, metalness_tex; /* PBR; -1 = none */
  int32_t compute_beer_lambert; /* Dielectric (dielectric.go:24) */
  double v[3]; /* Metal albedo | Dielectric RGB absorptionCoeff */
  double s;    /* Metal fuzz   | Dielectric refIdx               */
} izpi_material_spec;

/* ---- camera (internal/camera/camera.go:28) ------------------------------------------ */
typedef struct izpi_camera_spec {
  double look_from[3], look_at[3], vup[3];
  double vfov, aspect, aperture, focus_dist, time0, time1, exposure;
} izpi_camera_spec;

/* ---- scene --------------------------------------------------------------------------- */
enum {
  IZPI_WORLD_SLICE = 0, /* world = HitableSlice(prims)            scenes.go:153 */
  IZPI_WORLD_BVH4 = 1   /* world = HitableSlice{NewBVH4(prims)}   transport.go:76 */
};

enum {
  IZPI_BVH_REFERENCE = 0,   /* random-axis median split, the reference's tree node for node (bvh4.go:558-855) */
  IZPI_BVH_DEVICE_LBVH = 1  /* built on the GPU (Morton LBVH -> BVH4Node array); same closest hits, different tree */
};

typedef struct izpi_scene_spec {
  int32_t world_kind;
  int32_t n_prims;
  const izpi_prim_spec* prims; /* construction order == reference's hitables order */
  int32_t n_materials;
  int32_t n_textures;
  const izpi_material_spec* materials;
  const izpi_texture_spec* textures;
  int32_t n_spectral_textures;
  int32_t reserved;
  const izpi_spectral_texture_spec* spectral_textures;
  izpi_camera_spec camera;
  /* randomFunc injected into newBVH4 (hitable/bvh4.go:558; the reference's own tests
   * inject it, bvh4_test.go:57).  LCG constants of fastrandom.go:7-11:
   * state = (1664525*state + 1013904223) mod 2^32, value = state / 2^32.
   * bvh_rand_zero != 0 selects the tests' `func() float64 { return 0 }`. */
  uint64_t bvh_seed;
  int32_t bvh_rand_zero;
  int32_t bvh_builder; /* IZPI_BVH_REFERENCE (hitable.NewBVH4 restated on the host) | IZPI_BVH_DEVICE_LBVH */
} izpi_scene_spec;

/* The reference's flat BVH4 node, byte for byte (hitable/bvh4.go:23-39). */
typedef struct izpi_bvh4_node {
  float min_x[4], min_y[4], min_z[4];
  float max_x[4], max_y[4], max_z[4];
  int32_t child_index[4];     /* -1 = empty slot; inner: node index; leaf: first primitive */
  int32_t primitive_count[4]; /* >0 leaf, 0 inner */
} izpi_bvh4_node; /* 128 bytes */

#ifdef __cplusplus
}
#endif
#endif /* IZPI_SCENE_H */
