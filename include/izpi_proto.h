/*
 * izpi_proto.h -- scene ingest from the reference's protobuf scene format.
 *
 * Replaces, for a B200 box that loads `.izpi` / `.pbtxt` scenes itself or receives them from an unmodified
 * leader over SceneTransportService:
 *
 *   izpi_proto_scene_parse            <- proto.Unmarshal / prototext.Unmarshal into transport.Scene
 *                                        (internal/leader/leader.go:54-74; schema internal/proto/transport/transport.proto:1-281)
 *   izpi_proto_scene_append_triangles <- the StreamTriangles batches a worker receives
 *                                        (transport.proto:283-312; transport.NewTransport's `triangles` argument, transport.go:34-51)
 *   izpi_proto_scene_to_scene         <- (*Transport).ToScene (internal/transport/transport.go:53-92): camera with the aspect
 *                                        override, materials by name, triangles (embedded, then streamed) then spheres, every
 *                                        wire float widened float32 -> float64 (transport.go:595-622), DISPLACE operators through
 *                                        displacement.ApplyDisplacementMap (transport.go:633-646; run on the device, izpi_displace)
 *
 * The result is an izpi_scene_spec (include/izpi_scene.h) that izpi_host_scene_create consumes like any other scene.  The
 * protobuf wire format and text format are decoded by this library itself (no protobuf runtime dependency): unknown
 * fields are skipped in binary input and rejected in text input, as the Go runtime does.
 *
 * Not supported, with the reference's own behaviour noted: CHECKER / NOISE textures (toSceneTexture returns "unknown texture
 * type" for them too, transport.go:395-418), ISOTROPIC materials (only meaningful inside constant media, which the protobuf
 * format cannot express), per-vertex normals (ignored by toSceneTriangle, transport.go:595-650).
 */
#ifndef IZPI_PROTO_H
#define IZPI_PROTO_H

#include <stddef.h>
#include <stdint.h>

#include "izpi_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IZPI_PROTO_BINARY 0 /* .izpi: proto.Marshal(transport.Scene)      (leader.go:55-63) */
#define IZPI_PROTO_TEXT 1   /* .pbtxt: prototext                             (leader.go:64-72) */

#define IZPI_COLOUR_UNSPECIFIED 0
#define IZPI_COLOUR_RGB 1
#define IZPI_COLOUR_SPECTRAL 2 /* the leader switches the `colour` sampler to `spectral` (leader.go:77-80) */

/* A decoded image, keyed by the filename the scene refers to: the `textures` / `displacementMaps` maps handed to
 * transport.NewTransport (transport.go:34-51).  fp64 RGBA, row-major, index (y*W+x)*4+c (texture.NewFromRawData). */
typedef struct izpi_proto_image {
  const char* filename;
  int32_t width, height;
  const double* pixels_rgba;
} izpi_proto_image;

/* An extra entry for SpectralConstantTexture.from_light_source_library.  The reference's whole library
 * (internal/lightsources/lightsources.go:6-466, lookup :468-471) is built in: 39 tabulated SPDs (spectral.NewCIESPD, 75 samples,
 * 380..750 nm @ 5 nm) and the three blackbodies (incandescent_2800k, halogen_3200k, cie_illuminant_a_2856k:
 * spectral.NewBlackbodySPD, spectral.go:275-320), so a scene naming any of the 42 keys converts stand-alone.  Entries given
 * here are looked up FIRST (a deployment with a patched library).  A name that is in neither place is unknown to the
 * reference too, which falls back to CIE illuminant A (transport.go:483-490); so does this library. */
typedef struct izpi_proto_spd {
  const char* name;
  int32_t n;
  const double* wavelengths;
  const double* values;
} izpi_proto_spd;

typedef struct izpi_proto_options {
  double aspect_override; /* leader.go:46: XSize/YSize; 0 = Camera.aspect (transport.go:545-549) */
  int32_t n_textures, n_displacement_maps;
  const izpi_proto_image* textures;
  const izpi_proto_image* displacement_maps;
  int32_t n_light_sources, reserved;
  const izpi_proto_spd* light_sources;
  izpi_ctx* displace_ctx; /* device context for DISPLACE operators; may be NULL when the scene has none */
  uint64_t bvh_seed;      /* forwarded to izpi_scene_spec */
  int32_t bvh_rand_zero, bvh_builder;
} izpi_proto_options;

typedef struct izpi_proto_scene izpi_proto_scene;

int izpi_proto_scene_parse(const void* buf, size_t len, int32_t format, izpi_proto_scene** out);
/* One serialized StreamTrianglesResponse (binary); triangles are appended after the embedded ones (transport.go:572-583). */
int izpi_proto_scene_append_triangles(izpi_proto_scene* s, const void* buf, size_t len);
/* Runs ToScene.  Image pixel pointers in `opt` are borrowed until the scene spec has been consumed. */
int izpi_proto_scene_to_scene(izpi_proto_scene* s, const izpi_proto_options* opt);
/* Valid after izpi_proto_scene_to_scene, until destroy. */
const izpi_scene_spec* izpi_proto_scene_spec(const izpi_proto_scene* s);

const char* izpi_proto_scene_name(const izpi_proto_scene* s);
int32_t izpi_proto_scene_colour_representation(const izpi_proto_scene* s);
uint64_t izpi_proto_scene_total_triangles(const izpi_proto_scene* s); /* Scene.total_triangles (transport.proto:279) */
int32_t izpi_proto_scene_stream_triangles(const izpi_proto_scene* s);  /* Scene.stream_triangles (transport.proto:278) */
int64_t izpi_proto_scene_num_parsed_triangles(const izpi_proto_scene* s);
/* Scene.spectral_background widened to fp64 (control.proto:64-67 carries the same table to workers); returns the count. */
int32_t izpi_proto_scene_background(const izpi_proto_scene* s, const double** wavelengths, const double** values);
/* Filenames the scene needs decoded: image_textures (which = 0) / displacement_maps (which = 1) map values
 * (leader.go:83-110).  Returns the count; name i through *filename (owned by the scene). */
int32_t izpi_proto_scene_num_images(const izpi_proto_scene* s, int32_t which);
const char* izpi_proto_scene_image_filename(const izpi_proto_scene* s, int32_t which, int32_t i);
void izpi_proto_scene_destroy(izpi_proto_scene* s);

#ifdef __cplusplus
}
#endif
#endif /* IZPI_PROTO_H */
