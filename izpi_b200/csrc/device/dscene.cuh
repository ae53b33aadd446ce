// dscene.cuh -- the scene as it lives in HBM (uploaded once by izpi_scene_upload) and the
// per-context state.  Layout rationale in DESIGN.md §3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

#include "../../../include/izpi_cuda.h"
#include "../common/vecmath.h"

namespace izpi {

void set_error(const std::string& msg);

struct DTexture {  // texture.Constant / texture.ImageTxt
  int32_t type, width, height, pad;
  double color[3];
  const double* pixels;  // device, RGBA fp64 (32 B per texel, 16-B aligned)
};
struct DSpectralTexture {  // texture.SpectralConstant
  int32_t type, n;
  int32_t sorted, pad;  // TABULATED: wavelengths strictly increasing -> the bracketing pair is found by bisection (same pair as the reference's scan)
  double peak, centre, width;
  const double* wavelengths;  // device
  const double* values;       // device
};

struct DScene {
  int32_t world_kind, n_nodes, n_prims, n_xforms, n_lights, n_materials, dielectric_has_world;
  int32_t g4_ok;                // nodes_t is usable (reference-shaped tree: leaves are own slot-0-only nodes)
  int32_t root_is_leaf;
  int32_t g4_need;              // most stack entries a ray can need in THIS tree (<= 64, bvh4.go:71); picks the slab size of the 2-lane kernel
  int32_t class_mask;           // bit c set <=> some primitive carries a material of class c (IZPI_MAT_*): shade launches of absent classes are skipped
  int32_t n_textures, n_spectex;
  float world_min[3], world_max[3];  // union of the root node's child boxes (BVH4 worlds): quantises ray origins for the coherence sort
  int32_t scalar_need;          // stack entries the thread-per-ray traversal can need in this tree (<= 64): sizes its shared-memory stack
  int32_t pad0;
  // Material bins of the wavefront renderer (north_star 3: paths sorted by material ID between bounces).  A hit is filed
  // under mat_bin[material]; a bin holds ONE class (IZPI_MAT_*), every material with image textures has a bin of its own
  // (a warp of the shade kernel then samples one texture set), materials without image textures share their class's bin.
  int32_t n_bins;
  uint8_t bin_class[32];
  const uint8_t* mat_bin;
  const uint8_t* mat_flags;     // per material: kMatNeedsUV when any of its textures is an image (else sphere UVs -- atan2 + asin -- are never read)
  const float4* nodes;          // 8 x float4 per BVH4Node, verbatim SoA layout, 128-B aligned
  const float4* nodes_t;        // child-major copy: 4 x {minx miny minz maxx | maxy maxz ref cnt}, leaf-nodes folded in, empty slots NaN (context.cu)
  const izpi_prim_rec* prims;   // 80-B records in world order, 16-B aligned
  const izpi_tri_attr* attrs;   // 128-B records, same index
  const izpi_xform* xforms;
  const int32_t* lights;        // record indices
  const izpi_material_spec* materials;
  const DTexture* textures;
  const DSpectralTexture* spectex;
  izpi_camera camera;
};

constexpr uint8_t kMatNeedsUV = 1;
constexpr int kMaxBins = 32;

#define IZ_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      izpi::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
      return IZPI_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

}  // namespace izpi

struct RenderState;  // render.cu

namespace izpi {
// One cudaMalloc'ed piece of the uploaded scene.  The scene in HBM is the ordered list of these blocks plus the pointer
// tables that refer into them (DScene, the texture tables); that is what izpi_scene_image_* ships to another device.
struct SceneBlock {
  void* d;
  size_t bytes;
};
// cudaOccupancyMaxActiveBlocksPerMultiprocessor results, per CONTEXT (= per device): 0 = not asked yet
struct OccupancyCache {
  int trace_g2[2] = {0, 0}, trace_g4 = 0, trace_scalar = 0;
  int ext = 0, ext4 = 0, ext2[2] = {0, 0};
  size_t ext_smem = 0;  // shared memory per block the `ext` figure was asked for (depends on the uploaded tree)
};
}  // namespace izpi

struct izpi_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second lane of the copy/compute pipeline of izpi_trace_closest
  bool has_scene = false;
  int node_stragglers = 4;    // IZPI_NODE_STRAGGLERS: node-phase exit threshold of the 4-lanes-per-ray kernels
  int pair_stragglers = 7;    // IZPI_PAIR_STRAGGLERS: the same for the 2-lanes-per-ray kernel (pairs)
  int trace_lanes = 2;        // IZPI_TRACE_LANES=2|4: lanes per ray of izpi_trace_closest
  bool force_scalar = false;  // IZPI_FORCE_SCALAR=1: thread-per-ray traversal even for reference-shaped trees
  izpi::DScene scene{};
  std::vector<izpi::SceneBlock> scene_blocks;         // freed on re-upload / destroy
  std::vector<izpi::DTexture> h_textures;             // host mirrors of the two pointer tables (re-based by izpi_scene_image_commit)
  std::vector<izpi::DSpectralTexture> h_spectex;
  std::vector<izpi::SceneBlock> adopt_src;            // izpi_scene_image_adopt: the exporter's block table, until commit
  bool cie_uploaded = false;  // c_cie lives in __constant__ memory, which is per device: tracked per context
  izpi::OccupancyCache occ;
  // trace scratch (grown on demand)
  double* d_org = nullptr; double* d_dir = nullptr; int32_t* d_ids = nullptr; double* d_t = nullptr;
  int64_t ray_capacity = 0;
  unsigned long long* d_counters = nullptr;  // [0] work-queue head, [1] nodes visited, [2] primitive tests
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint64_t launches = 0;
  RenderState* render = nullptr;
  void* displace = nullptr;  // result of the last izpi_displace (displace.cu)
  void* bvh_build = nullptr;  // result of the last izpi_bvh4_build (bvh_build.cu)
  // Device group (izpi_ctx_create with n_devices > 1): this context is the group's first device; `subs` are single-device
  // contexts for device_ids[1..].  Scene uploads are replicated device-to-device, ray batches and tiles are split across
  // the members by one host thread per device (group.cpp-style code in context.cu / render.cu / trace.cu).
  std::vector<izpi_ctx*> subs;
  uint64_t group_cursor = 0;  // shared tile cursor of the group's current izpi_render_tiles call
};

namespace izpi {
// context.cu: fn(member, index) on one host thread per member of the device group (inline for a single device)
int group_run(izpi_ctx* ctx, const std::function<int(izpi_ctx*, int)>& fn);
}  // namespace izpi
