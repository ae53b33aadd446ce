// context.cu -- izpi_ctx lifetime, device groups, the one-time scene upload and its device-to-device replication
// (include/izpi_cuda.h).
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "dscene.cuh"

using namespace izpi;

void render_state_free(izpi_ctx* ctx);    // render.cu
void displace_result_free(izpi_ctx* ctx);  // displace.cu
void bvh_build_result_free(izpi_ctx* ctx);  // bvh_build.cu

namespace izpi {

// One host thread per member of a device group (the reference starts one worker goroutine per core from one process,
// renderer.go:126-138; here one per GPU).  fn(member, index) runs with the member's device current; the first failure's
// code and message are handed back to the calling thread.  A single-device context runs fn inline.
int group_run(izpi_ctx* ctx, const std::function<int(izpi_ctx*, int)>& fn) {
  if (ctx->subs.empty()) return fn(ctx, 0);
  const int n = 1 + (int)ctx->subs.size();
  std::vector<int> rc((size_t)n, IZPI_OK);
  std::vector<std::string> msg((size_t)n);
  auto body = [&](int i) {
    izpi_ctx* m = i == 0 ? ctx : ctx->subs[(size_t)i - 1];
    if (cudaSetDevice(m->device) != cudaSuccess) { rc[i] = IZPI_ECUDA; msg[i] = "cudaSetDevice failed"; return; }
    rc[i] = fn(m, i);
    if (rc[i] != IZPI_OK) msg[i] = izpi_last_error();
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++) th.emplace_back(body, i);
  body(0);
  for (auto& t : th) t.join();
  cudaSetDevice(ctx->device);
  for (int i = 0; i < n; i++)
    if (rc[i] != IZPI_OK) { set_error("device " + std::to_string(i == 0 ? ctx->device : ctx->subs[(size_t)i - 1]->device) + ": " + msg[i]); return rc[i]; }
  return IZPI_OK;
}

}  // namespace izpi

namespace {

template <typename T>
int upload(izpi_ctx* ctx, const T* host, size_t count, const T** dev) {
  *dev = nullptr;
  if (count == 0) return IZPI_OK;
  void* p = nullptr;
  IZ_CUDA(cudaMalloc(&p, count * sizeof(T)));  // cudaMalloc is 256-byte aligned
  ctx->scene_blocks.push_back({p, count * sizeof(T)});
  IZ_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dev = static_cast<const T*>(p);
  return IZPI_OK;
}

void free_scene(izpi_ctx* ctx) {
  for (const SceneBlock& b : ctx->scene_blocks) cudaFree(b.d);
  ctx->scene_blocks.clear();
  ctx->h_textures.clear();
  ctx->h_spectex.clear();
  ctx->adopt_src.clear();
  ctx->has_scene = false;
}

int create_one(int dev, izpi_ctx** out) {
  IZ_CUDA(cudaSetDevice(dev));
  auto* ctx = new izpi_ctx();
  ctx->device = dev;
  { const char* e = getenv("IZPI_FORCE_SCALAR"); ctx->force_scalar = e && e[0] == '1'; }
  { const char* e = getenv("IZPI_NODE_STRAGGLERS"); if (e && e[0] >= '0' && e[0] <= '7') ctx->node_stragglers = e[0] - '0'; }
  { const char* e = getenv("IZPI_PAIR_STRAGGLERS"); if (e) { int v = atoi(e); if (v >= 0 && v <= 15) ctx->pair_stragglers = v; } }
  { const char* e = getenv("IZPI_TRACE_LANES"); if (e && (e[0] == '2' || e[0] == '4')) ctx->trace_lanes = e[0] - '0'; }
  *out = ctx;  // from here on izpi_ctx_destroy cleans up whatever exists
  cudaDeviceProp prop;
  IZ_CUDA(cudaGetDeviceProperties(&prop, dev));
  ctx->sm_count = prop.multiProcessorCount;
  IZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  IZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  IZ_CUDA(cudaEventCreate(&ctx->ev0));
  IZ_CUDA(cudaEventCreate(&ctx->ev1));
  IZ_CUDA(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
  IZ_CUDA(cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
  return IZPI_OK;
}

void destroy_one(izpi_ctx* ctx) {
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  render_state_free(ctx);
  displace_result_free(ctx);
  bvh_build_result_free(ctx);
  free_scene(ctx);
  cudaFree(ctx->d_org); cudaFree(ctx->d_dir); cudaFree(ctx->d_ids); cudaFree(ctx->d_t); cudaFree(ctx->d_counters);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  delete ctx;
}

// ---- BVH4 validation (every tree, whatever its shape) --------------------------------------------------------------
// Checks what (*BVH4).Hit relies on (bvh4.go:86-89 bounds-checks the node index, the primitive ranges index
// b.Primitives) and bounds the traversal stack.  A node with k hit children leaves at most k-1 entries below the child
// being visited (bvh4.go:137-160), so need(i) = (k_i - 1) + max over inner children need(c).  `inner_only` counts the
// pushes of the reference / scalar traversal (leaf slots are tested on the spot); the cooperative kernels also push
// (folded) leaves.  The order is a DFS from the root, so trees in any numbering are handled and cycles are refused.
int validate_bvh4(const izpi_scene_desc* d, int& need_scalar) {
  const int nn = d->n_nodes;
  need_scalar = 0;
  if (nn == 0) return IZPI_OK;
  for (int i = 0; i < nn; i++) {
    const izpi_bvh4_node& n = d->nodes[i];
    for (int k = 0; k < 4; k++) {
      const int ci = n.child_index[k], pc = n.primitive_count[k];
      if (ci == -1) continue;
      const bool bad = ci < -1 || pc < 0 || (pc == 0 && ci >= nn) || (pc > 0 && (int64_t)ci + pc > (int64_t)d->n_prims);
      if (bad) {
        set_error("izpi_scene_upload: BVH4 node " + std::to_string(i) + " slot " + std::to_string(k) + " refers outside the node / primitive arrays");
        return IZPI_EINVAL;
      }
    }
  }
  // iterative DFS: colour 0 = unseen, 1 = on the path, 2 = done
  std::vector<uint8_t> colour((size_t)nn, 0);
  std::vector<int> need((size_t)nn, 0);
  std::vector<std::pair<int, int>> st;  // (node, next slot)
  st.push_back({0, 0});
  colour[0] = 1;
  while (!st.empty()) {
    auto& top = st.back();
    const int i = top.first;
    const izpi_bvh4_node& n = d->nodes[i];
    if (top.second < 4) {
      const int k = top.second++;
      if (n.child_index[k] == -1 || n.primitive_count[k] > 0) continue;
      const int c = n.child_index[k];
      if (colour[c] == 1) { set_error("izpi_scene_upload: the BVH4 contains a cycle through node " + std::to_string(c)); return IZPI_EINVAL; }
      if (colour[c] == 0) { colour[c] = 1; st.push_back({c, 0}); }
      continue;
    }
    int k_inner = 0, deepest = 0;
    for (int k = 0; k < 4; k++)
      if (n.child_index[k] != -1 && n.primitive_count[k] == 0) { k_inner++; if (need[n.child_index[k]] > deepest) deepest = need[n.child_index[k]]; }
    need[i] = (k_inner > 0 ? k_inner - 1 : 0) + deepest;
    colour[i] = 2;
    st.pop_back();
  }
  need_scalar = need[0];
  return IZPI_OK;
}

// ---- replication -----------------------------------------------------------------------------------------------------
struct ImageHeader {
  uint64_t magic, total_bytes;
  int32_t n_blocks, n_tex, n_spectex, reserved;
  DScene scene;
};
constexpr uint64_t kImageMagic = 0x495a5049494d4731ull;  // "IZPIIMG1"

size_t image_bytes(const izpi_ctx* c) {
  return sizeof(ImageHeader) + c->scene_blocks.size() * sizeof(SceneBlock) + c->h_textures.size() * sizeof(DTexture) +
         c->h_spectex.size() * sizeof(DSpectralTexture);
}

// pointer into the exporter's blocks -> the same offset of the local block
template <typename T>
bool rebase(const T*& p, const std::vector<SceneBlock>& src, const std::vector<SceneBlock>& dst) {
  if (!p) return true;
  const char* q = reinterpret_cast<const char*>(p);
  for (size_t i = 0; i < src.size(); i++) {
    const char* b = static_cast<const char*>(src[i].d);
    if (q >= b && q < b + src[i].bytes) { p = reinterpret_cast<const T*>(static_cast<const char*>(dst[i].d) + (q - b)); return true; }
  }
  return false;
}

}  // namespace

extern "C" {

int izpi_ctx_create(int n_devices, const int* device_ids, izpi_ctx** out) {
  if (!out) { set_error("izpi_ctx_create: out is NULL"); return IZPI_EINVAL; }
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) { set_error("izpi_ctx_create: n_devices must be 1..64"); return IZPI_EINVAL; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error(std::string("izpi_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    return IZPI_ECUDA;
  }
  std::vector<int> devs((size_t)n_devices);
  for (int i = 0; i < n_devices; i++) {
    devs[i] = device_ids ? device_ids[i] : i;
    if (devs[i] < 0 || devs[i] >= count) { set_error("izpi_ctx_create: device id out of range"); return IZPI_EINVAL; }
    for (int k = 0; k < i; k++) if (devs[k] == devs[i]) { set_error("izpi_ctx_create: device listed twice"); return IZPI_EINVAL; }
  }
  izpi_ctx* ctx = nullptr;
  int rc = create_one(devs[0], &ctx);
  for (int i = 1; i < n_devices && rc == IZPI_OK; i++) {
    izpi_ctx* sub = nullptr;
    rc = create_one(devs[i], &sub);
    if (sub) ctx->subs.push_back(sub);
  }
  if (rc == IZPI_OK && n_devices > 1) {
    // peer access in both directions: scene replication and the canvas gather then go over NVLink without host staging
    // (when a pair cannot be peers the copies still work, staged by the driver)
    for (int a = 0; a < n_devices; a++) {
      cudaSetDevice(devs[a]);
      for (int b = 0; b < n_devices; b++) {
        if (a == b) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) == cudaSuccess && can) {
          cudaError_t pe = cudaDeviceEnablePeerAccess(devs[b], 0);
          if (pe != cudaSuccess) cudaGetLastError();  // cudaErrorPeerAccessAlreadyEnabled: fine
        }
      }
    }
    cudaSetDevice(devs[0]);
  }
  if (rc != IZPI_OK) {
    std::string keep = izpi_last_error();
    if (ctx) izpi_ctx_destroy(ctx);
    set_error(keep);
    return rc;
  }
  *out = ctx;
  return IZPI_OK;
}

void izpi_ctx_destroy(izpi_ctx* ctx) {
  if (!ctx) return;
  for (izpi_ctx* s : ctx->subs) destroy_one(s);
  ctx->subs.clear();
  destroy_one(ctx);
}

int izpi_ctx_num_devices(const izpi_ctx* ctx) { return ctx ? 1 + (int)ctx->subs.size() : 0; }

static int upload_one(izpi_ctx* ctx, const izpi_scene_desc* d) {
  IZ_CUDA(cudaSetDevice(ctx->device));
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  free_scene(ctx);
  DScene& s = ctx->scene;
  std::memset(&s, 0, sizeof(s));
  s.world_kind = d->world_kind; s.n_nodes = d->n_nodes; s.n_prims = d->n_prims; s.n_xforms = d->n_xforms;
  s.n_lights = d->n_lights; s.n_materials = d->n_materials; s.dielectric_has_world = d->dielectric_has_world;
  s.n_textures = d->n_textures; s.n_spectex = d->n_spectral_textures;
  s.camera = d->camera;
  int rc;
  int need_scalar = 0;
  if (d->world_kind == IZPI_WORLD_BVH4) {
    if ((rc = validate_bvh4(d, need_scalar)) != IZPI_OK) return rc;
    if (need_scalar > 64) {
      // (*BVH4).Hit indexes its [64]int32 stack without a bound check and panics on such a tree (bvh4.go:71,141-145)
      set_error("izpi_scene_upload: the BVH4 can need " + std::to_string(need_scalar) + " traversal stack entries; the reference's stack holds 64 (bvh4.go:71)");
      return IZPI_EINVAL;
    }
  }
  s.scalar_need = need_scalar;
  for (int k = 0; k < 3; k++) { s.world_min[k] = 0.0f; s.world_max[k] = 1.0f; }
  if (d->world_kind == IZPI_WORLD_BVH4 && d->n_nodes > 0) {
    const izpi_bvh4_node& n0 = d->nodes[0];
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = 0; k < 4; k++) {
      if (n0.child_index[k] == -1) continue;
      const float lo[3] = {n0.min_x[k], n0.min_y[k], n0.min_z[k]}, hi[3] = {n0.max_x[k], n0.max_y[k], n0.max_z[k]};
      for (int a = 0; a < 3; a++) { if (lo[a] < mn[a]) mn[a] = lo[a]; if (hi[a] > mx[a]) mx[a] = hi[a]; }
    }
    for (int a = 0; a < 3; a++) if (mx[a] > mn[a]) { s.world_min[a] = mn[a]; s.world_max[a] = mx[a]; }
  }
  for (int i = 0; i < d->n_prims; i++) {
    const int m = (int)((d->prims[i].tag >> 4) & 0x3fffu);
    if (d->n_materials > 0 && m >= d->n_materials) { set_error("izpi_scene_upload: primitive record refers to a missing material"); return IZPI_EINVAL; }
    if (d->n_materials > 0 && d->materials) s.class_mask |= 1 << d->materials[m].type;
  }
  for (int i = 0; i < d->n_lights; i++)
    if (d->lights[i] < 0 || d->lights[i] >= d->n_prims) { set_error("izpi_scene_upload: light index outside the primitive array"); return IZPI_EINVAL; }
  const izpi_bvh4_node* dn = nullptr;
  if ((rc = upload(ctx, d->nodes, (size_t)d->n_nodes, &dn)) != IZPI_OK) return rc;
  s.nodes = reinterpret_cast<const float4*>(dn);
  {
    // Child-major node copy for the cooperative traversals (intersect_g4.cuh / intersect_g2.cuh).  A child that is one of
    // the reference's leaf-nodes (own node, slot 0 only, bvh4.go:737-760) whose box equals the parent's
    // slot box bit for bit (bvh4.go:750-755 vs :782-787) is folded into the parent slot as
    // (first primitive, count).  Trees of any other shape keep the scalar traversal.
    struct Child { float mnx, mny, mnz, mxx, mxy, mxz; int32_t idx, cnt; };
    static_assert(sizeof(Child) == 32, "child record");
    const int nn = d->n_nodes;
    std::vector<Child> t((size_t)nn * 4);
    auto pure_leaf = [&](const izpi_bvh4_node& n) {
      return n.primitive_count[0] > 0 && n.child_index[0] >= 0 && n.child_index[1] == -1 && n.child_index[2] == -1 && n.child_index[3] == -1;
    };
    bool ok = d->world_kind == IZPI_WORLD_BVH4 && nn > 0 && d->n_prims < (1 << 29);
    for (int i = 0; i < nn && ok; i++) {
      const izpi_bvh4_node& n = d->nodes[i];
      const bool self_leaf = pure_leaf(n);
      for (int k = 0; k < 4; k++) {
        Child& c = t[(size_t)i * 4 + k];
        c.mnx = n.min_x[k]; c.mny = n.min_y[k]; c.mnz = n.min_z[k]; c.mxx = n.max_x[k]; c.mxy = n.max_y[k]; c.mxz = n.max_z[k];
        c.idx = n.child_index[k]; c.cnt = 0;
        if (n.child_index[k] == -1) continue;
        if (n.primitive_count[k] > 0) {
          // a direct leaf slot is only representable when its node is a pure leaf-node reached as the root
          if (!self_leaf) ok = false;
          if (n.primitive_count[k] > 4) ok = false;
          c.cnt = n.primitive_count[k];
          continue;
        }
        const izpi_bvh4_node& ch = d->nodes[n.child_index[k]];
        if (pure_leaf(ch)) {
          bool same = ch.min_x[0] == n.min_x[k] && ch.min_y[0] == n.min_y[k] && ch.min_z[0] == n.min_z[k] &&
                      ch.max_x[0] == n.max_x[k] && ch.max_y[0] == n.max_y[k] && ch.max_z[0] == n.max_z[k];
          if (!same || ch.primitive_count[0] > 4) { ok = false; break; }
          c.idx = ch.child_index[0]; c.cnt = ch.primitive_count[0];
        } else {
          for (int q = 0; q < 4; q++) if (ch.child_index[q] != -1 && ch.primitive_count[q] > 0) ok = false;  // mixed node
        }
      }
    }
    s.g4_need = 64;
    if (ok) {
      // Stack entries a ray can need in the FOLDED tree (leaf children are pushed like nodes): same recurrence as
      // validate_bvh4 with k = all valid children, in the DFS post-order that validate_bvh4 has shown to be acyclic.
      std::vector<int> need((size_t)nn, -1);
      std::vector<std::pair<int, int>> st;
      st.push_back({0, 0});
      while (!st.empty()) {
        auto& top = st.back();
        const int i = top.first;
        if (top.second < 4) {
          const Child& c = t[(size_t)i * 4 + top.second++];
          if (c.idx != -1 && c.cnt == 0 && need[c.idx] < 0) st.push_back({c.idx, 0});
          continue;
        }
        int k = 0, deepest = 0;
        for (int q = 0; q < 4; q++) {
          const Child& c = t[(size_t)i * 4 + q];
          if (c.idx == -1) continue;
          k++;
          if (c.cnt == 0 && need[c.idx] > deepest) deepest = need[c.idx];
        }
        need[i] = (k > 0 ? k - 1 : 0) + deepest;
        st.pop_back();
      }
      if (need[0] > 64) ok = false;  // beyond the 64-entry slabs: the scalar traversal (bounded above) handles it
      else s.g4_need = need[0];
    }
    if (ok) {
      // final form of a child record: {minx miny minz maxx | maxy maxz ref cnt} with ref = node index for an inner child and
      // ~((first primitive << 2) | (count - 1)) for a (folded) leaf; an empty slot gets NaN bounds, which fail every
      // comparison of the slab test, so the traversal needs no `ChildIndex == -1` branch (bvh4.go:108-110)
      const float qnan = __builtin_nanf("");
      for (Child& c : t) {
        if (c.idx == -1) { c.mnx = c.mny = c.mnz = c.mxx = c.mxy = c.mxz = qnan; c.idx = 0; c.cnt = 0; }
        else if (c.cnt > 0) c.idx = ~((c.idx << 2) | (c.cnt - 1));
      }
    }
    s.g4_ok = ok ? 1 : 0;
    s.root_is_leaf = (ok && pure_leaf(d->nodes[0])) ? 1 : 0;
    if (ok) {
      const Child* dt = nullptr;
      if ((rc = upload(ctx, t.data(), t.size(), &dt)) != IZPI_OK) return rc;
      IZ_CUDA(cudaStreamSynchronize(ctx->stream));  // `t` is a local
      s.nodes_t = reinterpret_cast<const float4*>(dt);
    }
  }
  if ((rc = upload(ctx, d->prims, (size_t)d->n_prims, &s.prims)) != IZPI_OK) return rc;
  if (d->tri_attrs && (rc = upload(ctx, d->tri_attrs, (size_t)d->n_prims, &s.attrs)) != IZPI_OK) return rc;
  if ((rc = upload(ctx, d->xforms, (size_t)d->n_xforms, &s.xforms)) != IZPI_OK) return rc;
  if ((rc = upload(ctx, d->lights, (size_t)d->n_lights, &s.lights)) != IZPI_OK) return rc;
  if ((rc = upload(ctx, d->materials, (size_t)d->n_materials, &s.materials)) != IZPI_OK) return rc;
  {
    // u, v of a hit are read by image textures only (Constant / SpectralConstant ignore them: constant.go:21,
    // spectral_constant.go:65).  A material without image textures lets the shading skip the sphere's atan2 + asin.
    std::vector<uint8_t> flags((size_t)d->n_materials, 0);
    auto img = [&](int t) { return t >= 0 && t < d->n_textures && d->textures[t].type == IZPI_TEX_IMAGE; };
    auto simg = [&](int t) { return t >= 0 && t < d->n_spectral_textures && d->spectral_textures[t].type == IZPI_SPEC_IMAGE; };
    for (int i = 0; i < d->n_materials; i++) {
      const izpi_material_spec& m = d->materials[i];
      if (img(m.tex) || img(m.normal_tex) || img(m.roughness_tex) || img(m.metalness_tex) || simg(m.spectral_tex) || simg(m.spectral_absorption_tex))
        flags[i] |= kMatNeedsUV;
    }
    if ((rc = upload(ctx, flags.data(), flags.size(), &s.mat_flags)) != IZPI_OK) return rc;
    // material -> bin (dscene.cuh): per class one shared bin for the materials without image textures, then one bin per
    // image-textured material while bins last (kMaxBins); the overflow shares the class's textured bins round-robin.
    // IZPI_MATERIAL_BINS=0 files everything under its class bin (the round-1 behaviour; for A/B measurements).
    const char* env = getenv("IZPI_MATERIAL_BINS");
    const bool by_material = !(env && env[0] == '0');
    std::vector<uint8_t> bin((size_t)d->n_materials, 0);
    int base[5] = {-1, -1, -1, -1, -1};   // first bin of a class: its plain materials, or its first textured one when it has no plain ones
    bool has_plain[5] = {false, false, false, false, false}, base_taken[5] = {false, false, false, false, false};
    std::vector<int> pool[5];             // bins holding textured materials of a class
    s.n_bins = 0;
    for (int i = 0; i < d->n_materials; i++) {
      const int cls = d->materials[i].type;
      if (cls < 0 || cls > IZPI_MAT_PBR) { set_error("izpi_scene_upload: unknown material type"); return IZPI_EINVAL; }
      if (!(flags[i] & kMatNeedsUV) || !by_material) has_plain[cls] = true;
    }
    for (int c = 0; c < 5; c++)
      if ((s.class_mask >> c) & 1) { s.bin_class[s.n_bins] = (uint8_t)c; base[c] = s.n_bins++; }
    size_t rr = 0;
    for (int i = 0; i < d->n_materials; i++) {
      const int cls = d->materials[i].type;
      if (base[cls] < 0) continue;  // no primitive carries this class: never binned
      if (!(flags[i] & kMatNeedsUV) || !by_material) { bin[i] = (uint8_t)base[cls]; continue; }
      if (!has_plain[cls] && !base_taken[cls]) { base_taken[cls] = true; pool[cls].push_back(base[cls]); bin[i] = (uint8_t)base[cls]; continue; }
      if (s.n_bins < kMaxBins) { s.bin_class[s.n_bins] = (uint8_t)cls; pool[cls].push_back(s.n_bins); bin[i] = (uint8_t)s.n_bins++; continue; }
      bin[i] = (uint8_t)(pool[cls].empty() ? base[cls] : pool[cls][rr++ % pool[cls].size()]);  // out of bins: share
    }
    if ((rc = upload(ctx, bin.data(), bin.size(), &s.mat_bin)) != IZPI_OK) return rc;
    IZ_CUDA(cudaStreamSynchronize(ctx->stream));  // `flags`, `bin` are locals
  }
  // textures: pixel arrays first, then the table that points at them
  std::vector<DTexture>& tex = ctx->h_textures;
  tex.assign((size_t)d->n_textures, DTexture{});
  for (int i = 0; i < d->n_textures; i++) {
    const izpi_texture_spec& t = d->textures[i];
    DTexture& o = tex[i];
    std::memset(&o, 0, sizeof(o));
    o.type = t.type; o.width = t.width; o.height = t.height;
    for (int k = 0; k < 3; k++) o.color[k] = t.color[k];
    if (t.type == IZPI_TEX_IMAGE) {
      if (!t.pixels || t.width <= 0 || t.height <= 0) { set_error("izpi_scene_upload: image texture without pixels"); return IZPI_EINVAL; }
      if ((rc = upload(ctx, t.pixels, (size_t)t.width * t.height * 4, &o.pixels)) != IZPI_OK) return rc;
    }
  }
  if ((rc = upload(ctx, tex.data(), tex.size(), &s.textures)) != IZPI_OK) return rc;
  std::vector<DSpectralTexture>& st = ctx->h_spectex;
  st.assign((size_t)d->n_spectral_textures, DSpectralTexture{});
  for (int i = 0; i < d->n_spectral_textures; i++) {
    const izpi_spectral_texture_spec& t = d->spectral_textures[i];
    DSpectralTexture& o = st[i];
    std::memset(&o, 0, sizeof(o));
    o.type = t.type; o.n = t.n; o.peak = t.peak; o.centre = t.centre; o.width = t.width;
    if (t.type == IZPI_SPEC_TABULATED) {
      if (t.n < 0 || (t.n > 0 && (!t.wavelengths || !t.values))) { set_error("izpi_scene_upload: tabulated spectral texture without samples"); return IZPI_EINVAL; }
      o.sorted = 1;
      for (int k = 0; k + 1 < t.n; k++) if (!(t.wavelengths[k] < t.wavelengths[k + 1])) o.sorted = 0;
      if ((rc = upload(ctx, t.wavelengths, (size_t)t.n, &o.wavelengths)) != IZPI_OK) return rc;
      if ((rc = upload(ctx, t.values, (size_t)t.n, &o.values)) != IZPI_OK) return rc;
    }
  }
  if ((rc = upload(ctx, st.data(), st.size(), &s.spectex)) != IZPI_OK) return rc;
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));  // borrowed host memory is free to go after this call
  ctx->has_scene = true;
  return IZPI_OK;
}

// ---- scene image: what another device needs to own a copy of an uploaded scene ---------------------------------------
int izpi_scene_image_size(izpi_ctx* src, uint64_t* header_bytes, int32_t* n_blocks) {
  if (!src || !header_bytes || !n_blocks) { set_error("izpi_scene_image_size: bad argument"); return IZPI_EINVAL; }
  if (!src->has_scene) { set_error("izpi_scene_image_size: no scene uploaded"); return IZPI_ESTATE; }
  *header_bytes = image_bytes(src);
  *n_blocks = (int32_t)src->scene_blocks.size();
  return IZPI_OK;
}

int izpi_scene_image_export(izpi_ctx* src, void* header, uint64_t* block_bytes, void** d_blocks) {
  if (!src || !header) { set_error("izpi_scene_image_export: bad argument"); return IZPI_EINVAL; }
  if (!src->has_scene) { set_error("izpi_scene_image_export: no scene uploaded"); return IZPI_ESTATE; }
  ImageHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kImageMagic; h.total_bytes = image_bytes(src);
  h.n_blocks = (int32_t)src->scene_blocks.size(); h.n_tex = (int32_t)src->h_textures.size(); h.n_spectex = (int32_t)src->h_spectex.size();
  h.scene = src->scene;
  char* o = static_cast<char*>(header);
  std::memcpy(o, &h, sizeof(h)); o += sizeof(h);
  if (h.n_blocks) std::memcpy(o, src->scene_blocks.data(), (size_t)h.n_blocks * sizeof(SceneBlock));
  o += (size_t)h.n_blocks * sizeof(SceneBlock);
  if (h.n_tex) std::memcpy(o, src->h_textures.data(), (size_t)h.n_tex * sizeof(DTexture));
  o += (size_t)h.n_tex * sizeof(DTexture);
  if (h.n_spectex) std::memcpy(o, src->h_spectex.data(), (size_t)h.n_spectex * sizeof(DSpectralTexture));
  for (int i = 0; i < h.n_blocks; i++) {
    if (block_bytes) block_bytes[i] = src->scene_blocks[i].bytes;
    if (d_blocks) d_blocks[i] = src->scene_blocks[i].d;
  }
  return IZPI_OK;
}

int izpi_scene_image_adopt(izpi_ctx* dst, const void* header, uint64_t header_bytes, void** d_blocks_out) {
  if (!dst || !header || header_bytes < sizeof(ImageHeader)) { set_error("izpi_scene_image_adopt: bad argument"); return IZPI_EINVAL; }
  ImageHeader h;
  std::memcpy(&h, header, sizeof(h));
  if (h.magic != kImageMagic || h.total_bytes != header_bytes || h.n_blocks < 0 || h.n_tex < 0 || h.n_spectex < 0 ||
      sizeof(ImageHeader) + (size_t)h.n_blocks * sizeof(SceneBlock) + (size_t)h.n_tex * sizeof(DTexture) +
              (size_t)h.n_spectex * sizeof(DSpectralTexture) != header_bytes) {
    set_error("izpi_scene_image_adopt: not a scene image of this library build");
    return IZPI_EINVAL;
  }
  IZ_CUDA(cudaSetDevice(dst->device));
  IZ_CUDA(cudaStreamSynchronize(dst->stream));
  free_scene(dst);
  const char* p = static_cast<const char*>(header) + sizeof(h);
  dst->adopt_src.resize((size_t)h.n_blocks);
  if (h.n_blocks) std::memcpy(dst->adopt_src.data(), p, (size_t)h.n_blocks * sizeof(SceneBlock));
  p += (size_t)h.n_blocks * sizeof(SceneBlock);
  dst->h_textures.resize((size_t)h.n_tex);
  if (h.n_tex) std::memcpy(dst->h_textures.data(), p, (size_t)h.n_tex * sizeof(DTexture));
  p += (size_t)h.n_tex * sizeof(DTexture);
  dst->h_spectex.resize((size_t)h.n_spectex);
  if (h.n_spectex) std::memcpy(dst->h_spectex.data(), p, (size_t)h.n_spectex * sizeof(DSpectralTexture));
  dst->scene = h.scene;
  for (int i = 0; i < h.n_blocks; i++) {
    void* d = nullptr;
    IZ_CUDA(cudaMalloc(&d, dst->adopt_src[i].bytes));
    dst->scene_blocks.push_back({d, dst->adopt_src[i].bytes});
    if (d_blocks_out) d_blocks_out[i] = d;
  }
  return IZPI_OK;
}

int izpi_scene_image_commit(izpi_ctx* dst) {
  if (!dst) { set_error("izpi_scene_image_commit: bad argument"); return IZPI_EINVAL; }
  if (dst->adopt_src.size() != dst->scene_blocks.size() || (dst->adopt_src.empty() && dst->scene.n_prims > 0)) {
    set_error("izpi_scene_image_commit: izpi_scene_image_adopt has not been called");
    return IZPI_ESTATE;
  }
  IZ_CUDA(cudaSetDevice(dst->device));
  const auto& src = dst->adopt_src;
  const auto& loc = dst->scene_blocks;
  DScene& s = dst->scene;
  bool ok = rebase(s.nodes, src, loc) && rebase(s.nodes_t, src, loc) && rebase(s.prims, src, loc) && rebase(s.attrs, src, loc) &&
            rebase(s.xforms, src, loc) && rebase(s.lights, src, loc) && rebase(s.materials, src, loc) && rebase(s.textures, src, loc) &&
            rebase(s.spectex, src, loc) && rebase(s.mat_flags, src, loc) && rebase(s.mat_bin, src, loc);
  for (DTexture& t : dst->h_textures) ok = ok && rebase(t.pixels, src, loc);
  for (DSpectralTexture& t : dst->h_spectex) ok = ok && rebase(t.wavelengths, src, loc) && rebase(t.values, src, loc);
  if (!ok) { set_error("izpi_scene_image_commit: the image's pointer tables do not match its blocks"); return IZPI_EINVAL; }
  // the two pointer tables arrived with the exporter's addresses: overwrite them with the re-based ones
  if (!dst->h_textures.empty())
    IZ_CUDA(cudaMemcpyAsync(const_cast<DTexture*>(s.textures), dst->h_textures.data(), dst->h_textures.size() * sizeof(DTexture),
                            cudaMemcpyHostToDevice, dst->stream));
  if (!dst->h_spectex.empty())
    IZ_CUDA(cudaMemcpyAsync(const_cast<DSpectralTexture*>(s.spectex), dst->h_spectex.data(), dst->h_spectex.size() * sizeof(DSpectralTexture),
                            cudaMemcpyHostToDevice, dst->stream));
  IZ_CUDA(cudaStreamSynchronize(dst->stream));
  dst->adopt_src.clear();
  dst->has_scene = true;
  return IZPI_OK;
}

int izpi_scene_upload(izpi_ctx* ctx, const izpi_scene_desc* d) {
  if (!ctx || !d) { set_error("izpi_scene_upload: bad argument"); return IZPI_EINVAL; }
  if (d->n_prims < 0 || d->n_nodes < 0 || (d->n_prims > 0 && !d->prims) || (d->n_nodes > 0 && !d->nodes) ||
      d->n_lights < 0 || (d->n_lights > 0 && !d->lights) || d->n_materials < 0 || d->n_textures < 0 || d->n_spectral_textures < 0) {
    set_error("izpi_scene_upload: inconsistent primitive / node arrays");
    return IZPI_EINVAL;
  }
  if (d->world_kind == IZPI_WORLD_BVH4 && d->n_prims > 0 && d->n_nodes == 0) {
    set_error("izpi_scene_upload: BVH4 world without nodes");
    return IZPI_EINVAL;
  }
  int rc = upload_one(ctx, d);
  if (rc != IZPI_OK || ctx->subs.empty()) return rc;
  // Device group: the scene was flattened and validated ONCE; every other member receives the blocks device-to-device
  // (one cudaMemcpyPeerAsync per block, NVLink when the pair are peers) and re-bases the pointer tables.
  std::vector<uint8_t> header(image_bytes(ctx));
  if ((rc = izpi_scene_image_export(ctx, header.data(), nullptr, nullptr)) != IZPI_OK) return rc;
  rc = group_run(ctx, [&](izpi_ctx* m, int i) -> int {
    if (i == 0) return IZPI_OK;
    std::vector<void*> blocks(ctx->scene_blocks.size());
    int r = izpi_scene_image_adopt(m, header.data(), header.size(), blocks.data());
    if (r != IZPI_OK) return r;
    for (size_t b = 0; b < blocks.size(); b++)
      IZ_CUDA(cudaMemcpyPeerAsync(blocks[b], m->device, ctx->scene_blocks[b].d, ctx->device, ctx->scene_blocks[b].bytes, m->stream));
    IZ_CUDA(cudaStreamSynchronize(m->stream));
    return izpi_scene_image_commit(m);
  });
  return rc;
}

}  // extern "C"
