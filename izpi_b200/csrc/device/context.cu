// context.cu -- izpi_ctx lifetime and the one-time scene upload (include/izpi_cuda.h).
#include <cstdlib>
#include <cstring>

#include "dscene.cuh"

using namespace izpi;

void render_state_free(izpi_ctx* ctx);    // render.cu
void displace_result_free(izpi_ctx* ctx);  // displace.cu
void bvh_build_result_free(izpi_ctx* ctx);  // bvh_build.cu

namespace {

template <typename T>
int upload(izpi_ctx* ctx, const T* host, size_t count, const T** dev, size_t align_bytes = 256) {
  *dev = nullptr;
  if (count == 0) return IZPI_OK;
  void* p = nullptr;
  IZ_CUDA(cudaMalloc(&p, count * sizeof(T)));  // cudaMalloc is 256-byte aligned
  (void)align_bytes;
  ctx->scene_allocs.push_back(p);
  IZ_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dev = static_cast<const T*>(p);
  return IZPI_OK;
}

void free_scene(izpi_ctx* ctx) {
  for (void* p : ctx->scene_allocs) cudaFree(p);
  ctx->scene_allocs.clear();
  ctx->has_scene = false;
}

}  // namespace

extern "C" {

int izpi_ctx_create(int n_devices, const int* device_ids, izpi_ctx** out) {
  if (!out) { set_error("izpi_ctx_create: out is NULL"); return IZPI_EINVAL; }
  *out = nullptr;
  if (n_devices != 1) { set_error("izpi_ctx_create: one device per context (process-per-GPU model)"); return IZPI_EINVAL; }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error(std::string("izpi_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    return IZPI_ECUDA;
  }
  int dev = device_ids ? device_ids[0] : 0;
  if (dev < 0 || dev >= count) { set_error("izpi_ctx_create: device id out of range"); return IZPI_EINVAL; }
  IZ_CUDA(cudaSetDevice(dev));
  auto* ctx = new izpi_ctx();
  ctx->device = dev;
  { const char* e = getenv("IZPI_FORCE_SCALAR"); ctx->force_scalar = e && e[0] == '1'; }
  { const char* e = getenv("IZPI_NODE_STRAGGLERS"); if (e && e[0] >= '0' && e[0] <= '7') ctx->node_stragglers = e[0] - '0'; }
  { const char* e = getenv("IZPI_PAIR_STRAGGLERS"); if (e) { int v = atoi(e); if (v >= 0 && v <= 15) ctx->pair_stragglers = v; } }
  { const char* e = getenv("IZPI_TRACE_LANES"); if (e && (e[0] == '2' || e[0] == '4')) ctx->trace_lanes = e[0] - '0'; }
  cudaDeviceProp prop;
  IZ_CUDA(cudaGetDeviceProperties(&prop, dev));
  ctx->sm_count = prop.multiProcessorCount;
  IZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  IZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  IZ_CUDA(cudaEventCreate(&ctx->ev0));
  IZ_CUDA(cudaEventCreate(&ctx->ev1));
  IZ_CUDA(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
  IZ_CUDA(cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
  *out = ctx;
  return IZPI_OK;
}

void izpi_ctx_destroy(izpi_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
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

int izpi_scene_upload(izpi_ctx* ctx, const izpi_scene_desc* d) {
  if (!ctx || !d) { set_error("izpi_scene_upload: bad argument"); return IZPI_EINVAL; }
  if (d->n_prims < 0 || d->n_nodes < 0 || (d->n_prims > 0 && !d->prims) || (d->n_nodes > 0 && !d->nodes)) {
    set_error("izpi_scene_upload: inconsistent primitive / node arrays");
    return IZPI_EINVAL;
  }
  if (d->world_kind == IZPI_WORLD_BVH4 && d->n_prims > 0 && d->n_nodes == 0) {
    set_error("izpi_scene_upload: BVH4 world without nodes");
    return IZPI_EINVAL;
  }
  IZ_CUDA(cudaSetDevice(ctx->device));
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  free_scene(ctx);
  DScene& s = ctx->scene;
  std::memset(&s, 0, sizeof(s));
  s.world_kind = d->world_kind; s.n_nodes = d->n_nodes; s.n_prims = d->n_prims; s.n_xforms = d->n_xforms;
  s.n_lights = d->n_lights; s.n_materials = d->n_materials; s.dielectric_has_world = d->dielectric_has_world;
  s.camera = d->camera;
  int rc;
  const izpi_bvh4_node* dn = nullptr;
  if ((rc = upload(ctx, d->nodes, (size_t)d->n_nodes, &dn)) != IZPI_OK) return rc;
  s.nodes = reinterpret_cast<const float4*>(dn);
  {
    // Child-major node copy for the 4-lanes-per-ray traversal (intersect_g4.cuh).  A child that is one of
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
    bool ok = nn > 0 && d->n_prims < (1 << 29);
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
          if (n.primitive_count[k] > 4 || n.child_index[k] + n.primitive_count[k] > d->n_prims) ok = false;
          c.cnt = n.primitive_count[k];
          continue;
        }
        if (n.child_index[k] < 0 || n.child_index[k] >= nn) { ok = false; break; }
        const izpi_bvh4_node& ch = d->nodes[n.child_index[k]];
        if (pure_leaf(ch)) {
          bool same = ch.min_x[0] == n.min_x[k] && ch.min_y[0] == n.min_y[k] && ch.min_z[0] == n.min_z[k] &&
                      ch.max_x[0] == n.max_x[k] && ch.max_y[0] == n.max_y[k] && ch.max_z[0] == n.max_z[k];
          if (!same || ch.primitive_count[0] > 4 || ch.child_index[0] + ch.primitive_count[0] > d->n_prims) { ok = false; break; }
          c.idx = ch.child_index[0]; c.cnt = ch.primitive_count[0];
        } else {
          for (int q = 0; q < 4; q++) if (ch.child_index[q] != -1 && ch.primitive_count[q] > 0) ok = false;  // mixed node
        }
      }
    }
    s.g4_need = 64;
    if (ok) {
      // Stack entries a ray can need in this tree: visiting a node with k valid children leaves at most k-1 entries below
      // the child being visited (bvh4.go:137-160; folded leaves are pushed like nodes).  When every child index exceeds its
      // parent's (pre-order numbering: NewBVH4 and the device build) one reverse sweep gives the bound; otherwise the
      // reference's fixed 64 stands.
      std::vector<int> need((size_t)nn, 0);
      bool ordered = true;
      for (int i = nn - 1; i >= 0 && ordered; i--) {
        int k = 0, deepest = 0;
        for (int q = 0; q < 4; q++) {
          const Child& c = t[(size_t)i * 4 + q];
          if (c.idx == -1) continue;
          k++;
          if (c.cnt == 0) {
            if (c.idx <= i) { ordered = false; break; }
            if (need[c.idx] > deepest) deepest = need[c.idx];
          }
        }
        need[i] = k > 0 ? (k - 1) + deepest : 0;
      }
      if (ordered && need[0] > 64) {
        // (*BVH4).Hit indexes its [64]int32 stack without a bound check and panics on such a tree (bvh4.go:71,141-145);
        // the kernels' slabs hold 64 entries as well, so refuse it instead of corrupting shared memory
        set_error("izpi_scene_upload: the BVH4 can need " + std::to_string(need[0]) + " traversal stack entries; the reference's stack holds 64 (bvh4.go:71)");
        return IZPI_EINVAL;
      }
      s.g4_need = ordered ? need[0] : 64;
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
  // textures: pixel arrays first, then the table that points at them
  std::vector<DTexture> tex((size_t)d->n_textures);
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
  std::vector<DSpectralTexture> st((size_t)d->n_spectral_textures);
  for (int i = 0; i < d->n_spectral_textures; i++) {
    const izpi_spectral_texture_spec& t = d->spectral_textures[i];
    DSpectralTexture& o = st[i];
    std::memset(&o, 0, sizeof(o));
    o.type = t.type; o.n = t.n; o.peak = t.peak; o.centre = t.centre; o.width = t.width;
    if (t.type == IZPI_SPEC_TABULATED) {
      if ((rc = upload(ctx, t.wavelengths, (size_t)t.n, &o.wavelengths)) != IZPI_OK) return rc;
      if ((rc = upload(ctx, t.values, (size_t)t.n, &o.values)) != IZPI_OK) return rc;
    }
  }
  if ((rc = upload(ctx, st.data(), st.size(), &s.spectex)) != IZPI_OK) return rc;
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));  // borrowed host memory is free to go after this call
  ctx->has_scene = true;
  return IZPI_OK;
}

}  // extern "C"
