// host_scene.cpp -- host runtime: object constructors, BVH4, flattening, tile scheduling.
// See include/izpi_host.h.  Reference file:line cited per step.
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/izpi_host.h"
#include "../common/vecmath.h"
#include "bvh4_builder.hpp"

namespace izpi {
void set_error(const std::string& msg);  // error.cpp
}
using namespace izpi;

struct izpi_host_scene {
  int32_t world_kind = IZPI_WORLD_BVH4;
  std::vector<izpi_prim_rec> recs;  // world order
  std::vector<izpi_tri_attr> attrs;
  std::vector<izpi_xform> xforms;
  BVH4Build bvh;
  std::vector<int32_t> light_recs, light_ids;
  std::vector<izpi_material_spec> materials;
  std::vector<izpi_texture_spec> textures;
  std::vector<izpi_spectral_texture_spec> spectex;
  std::vector<std::vector<double>> owned;  // deep copies of SPD tables
  izpi_camera camera;
};

namespace {

inline d3 rd3(const double* p) { return mk(p[0], p[1], p[2]); }
inline void wr3(double* o, d3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
inline double fmax2(double a, double b) { return a > b ? a : b; }  // math.Max for ordinary values

void grow(BoxD& a, const BoxD& b) {  // aabb.SurroundingBox (aabb.go:26-39)
  for (int k = 0; k < 3; k++) {
    if (b.mn[k] < a.mn[k]) a.mn[k] = b.mn[k];
    if (b.mx[k] > a.mx[k]) a.mx[k] = b.mx[k];
  }
}

// rect bounding box: k-1e-4 .. k+1e-3 on the plane axis (xyrect.go:90-101, xzrect.go:92-104, yzrect.go:90-101)
BoxD rect_box(int axis, double a0, double a1, double b0, double b1, double k) {
  int ia = axis == 0 ? 1 : 0, ib = axis == 2 ? 1 : 2;
  BoxD b;
  b.mn[axis] = k - 0.0001; b.mx[axis] = k + 0.001;
  b.mn[ia] = a0; b.mx[ia] = a1; b.mn[ib] = b0; b.mx[ib] = b1;
  return b;
}

// NewTriangleWithUV -> NewTriangleWithUVAndNormal (triangle.go:60-134)
void make_triangle(const double* p, izpi_prim_rec& rec, izpi_tri_attr& at, BoxD& box) {
  d3 v0 = rd3(p), v1 = rd3(p + 3), v2 = rd3(p + 6);
  double u0 = p[9], w0 = p[10], u1 = p[11], w1 = p[12], u2 = p[13], w2 = p[14];
  d3 e1 = v1 - v0, e2 = v2 - v0;
  d3 n = cross(e1, e2);
  d3 normal = unit(n);
  double dU1 = u1 - u0, dU2 = u2 - u0, dV1 = w1 - w0, dV2 = w2 - w0;
  double f = 1.0 / (dU1 * dV2 - dU2 * dV1);
  d3 tangent = unit(mk(f * (dV2 * e1.x - dV1 * e2.x), f * (dV2 * e1.y - dV1 * e2.y), f * (dV2 * e1.z - dV1 * e2.z)));
  d3 bitangent = unit(mk(f * (-dU2 * e1.x + dU1 * e2.x), f * (-dU2 * e1.y + dU1 * e2.y), f * (-dU2 * e1.z + dU1 * e2.z)));
  wr3(rec.a, v0); wr3(rec.a + 3, e1); wr3(rec.a + 6, e2);
  wr3(at.normal, normal); wr3(at.tangent, tangent); wr3(at.bitangent, bitangent);
  at.uv[0] = u0; at.uv[1] = w0; at.uv[2] = u1; at.uv[3] = w1; at.uv[4] = u2; at.uv[5] = w2;
  at.area = len(n) / 2.0;
  wr3(at.vertex1, v1); wr3(at.vertex2, v2);
  // vec3.Min3 / Max3, epsilon relative to the largest extent (triangle.go:100-113)
  double mn[3], mx[3];
  for (int k = 0; k < 3; k++) {
    double a = comp(v0, k), b = comp(v1, k), c = comp(v2, k);
    double lo = DBL_MAX, hi = -DBL_MAX;
    if (a < lo) lo = a; if (b < lo) lo = b; if (c < lo) lo = c;
    if (a > hi) hi = a; if (b > hi) hi = b; if (c > hi) hi = c;
    mn[k] = lo; mx[k] = hi;
  }
  double maxDim = fmax2(mx[0] - mn[0], fmax2(mx[1] - mn[1], mx[2] - mn[2]));
  double eps = fmax2(maxDim * 1e-4, 1e-6);
  for (int k = 0; k < 3; k++) { box.mn[k] = mn[k] - eps; box.mx[k] = mx[k] + eps; }
}

// RotateY bounding box (rotate_y.go:27-79)
BoxD rotate_box(const BoxD& b, double s, double c) {
  BoxD r;
  for (int k = 0; k < 3; k++) { r.mn[k] = DBL_MAX; r.mx[k] = -DBL_MAX; }
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
    double x = (double)i * b.mx[0] + (1.0 - (double)i) * b.mn[0];
    double y = (double)j * b.mx[1] + (1.0 - (double)j) * b.mn[1];
    double z = (double)k * b.mx[2] + (1.0 - (double)k) * b.mn[2];
    double t[3] = {c * x + s * z, y, -s * x + c * z};
    for (int a = 0; a < 3; a++) { if (t[a] > r.mx[a]) r.mx[a] = t[a]; if (t[a] < r.mn[a]) r.mn[a] = t[a]; }
  }
  return r;
}

bool material_is_emitter(const izpi_material_spec& m) {  // diffuselight.go:66, dielectric.go:215
  return m.type == IZPI_MAT_DIFFUSE_LIGHT || m.type == IZPI_MAT_DIELECTRIC;
}

}  // namespace

extern "C" {

int izpi_host_scene_create(const izpi_scene_spec* spec, int threads, izpi_host_scene** out) {
  if (!spec || !out || spec->n_prims < 0) { set_error("izpi_host_scene_create: bad argument"); return IZPI_EINVAL; }
  if (spec->n_materials > (1 << 14) - 1) { set_error("too many materials (max 16383)"); return IZPI_EINVAL; }
  auto* s = new izpi_host_scene();
  s->world_kind = spec->world_kind;
  const int n = spec->n_prims;
  std::vector<izpi_prim_rec> recs(n);
  std::vector<izpi_tri_attr> attrs(n);
  std::vector<BoxD> boxes(n);
  std::memset(attrs.data(), 0, sizeof(izpi_tri_attr) * (size_t)n);
  // constructors; triangles (the bulk) in parallel
  int nt = threads < 1 ? 1 : threads;
  std::vector<int> bad(nt, 0);
  std::vector<std::vector<std::pair<int, izpi_xform>>> xf(nt);
  auto construct = [&](int tid) {
    for (int i = (int)((int64_t)n * tid / nt); i < (int)((int64_t)n * (tid + 1) / nt); i++) {
      const izpi_prim_spec& ps = spec->prims[i];
      izpi_prim_rec& r = recs[i];
      std::memset(&r, 0, sizeof(r));
      r.orig_id = i;
      if (ps.material < 0 || ps.material >= spec->n_materials || ps.type < 0 || ps.type > IZPI_PRIM_BOX) { bad[tid] = 1; continue; }
      const double* p = ps.p;
      BoxD& b = boxes[i];
      switch (ps.type) {
        case IZPI_PRIM_TRIANGLE: make_triangle(p, r, attrs[i], b); break;
        case IZPI_PRIM_SPHERE:  // sphere.go:114-123 (center0 == center1)
          for (int k = 0; k < 4; k++) r.a[k] = p[k];
          for (int k = 0; k < 3; k++) { b.mn[k] = p[k] - p[3]; b.mx[k] = p[k] + p[3]; }
          break;
        case IZPI_PRIM_XYRECT: case IZPI_PRIM_XZRECT: case IZPI_PRIM_YZRECT:
          for (int k = 0; k < 5; k++) r.a[k] = p[k];
          b = rect_box(ps.type == IZPI_PRIM_YZRECT ? 0 : (ps.type == IZPI_PRIM_XZRECT ? 1 : 2), p[0], p[1], p[2], p[3], p[4]);
          break;
        case IZPI_PRIM_BOX: {  // box.go:23-46: six rects, sides order +z -z +y -y +x -x
          for (int k = 0; k < 6; k++) r.a[k] = p[k];
          b = rect_box(2, p[0], p[3], p[1], p[4], p[5]);
          grow(b, rect_box(2, p[0], p[3], p[1], p[4], p[2]));
          grow(b, rect_box(1, p[0], p[3], p[2], p[5], p[4]));
          grow(b, rect_box(1, p[0], p[3], p[2], p[5], p[1]));
          grow(b, rect_box(0, p[1], p[4], p[2], p[5], p[3]));
          grow(b, rect_box(0, p[1], p[4], p[2], p[5], p[0]));
          break;
        }
      }
      uint32_t xform1 = 0;
      if (ps.wrap & (IZPI_WRAP_ROTATE_Y | IZPI_WRAP_TRANSLATE)) {
        izpi_xform x;
        std::memset(&x, 0, sizeof(x));
        x.cos_theta = 1.0;
        if (ps.wrap & IZPI_WRAP_ROTATE_Y) {  // rotate_y.go:28-30
          double radians = (M_PI / 180.0) * ps.rotate_y_deg;
          x.sin_theta = std::sin(radians); x.cos_theta = std::cos(radians); x.has_rotate = 1;
          b = rotate_box(b, x.sin_theta, x.cos_theta);
        }
        if (ps.wrap & IZPI_WRAP_TRANSLATE) {  // translate.go:49-55
          for (int k = 0; k < 3; k++) { x.offset[k] = ps.translate[k]; b.mn[k] = b.mn[k] + x.offset[k]; b.mx[k] = b.mx[k] + x.offset[k]; }
          x.has_translate = 1;
        }
        xf[tid].push_back({i, x});
        xform1 = 1;  // patched to the real index below
      }
      r.tag = IZPI_TAG(ps.type, (ps.wrap & IZPI_WRAP_FLIP) != 0, ps.material, xform1);
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(construct, t);
    construct(0);
    for (auto& t : th) t.join();
  }
  for (int t = 0; t < nt; t++) if (bad[t]) { delete s; set_error("primitive with invalid type or material index"); return IZPI_EINVAL; }
  for (int t = 0; t < nt; t++)
    for (auto& pr : xf[t]) {
      if (s->xforms.size() >= (1u << 14) - 2) { delete s; set_error("too many transformed primitives"); return IZPI_EINVAL; }
      s->xforms.push_back(pr.second);
      recs[pr.first].tag = (recs[pr.first].tag & ~(0x3fffu << 18)) | ((uint32_t)s->xforms.size() << 18);
    }

  // world order
  std::vector<int32_t> rec_of(n);
  if (spec->world_kind == IZPI_WORLD_BVH4 && n > 0) {
    if (spec->bvh_builder == IZPI_BVH_DEVICE_LBVH) {
      s->bvh = BuildBVH4Device(boxes);
      if (s->bvh.nodes.empty()) { delete s; return IZPI_ECUDA; }  // message set by the builder; no host fallback
    } else {
      s->bvh = NewBVH4(boxes, spec->bvh_seed, spec->bvh_rand_zero != 0, threads);  // transport.go:76
    }
    s->recs.resize(n); s->attrs.resize(n);
    // world order = the BVH's leaf order: a 256-byte gather per primitive (2.9 GB for config 5), split over the host threads
    // (perm is a permutation, so every thread writes its own slots)
    auto permute = [&](int tid) {
      for (int i = (int)((int64_t)n * tid / nt); i < (int)((int64_t)n * (tid + 1) / nt); i++) {
        int32_t src = s->bvh.perm[i];
        s->recs[i] = recs[src]; s->attrs[i] = attrs[src]; rec_of[src] = i;
      }
    };
    {
      std::vector<std::thread> th;
      for (int t = 1; t < nt; t++) th.emplace_back(permute, t);
      permute(0);
      for (auto& t : th) t.join();
    }
  } else {
    s->recs.swap(recs); s->attrs.swap(attrs);
    for (int i = 0; i < n; i++) rec_of[i] = i;
  }
  // lights = every hitable whose IsEmitter() is true, in hitables order (transport.go:67-72)
  for (int i = 0; i < n; i++)
    if (material_is_emitter(spec->materials[spec->prims[i].material])) { s->light_recs.push_back(rec_of[i]); s->light_ids.push_back(i); }

  s->materials.assign(spec->materials, spec->materials + spec->n_materials);
  s->textures.assign(spec->textures, spec->textures + spec->n_textures);  // image pixels stay borrowed until upload
  s->spectex.assign(spec->spectral_textures, spec->spectral_textures + spec->n_spectral_textures);
  for (auto& t : s->spectex) {
    if (t.type == IZPI_SPEC_IMAGE && (t.n < 0 || t.n >= spec->n_textures || spec->textures[t.n].type != IZPI_TEX_IMAGE)) {
      delete s; set_error("spectral image texture refers to a missing image texture"); return IZPI_EINVAL;
    }
    if (t.type == IZPI_SPEC_TABULATED) {
      s->owned.emplace_back(t.wavelengths, t.wavelengths + t.n); t.wavelengths = s->owned.back().data();
      s->owned.emplace_back(t.values, t.values + t.n); t.values = s->owned.back().data();
    }
  }

  {  // camera.New (camera/camera.go:28-58)
    const izpi_camera_spec& c = spec->camera;
    d3 lookFrom = rd3(c.look_from), lookAt = rd3(c.look_at), vup = rd3(c.vup);
    izpi_camera& o = s->camera;
    o.lens_radius = c.aperture / 2.0;
    double theta = c.vfov * M_PI / 180;
    double halfHeight = std::tan(theta / 2.0);
    double halfWidth = c.aspect * halfHeight;
    d3 w = unit(lookFrom - lookAt);
    d3 u = unit(cross(vup, w));
    d3 v = cross(w, u);
    d3 llc = ((lookFrom - u * (halfWidth * c.focus_dist)) - v * (halfHeight * c.focus_dist)) - w * c.focus_dist;
    wr3(o.u, u); wr3(o.v, v); wr3(o.origin, lookFrom); wr3(o.lower_left_corner, llc);
    wr3(o.horizontal, u * (2.0 * halfWidth * c.focus_dist));
    wr3(o.vertical, v * (2.0 * halfHeight * c.focus_dist));
    o.time0 = c.time0; o.time1 = c.time1; o.exposure = c.exposure;
  }
  *out = s;
  return IZPI_OK;
}

void izpi_host_scene_destroy(izpi_host_scene* s) { delete s; }
int32_t izpi_host_scene_num_nodes(const izpi_host_scene* s) { return s ? (int32_t)s->bvh.nodes.size() : 0; }
int izpi_host_scene_bvh(const izpi_host_scene* s, izpi_bvh4_node* nodes, int32_t* perm) {
  if (!s) return IZPI_EINVAL;
  if (nodes) std::memcpy(nodes, s->bvh.nodes.data(), s->bvh.nodes.size() * sizeof(izpi_bvh4_node));
  if (perm) std::memcpy(perm, s->bvh.perm.data(), s->bvh.perm.size() * sizeof(int32_t));
  return IZPI_OK;
}
int32_t izpi_host_scene_num_lights(const izpi_host_scene* s) { return s ? (int32_t)s->light_ids.size() : 0; }
int izpi_host_scene_lights(const izpi_host_scene* s, int32_t* ids) {
  if (!s || !ids) return IZPI_EINVAL;
  std::memcpy(ids, s->light_ids.data(), s->light_ids.size() * sizeof(int32_t));
  return IZPI_OK;
}

int izpi_host_scene_desc(const izpi_host_scene* s, izpi_scene_desc* d) {
  if (!s || !d) { set_error("izpi_host_scene_desc: bad argument"); return IZPI_EINVAL; }
  std::memset(d, 0, sizeof(*d));
  d->world_kind = s->world_kind;
  d->n_nodes = (int32_t)s->bvh.nodes.size(); d->nodes = s->bvh.nodes.data();
  d->n_prims = (int32_t)s->recs.size(); d->prims = s->recs.data(); d->tri_attrs = s->attrs.data();
  d->n_xforms = (int32_t)s->xforms.size(); d->xforms = s->xforms.data();
  d->n_lights = (int32_t)s->light_recs.size(); d->lights = s->light_recs.data();
  d->n_materials = (int32_t)s->materials.size(); d->materials = s->materials.data();
  d->n_textures = (int32_t)s->textures.size(); d->textures = s->textures.data();
  d->n_spectral_textures = (int32_t)s->spectex.size(); d->spectral_textures = s->spectex.data();
  d->camera = s->camera;
  d->dielectric_has_world = s->world_kind == IZPI_WORLD_BVH4 ? 1 : 0;  // transport.go:83-89 vs scenes.go (never set)
  return IZPI_OK;
}

int izpi_host_scene_upload(const izpi_host_scene* s, izpi_ctx* ctx) {
  izpi_scene_desc d;
  int rc = izpi_host_scene_desc(s, &d);
  if (rc != IZPI_OK) return rc;
  return izpi_scene_upload(ctx, &d);
}

void izpi_host_tiles(int32_t sx, int32_t sy, int32_t* step_x, int32_t* step_y) {
  static const int32_t sizes[] = {32, 25, 24, 20, 16, 12, 10, 8, 5, 4};
  int32_t ax = 0, ay = 0;
  for (int32_t v : sizes) if (sx % v == 0) { ax = v; break; }
  for (int32_t v : sizes) if (sy % v == 0) { ay = v; break; }
  if (step_x) *step_x = ax;
  if (step_y) *step_y = ay;
}

void izpi_host_walk_grid_spiral(int32_t size_x, int32_t size_y, int32_t* xy) {
  const int64_t total = (int64_t)size_x * size_y;
  if (total <= 0 || !xy) return;
  // walked cells, including the ring of cells just outside the grid that the spiral passes through: coordinates -1 .. size
  const int64_t w = (int64_t)size_x + 2 * ((int64_t)(size_x > size_y ? size_x : size_y) + 2);
  const int64_t off = (w - size_x) / 2;
  std::vector<uint8_t> seen((size_t)(w * w), 0);
  auto at = [&](int64_t x, int64_t y) -> uint8_t& { return seen[(size_t)((y + off) * w + (x + off))]; };
  static const int dx[4] = {0, 1, 0, -1}, dy[4] = {-1, 0, 1, 0};  // up, right, down, left (grid.go:13-18)
  int64_t x = size_x / 2, y = size_y / 2, n = 0, k = 0;
  at(x, y) = 1;
  xy[0] = (int32_t)x; xy[1] = (int32_t)y; n = 1;
  while (n < total) {
    const int d = (int)(k % 4);
    if (at(x + dx[d], y + dy[d])) { k--; continue; }  // already walked: keep the previous direction (grid.go:66-69)
    x += dx[d]; y += dy[d];
    at(x, y) = 1;
    if (x >= 0 && x < size_x && y >= 0 && y < size_y) { xy[2 * n] = (int32_t)x; xy[2 * n + 1] = (int32_t)y; n++; }
    k++;
  }
}

int izpi_host_claim_tiles(uint64_t* cursor, int32_t n_tiles, int64_t tile_paths, int64_t batch_paths, int32_t takers,
                          int32_t* begin, int32_t* end) {
  if (!cursor || !begin || !end || n_tiles <= 0) return 0;
  if (tile_paths < 1) tile_paths = 1;
  int64_t grab;
  if (takers <= 1) {
    grab = n_tiles;
  } else {
    const uint64_t seen = __atomic_load_n(cursor, __ATOMIC_RELAXED);  // a stale value only changes the size of the claim
    const int64_t left = seen < (uint64_t)n_tiles ? (int64_t)n_tiles - (int64_t)seen : 0;
    int64_t cap_tiles = batch_paths / tile_paths;
    if (cap_tiles < 1) cap_tiles = 1;
    int64_t min_tiles = (((int64_t)1 << 20) + tile_paths - 1) / tile_paths;
    if (min_tiles > cap_tiles) min_tiles = cap_tiles;
    grab = left / (2 * (int64_t)takers);
    if (grab > cap_tiles) grab = cap_tiles;
    if (grab < min_tiles) grab = min_tiles;
  }
  const uint64_t b = __atomic_fetch_add(cursor, (uint64_t)grab, __ATOMIC_RELAXED);
  if (b >= (uint64_t)n_tiles) return 0;
  *begin = (int32_t)b;
  *end = (int32_t)(b + (uint64_t)grab < (uint64_t)n_tiles ? b + (uint64_t)grab : (uint64_t)n_tiles);
  return 1;
}

int izpi_host_render(izpi_ctx* ctx, const izpi_render_config* cfg, int32_t tile_begin, int32_t tile_end, int32_t finish,
                     double* canvas, uint64_t* total_rays) {
  if (!ctx || !cfg) { set_error("izpi_host_render: bad argument"); return IZPI_EINVAL; }
  int32_t sx, sy;
  izpi_host_tiles(cfg->width, cfg->height, &sx, &sy);
  if (sx == 0 || sy == 0) {  // the reference divides by zero here (renderer.go:117)
    set_error("no tile size divides the image dimensions (common.Tiles)");
    return IZPI_EINVAL;
  }
  int32_t gx = cfg->width / sx, gy = cfg->height / sy;
  int32_t total = gx * gy;
  if (tile_end < 0 || tile_end > total) tile_end = total;
  if (tile_begin < 0) tile_begin = 0;
  int rc = izpi_render_setup(ctx, cfg);
  if (rc != IZPI_OK) return rc;
  std::vector<int32_t> walk((size_t)2 * total);
  izpi_host_walk_grid_spiral(gx, gy, walk.data());  // the reference's queueing order (renderer.go:151)
  std::vector<uint32_t> tiles;
  tiles.reserve((size_t)4 * (tile_end > tile_begin ? tile_end - tile_begin : 0));
  for (int32_t t = tile_begin; t < tile_end; t++) {  // workUnit bounds, renderer.go:183-186
    uint32_t tx = (uint32_t)walk[2 * (size_t)t], ty = (uint32_t)walk[2 * (size_t)t + 1];
    tiles.push_back(tx * sx); tiles.push_back(ty * sy);
    tiles.push_back(tx * sx + (sx - 1)); tiles.push_back(ty * sy + (sy - 1));
  }
  if (!tiles.empty()) {
    rc = izpi_render_tiles(ctx, (int32_t)(tiles.size() / 4), tiles.data(), nullptr);
    if (rc != IZPI_OK) return rc;
  }
  if (finish) return izpi_render_finish(ctx, canvas, total_rays);
  return IZPI_OK;
}

}  // extern "C"
