// oracle_api.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Scene construction from izpi_scene_spec (the role of transport.ToScene, transport.go:53-92, and
// of scenes.CornellBox, scenes.go:119-155), batched closest hit, and the tile render loop
// (render/rgb.go:12-57, render/spectral.go:14-106, renderer.go:108-222).
#include "oracle.h"
#include "oracle_shade.hpp"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace orc;

namespace orc {
bool MaterialIsEmitter(const Material* m) { return m && m->IsEmitter(); }
const Texture* MaterialNormalMap(const Material* m) { return (m && m->type == IZPI_MAT_PBR) ? m->normalMap : nullptr; }  // pbr.go:276, non_pbr.go:7
Vec3 TextureValue(const Texture* t, double u, double v, const Vec3&) { return t->Value(u, v); }
}  // namespace orc

struct oracle_scene {
  std::vector<std::unique_ptr<Hitable>> owned;
  std::vector<Hitable*> hitables;  // reference's `hitables` slice, construction order
  std::vector<Texture> textures;
  std::vector<SpectralTexture> spectex;
  std::vector<Material> materials;
  std::unique_ptr<BVH4> bvh;
  HitableSlice world, lights;
  std::vector<int32_t> light_ids;
  Camera camera;
  izpi_camera_spec camspec;
};

static std::unique_ptr<Hitable> make_prim(const izpi_prim_spec& ps, int32_t id, const std::vector<Material>& mats) {
  const Material* m = (ps.material >= 0 && ps.material < (int32_t)mats.size()) ? &mats[ps.material] : nullptr;
  const double* p = ps.p;
  std::unique_ptr<Hitable> h;
  switch (ps.type) {
    case IZPI_PRIM_TRIANGLE:
      h = std::make_unique<Triangle>(V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), p[9], p[10], p[11],
                                     p[12], p[13], p[14], m);
      break;
    case IZPI_PRIM_SPHERE: {
      auto s = std::make_unique<Sphere>();
      s->center0 = s->center1 = V(p[0], p[1], p[2]);
      s->time0 = 0; s->time1 = 1;  // transport.go:679, scenes.go:133
      s->radius = p[3]; s->material = m;
      h = std::move(s);
      break;
    }
    case IZPI_PRIM_XYRECT: case IZPI_PRIM_XZRECT: case IZPI_PRIM_YZRECT: {
      auto r = std::make_unique<Rect>();
      r->axis = ps.type == IZPI_PRIM_YZRECT ? 0 : (ps.type == IZPI_PRIM_XZRECT ? 1 : 2);
      r->a0 = p[0]; r->a1 = p[1]; r->b0 = p[2]; r->b1 = p[3]; r->k = p[4]; r->material = m;
      h = std::move(r);
      break;
    }
    case IZPI_PRIM_BOX:
      h = std::make_unique<Box>(V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), m);
      break;
    default:
      return nullptr;
  }
  h->prim_id = id;
  if (ps.wrap & IZPI_WRAP_ROTATE_Y) {
    auto r = std::make_unique<RotateY>();
    r->h = std::move(h); r->init(ps.rotate_y_deg); r->prim_id = id;
    h = std::move(r);
  }
  if (ps.wrap & IZPI_WRAP_TRANSLATE) {
    auto t = std::make_unique<Translate>();
    t->h = std::move(h); t->offset = V(ps.translate[0], ps.translate[1], ps.translate[2]); t->prim_id = id;
    h = std::move(t);
  }
  if (ps.wrap & IZPI_WRAP_FLIP) {
    auto f = std::make_unique<FlipNormals>();
    f->h = std::move(h); f->prim_id = id;
    h = std::move(f);
  }
  return h;
}

extern "C" {

uint8_t oracle_ray_aabb4(int flavour, const float org[3], const float inv[3], const float b[24], float tmax) {
  izpi_bvh4_node n;
  std::memcpy(n.min_x, b, 96);
  for (int i = 0; i < 4; i++) { n.child_index[i] = 0; n.primitive_count[i] = 0; }
  return RayAABB4(flavour, org, inv, n, tmax);
}
float oracle_conservative_float32_min(double v) { return conservativeFloat32Min(v); }
float oracle_conservative_float32_max(double v) { return conservativeFloat32Max(v); }
double oracle_lcg_next(uint64_t* state) { Rng r; r.state = *state; double v = r.Float64(); *state = r.state; return v; }
void oracle_sample_wavelength(double random, double* lambda, double* pdf) { SampleWavelength(random, *lambda, *pdf); }
void oracle_cie_values(double lambda, double* xyz) { GetCIEValues(lambda, xyz[0], xyz[1], xyz[2]); }

oracle_scene* oracle_scene_create(const izpi_scene_spec* spec) {
  auto* s = new oracle_scene();
  s->textures.resize(spec->n_textures);
  for (int i = 0; i < spec->n_textures; i++) {
    const izpi_texture_spec& t = spec->textures[i];
    Texture& o = s->textures[i];
    o.type = t.type; o.color = V(t.color[0], t.color[1], t.color[2]);
    o.sizeX = t.width; o.sizeY = t.height; o.pixels = t.pixels;  // borrowed: caller keeps images alive
  }
  s->spectex.resize(spec->n_spectral_textures);
  for (int i = 0; i < spec->n_spectral_textures; i++) {
    const izpi_spectral_texture_spec& t = spec->spectral_textures[i];
    SpectralTexture& o = s->spectex[i];
    o.type = t.type; o.peak = t.peak; o.centre = t.centre; o.width = t.width;
    if (t.type == IZPI_SPEC_IMAGE) o.image = (t.n >= 0 && t.n < (int)s->textures.size()) ? &s->textures[t.n] : nullptr;
    if (t.type == IZPI_SPEC_TABULATED) {
      o.spd.wavelengths.assign(t.wavelengths, t.wavelengths + t.n);
      o.spd.values.assign(t.values, t.values + t.n);
    }
  }
  auto tex = [&](int i) -> const Texture* { return (i >= 0 && i < (int)s->textures.size()) ? &s->textures[i] : nullptr; };
  auto stex = [&](int i) -> const SpectralTexture* { return (i >= 0 && i < (int)s->spectex.size()) ? &s->spectex[i] : nullptr; };
  s->materials.resize(spec->n_materials);
  for (int i = 0; i < spec->n_materials; i++) {
    const izpi_material_spec& m = spec->materials[i];
    Material& o = s->materials[i];
    o.type = m.type; o.tex = tex(m.tex); o.spectral = stex(m.spectral_tex);
    o.spectralAbsorption = stex(m.spectral_absorption_tex);
    o.normalMap = tex(m.normal_tex); o.roughness = tex(m.roughness_tex); o.metalness = tex(m.metalness_tex);
    o.computeBeerLambert = m.compute_beer_lambert != 0;
    o.v = V(m.v[0], m.v[1], m.v[2]); o.s = m.s;
  }
  for (int i = 0; i < spec->n_prims; i++) {
    auto h = make_prim(spec->prims[i], i, s->materials);
    if (!h) { delete s; return nullptr; }
    s->hitables.push_back(h.get());
    s->owned.push_back(std::move(h));
  }
  for (Hitable* h : s->hitables)  // transport.go:67-72, scenes.go:136-141
    if (h->IsEmitter()) { s->lights.hitables.push_back(h); s->light_ids.push_back(h->prim_id); }
  if (spec->world_kind == IZPI_WORLD_BVH4) {
    Rng lcg; lcg.mode = 0; lcg.state = spec->bvh_seed;
    s->bvh = newBVH4(s->hitables, &lcg, spec->bvh_rand_zero != 0);
    if (s->bvh) s->world.hitables.push_back(s->bvh.get());
    for (Material& m : s->materials)  // transport.go:83-89
      if (m.type == IZPI_MAT_DIELECTRIC) m.world = &s->world;
  } else {
    s->world.hitables = s->hitables;
  }
  s->camspec = spec->camera;
  s->camera.init(spec->camera);
  return s;
}
void oracle_scene_destroy(oracle_scene* s) { delete s; }
int32_t oracle_scene_num_nodes(const oracle_scene* s) { return s->bvh ? (int32_t)s->bvh->Nodes.size() : 0; }
void oracle_scene_bvh(const oracle_scene* s, izpi_bvh4_node* nodes, int32_t* perm) {
  if (!s->bvh) return;
  std::memcpy(nodes, s->bvh->Nodes.data(), s->bvh->Nodes.size() * sizeof(izpi_bvh4_node));
  std::memcpy(perm, s->bvh->PrimitiveIndices.data(), s->bvh->PrimitiveIndices.size() * sizeof(int32_t));
}
int32_t oracle_scene_num_lights(const oracle_scene* s) { return (int32_t)s->light_ids.size(); }
void oracle_scene_lights(const oracle_scene* s, int32_t* ids) { std::memcpy(ids, s->light_ids.data(), s->light_ids.size() * 4); }
void oracle_prim_bbox(const oracle_scene* s, int32_t prim, double* o) {
  AABB b; s->hitables[prim]->BoundingBox(b);
  o[0] = b.min.X; o[1] = b.min.Y; o[2] = b.min.Z; o[3] = b.max.X; o[4] = b.max.Y; o[5] = b.max.Z;
}
int oracle_triangle_fields(const oracle_scene* s, int32_t prim, double* o) {
  auto* t = dynamic_cast<Triangle*>(s->hitables[prim]);
  if (!t) return -1;
  const Vec3* vs[5] = {&t->edge1, &t->edge2, &t->normal, &t->tangent, &t->bitangent};
  for (int i = 0; i < 5; i++) { o[3 * i] = vs[i]->X; o[3 * i + 1] = vs[i]->Y; o[3 * i + 2] = vs[i]->Z; }
  o[15] = t->area;
  o[16] = t->bb.min.X; o[17] = t->bb.min.Y; o[18] = t->bb.min.Z; o[19] = t->bb.max.X; o[20] = t->bb.max.Y; o[21] = t->bb.max.Z;
  return 0;
}
double oracle_spectral_texture_value(const oracle_scene* s, int32_t i, double lambda) { return s->spectex[i].Value(lambda); }

int32_t oracle_hit(const oracle_scene* s, int flavour, const double org[3], const double dir[3], double tmin, double tmax,
                   double* o) {
  if (s->bvh) s->bvh->flavour = flavour;
  Ray r = NewRay(V(org[0], org[1], org[2]), V(dir[0], dir[1], dir[2]), 0);
  HitRecord rec; const Material* m;
  if (!s->world.Hit(r, tmin, tmax, rec, m)) return -1;
  o[0] = rec.t; o[1] = rec.u; o[2] = rec.v; o[3] = rec.p.X; o[4] = rec.p.Y; o[5] = rec.p.Z;
  o[6] = rec.normal.X; o[7] = rec.normal.Y; o[8] = rec.normal.Z;
  return rec.prim;
}

void oracle_trace(const oracle_scene* s, int flavour, int64_t n, const double* org, const double* dir, double tmin,
                  double tmax, int32_t* prim_id, double* t, oracle_stats* stats, int threads) {
  if (s->bvh) s->bvh->flavour = flavour;
  if (threads < 1) threads = 1;
  std::atomic<int64_t> next(0);
  const int64_t chunk = 4096;  // rays handed out in chunks, like tiles from the work queue (renderer.go:172-188)
  std::vector<Stats> st(threads);
  auto work = [&](int tid) {
    g_stats = stats ? &st[tid] : nullptr;
    for (;;) {
      int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      int64_t e = b + chunk < n ? b + chunk : n;
      for (int64_t i = b; i < e; i++) {
        Ray r = NewRay(V(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), 0);
        HitRecord rec; const Material* m;
        if (s->world.Hit(r, tmin, tmax, rec, m)) { prim_id[i] = rec.prim; t[i] = rec.t; }
        else { prim_id[i] = -1; t[i] = 0.0; }
      }
    }
    g_stats = nullptr;
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (auto& x : st) { stats->nodes += x.nodes; stats->tris += x.tris; stats->spheres += x.spheres; stats->others += x.others; }
    stats->rays = (uint64_t)n;
  }
}

void oracle_tiles(int32_t sx, int32_t sy, int32_t* stepx, int32_t* stepy) {  // common/tiles.go:3-24
  static const int sizes[10] = {32, 25, 24, 20, 16, 12, 10, 8, 5, 4};
  *stepx = 0; *stepy = 0;
  for (int s : sizes) if (sx % s == 0) { *stepx = s; break; }
  for (int s : sizes) if (sy % s == 0) { *stepy = s; break; }
}

// One pixel of renderRectRGB (render/rgb.go:27-41) / RenderPixelSpectral (render/spectral.go:75-106) / the AOV samplers:
// spp samples, mean.  rng / camRng are the work unit's LCGs (mode 0); mode 1 replays the device's counter RNG.
static void render_pixel(const oracle_scene* s, const oracle_render_params* p, const Camera& cam, Sampler& smp, int x, int y, Rng& rng,
                         Rng& camRng, double out[3]) {
  const int nx = p->width, ny = p->height;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int smpl = 0; smpl < p->spp; smpl++) {
    Rng* r = &rng; Rng* cr = &camRng;
    Rng ctr;
    if (p->rng_mode == 1) {
      ctr.mode = 1; ctr.key = Rng::stream_key(p->seed, (uint64_t)y * (uint64_t)nx + (uint64_t)x, (uint64_t)smpl); ctr.ctr = 0;
      r = &ctr; cr = &ctr;
    }
    if (p->sampler >= 2) {  // sampler/albedo.go:30-36, sampler/normal.go:28-34 through the RGB pixel loop
      double u = ((double)x + r->Float64()) / (double)nx;
      double v = ((double)y + r->Float64()) / (double)ny;
      Ray ray = cam.GetRay(u, v, 0, *cr);
      smp.numRays++;
      HitRecord rec; const Material* mat;
      Vec3 c;
      if (s->world.Hit(ray, 0.001, DBL_MAX, rec, mat)) c = p->sampler == 2 ? mat->Albedo(rec.u, rec.v) : rec.normal;
      c = DeNAN(c);
      a0 = a0 + c.X; a1 = a1 + c.Y; a2 = a2 + c.Z;
    } else if (p->sampler == 0) {  // rgb.go:30-37
      double u = ((double)x + r->Float64()) / (double)nx;
      double v = ((double)y + r->Float64()) / (double)ny;
      Ray ray = cam.GetRay(u, v, 0, *cr);
      Vec3 c = DeNAN(smp.Sample(ray, &s->world, &s->lights, 0, *r));
      a0 = a0 + c.X; a1 = a1 + c.Y; a2 = a2 + c.Z;
    } else {  // render/spectral.go:75-96
      double lambda, pdf;
      SampleWavelength(r->Float64(), lambda, pdf);
      if (pdf == 0) continue;
      double u = ((double)x + r->Float64()) / (double)nx;
      double v = ((double)y + r->Float64()) / (double)ny;
      Ray ray = cam.GetRay(u, v, lambda, *cr);
      double radiance = smp.SampleSpectral(ray, &s->world, &s->lights, 0, *r);
      double cx, cy, cz;
      GetCIEValues(lambda, cx, cy, cz);
      a0 += (radiance * cx) / pdf; a1 += (radiance * cy) / pdf; a2 += (radiance * cz) / pdf;
    }
  }
  if (p->sampler != 1) { a0 = a0 / (double)p->spp; a1 = a1 / (double)p->spp; a2 = a2 / (double)p->spp; }  // rgb.go:39
  else { double inv = 1.0 / (double)p->spp; a0 = a0 * inv; a1 = a1 * inv; a2 = a2 * inv; }             // spectral.go:99-103
  out[0] = a0; out[1] = a1; out[2] = a2;
}

static void init_sampler(Sampler& smp, const oracle_render_params* p) {
  smp.maxDepth = p->max_depth;
  smp.background = Vec3();  // colours.Black
  smp.spectralBackground.wavelengths = {380, 750};  // colours.SpectralBlack (colours.go:19): all-zero SPD
  smp.spectralBackground.values = {0, 0};
}

void oracle_render(const oracle_scene* s, const oracle_render_params* p, double* canvas, uint64_t* num_rays) {
  if (s->bvh) s->bvh->flavour = p->flavour;
  g_oracle_libm_jitter = p->libm_jitter;
  const int nx = p->width, ny = p->height;
  int threads = p->threads < 1 ? 1 : p->threads;
  std::atomic<int> nextRow(p->y0);
  std::atomic<uint64_t> rays(0);
  // camera.New with the aspect override W/H (leader.go:45, transport.go:534-539)
  Camera cam = s->camera;
  auto work = [&]() {
    Sampler smp;
    init_sampler(smp, p);
    for (;;) {
      int y = nextRow.fetch_add(1);
      if (y > p->y1) break;
      Rng rng, camRng;  // mode 0: one private LCG per work unit + the camera's own LCG (camera.go:14,46)
      rng.mode = 0; rng.state = Rng::mix64(p->seed + 0x1000003ull * (uint64_t)y) & 0xffffffffull;
      camRng.mode = 0; camRng.state = Rng::mix64(p->seed ^ (0xC0FFEEull + (uint64_t)y)) & 0xffffffffull;
      for (int x = p->x0; x <= p->x1; x++) {
        double a[3];
        render_pixel(s, p, cam, smp, x, y, rng, camRng, a);
        int row = ny - y;  // rgb.go:41: canvas.Set(x, ny-y, ...); row ny is out of bounds -> dropped
        if (row >= 0 && row < ny) {
          double* px = canvas + ((size_t)row * nx + x) * 4;
          px[0] = a[0]; px[1] = a[1]; px[2] = a[2]; px[3] = 1.0;
        }
      }
    }
    rays.fetch_add(smp.numRays);
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  g_oracle_libm_jitter = 0;
  if (num_rays) *num_rays = rays.load();
  if (p->sampler == 1 && p->epilogue) {  // renderer.go:216-219
    oracle_firefly_rejection(canvas, nx, ny);
    std::vector<double> tmp(canvas, canvas + (size_t)4 * nx * ny);
    oracle_xyz_to_rgb(tmp.data(), canvas, nx, ny, cam.exposure);
  }
}

// worker.RenderTile (internal/worker/render.go:17-75): one LCG from the pool for the whole tile (:24), per image row a fresh
// strip of stripSize doubles (:35) whose first 4*width entries receive the pixel means and alpha 1 (:51-55); rows are streamed in
// image order, y0 first, without the local path's row flip, and y == 0 is streamed like any other row.
void oracle_render_tile(const oracle_scene* s, const oracle_render_params* p, uint32_t strip_height, uint32_t x0, uint32_t y0,
                        uint32_t x1, uint32_t y1, double* rows, uint64_t* num_rays) {
  if (s->bvh) s->bvh->flavour = p->flavour;
  g_oracle_libm_jitter = p->libm_jitter;
  const size_t stripSize = (size_t)strip_height * 4 * (x1 - x0 + 1);
  Camera cam = s->camera;
  Sampler smp;
  init_sampler(smp, p);
  Rng rng, camRng;
  rng.mode = 0; rng.state = Rng::mix64(p->seed + 0x1000003ull * (uint64_t)y0 + x0) & 0xffffffffull;
  camRng.mode = 0; camRng.state = Rng::mix64(p->seed ^ (0xC0FFEEull + (uint64_t)y0)) & 0xffffffffull;
  for (uint32_t y = y0; y <= y1; y++) {
    double* pixels = rows + (size_t)(y - y0) * stripSize;
    for (size_t k = 0; k < stripSize; k++) pixels[k] = 0.0;  // make([]float64, stripSize)
    size_t i = 0;
    for (uint32_t x = x0; x <= x1; x++) {
      double a[3];
      render_pixel(s, p, cam, smp, (int)x, (int)y, rng, camRng, a);
      pixels[i] = a[0]; pixels[i + 1] = a[1]; pixels[i + 2] = a[2]; pixels[i + 3] = 1.0;
      i += 4;
    }
  }
  g_oracle_libm_jitter = 0;
  if (num_rays) *num_rays = smp.numRays;
}

void oracle_firefly_rejection(double* pix, int32_t width, int32_t height) {  // firefly_rejection.go:12-113
  if (width == 0 || height == 0) return;
  const int kernelRadius = 1; const double kThreshold = 2.5; const int minNeighbors = 3;
  std::vector<double> yValues((size_t)width * height);
  for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) yValues[(size_t)y * width + x] = pix[((size_t)y * width + x) * 4 + 1];
  for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {
    size_t pixelIdx = ((size_t)y * width + x) * 4;
    double currentY = yValues[(size_t)y * width + x];
    if (currentY <= 0) continue;
    double nb[8]; int n = 0;
    for (int dy = -kernelRadius; dy <= kernelRadius; dy++) for (int dx = -kernelRadius; dx <= kernelRadius; dx++) {
      if (dx == 0 && dy == 0) continue;
      int nx = x + dx, ny = y + dy;
      if (nx >= 0 && nx < width && ny >= 0 && ny < height) {
        double v = yValues[(size_t)ny * width + nx];
        if (v > 0) nb[n++] = v;
      }
    }
    if (n < minNeighbors) continue;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += nb[i];
    double mean = sum / (double)n;
    double varianceSum = 0.0;
    for (int i = 0; i < n; i++) { double d = nb[i] - mean; varianceSum += d * d; }
    double stddev = std::sqrt(varianceSum / (double)n);
    double threshold = mean + kThreshold * stddev;
    if (currentY > threshold && threshold > 0) {
      double ratio = threshold / currentY;
      pix[pixelIdx] *= ratio; pix[pixelIdx + 1] *= ratio; pix[pixelIdx + 2] *= ratio;
    }
  }
}

void oracle_xyz_to_rgb(const double* in, double* out, int32_t width, int32_t height, double exposure) {  // rgb_image.go:13-67
  static const double M[3][3] = {{1.6410234, -0.3248033, -0.2364247}, {-0.6636629, 1.6153316, 0.0167563}, {0.0117219, -0.0082845, 0.9883949}};
  for (size_t i = 0; i < (size_t)width * height; i++) {
    double x = in[4 * i] * exposure, y = in[4 * i + 1] * exposure, z = in[4 * i + 2] * exposure;
    out[4 * i] = M[0][0] * x + M[0][1] * y + M[0][2] * z;
    out[4 * i + 1] = M[1][0] * x + M[1][1] * y + M[1][2] * z;
    out[4 * i + 2] = M[2][0] * x + M[2][1] * y + M[2][2] * z;
    out[4 * i + 3] = in[4 * i + 3];
  }
}

}  // extern "C"

// ---- design aid: per-ray step sequence of the exact traversal (N = inner node visit, L = leaf-node visit whose
// box test passes and primitives are tested, F = leaf-node visit whose box test fails).  Used by
// profiles/sim_schedule.py to evaluate SIMT scheduling policies offline; not part of any parity check.
extern "C" int64_t oracle_trace_steps(const oracle_scene* s, int64_t n, const double* org, const double* dir, double tmin,
                                      double tmax, char* out, int64_t out_cap, int64_t* offsets) {
  const BVH4* bvh = s->bvh.get();
  if (!bvh) return -1;
  int64_t pos = 0;
  for (int64_t i = 0; i < n; i++) {
    offsets[i] = pos;
    Ray r = NewRay(V(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), 0);
    float inv[3] = {(float)(1.0 / r.direction.X), (float)(1.0 / r.direction.Y), (float)(1.0 / r.direction.Z)};
    float o[3] = {(float)r.origin.X, (float)r.origin.Y, (float)r.origin.Z};
    int32_t stack[64]; int sp = 0; int32_t cur = 0; double tM = tmax;
    HitRecord rec; const Material* m;
    while (cur != -1) {
      const izpi_bvh4_node& node = bvh->Nodes[cur];
      uint8_t mask = RayAABB4(BOX_SSE, o, inv, node, (float)tM);
      bool leafnode = node.primitive_count[0] > 0;
      if (pos < out_cap) out[pos++] = leafnode ? ((mask & 1) ? 'L' : 'F') : 'N';
      int32_t next = -1;
      for (int k = 0; k < 4; k++) {
        if (!((mask >> k) & 1) || node.child_index[k] == -1) continue;
        if (node.primitive_count[k] > 0) {
          for (int p = 0; p < node.primitive_count[k]; p++)
            if (bvh->Primitives[node.child_index[k] + p]->Hit(r, tmin, tM, rec, m)) tM = rec.t;
        } else if (next == -1) next = node.child_index[k];
        else stack[sp++] = node.child_index[k];
      }
      cur = next != -1 ? next : (sp > 0 ? stack[--sp] : -1);
    }
  }
  offsets[n] = pos;
  return pos;
}

// ---- replay of material/dielectric_test.go:47-216 (mockSceneGeometry, LCG(12345), 1000 scatter calls) -------------
namespace {
struct MockSceneGeometry : Hitable {  // dielectric_test.go:73-84: always "hits" at (0.5, 0.5, 1.5)
  bool Hit(const Ray&, double, double, HitRecord& rec, const Material*& mat) const override {
    rec.t = 2.0; rec.u = 0; rec.v = 0; rec.p = V(0.5, 0.5, 1.5); rec.normal = V(0, 0, 1); mat = nullptr;
    return true;
  }
  bool BoundingBox(AABB&) const override { return false; }
  bool IsEmitter() const override { return false; }
};
}  // namespace

extern "C" {
// kind 0: TestColoredGlassScattering (RGB, NewColoredDielectric(1.5, (0.1,0.2,0.3)))
// kind 1: TestSpectralColoredGlassScattering (gaussian refidx (1.5,550,50), gaussian absorption (0.5,480,60), lambda 480)
// kind 2: TestPathLengthCalculation
// out: [0] iterations run, [1] found attenuated, [2] found unattenuated, [3..5] first attenuated value(s), [6] path length
void oracle_probe_dielectric_test(int kind, double* out) {
  MockSceneGeometry mock;
  SpectralTexture refidx, absorb;
  refidx.type = IZPI_SPEC_GAUSSIAN; refidx.peak = 1.5; refidx.centre = 550.0; refidx.width = 50.0;
  absorb.type = IZPI_SPEC_GAUSSIAN; absorb.peak = 0.5; absorb.centre = 480.0; absorb.width = 60.0;
  Material m;
  m.type = IZPI_MAT_DIELECTRIC; m.world = &mock;
  for (int i = 0; i < 8; i++) out[i] = 0;
  if (kind == 2) {
    m.s = 1.5;
    HitRecord hr; hr.t = 1.0; hr.p = V(0.5, 0.5, 0.5); hr.normal = V(0.577, 0.577, 0.577);
    Ray r = NewRay(V(0, 0, -2), V(0, 0, 1), 0), scattered = r;
    out[6] = m.calculatePathLength(r, hr, scattered);
    return;
  }
  if (kind == 0) { m.s = 1.5; m.v = V(0.1, 0.2, 0.3); m.computeBeerLambert = true; }
  else { m.spectral = &refidx; m.spectralAbsorption = &absorb; }
  HitRecord hr; hr.t = 1.0; hr.p = V(0, 0, 1); hr.normal = V(0, 0, 1);
  Ray r = NewRay(V(0, 0, -1), V(0, 0, 1), 0, kind == 1 ? 480.0 : 0.0);
  Rng rng; rng.mode = 0; rng.state = 12345;  // fastrandom.New(12345, 4294967296, 1664525, 1013904223)
  bool foundA = false, foundN = false;
  int it = 0;
  for (; it < 1000; it++) {
    ScatterRecord s;
    if (kind == 0) {
      m.Scatter(r, hr, rng, s);
      const double eps = 1e-6;
      if (s.attenuation.X < 1.0 - eps || s.attenuation.Y < 1.0 - eps || s.attenuation.Z < 1.0 - eps) {
        if (!foundA) { out[3] = s.attenuation.X; out[4] = s.attenuation.Y; out[5] = s.attenuation.Z; }
        foundA = true;
      }
      if (s.attenuation.X >= 1.0 - eps && s.attenuation.Y >= 1.0 - eps && s.attenuation.Z >= 1.0 - eps) foundN = true;
    } else {
      m.SpectralScatter(r, hr, rng, s);
      if (s.spectralAtten < 1.0) { if (!foundA) out[3] = s.spectralAtten; foundA = true; }
      if (s.spectralAtten >= 1.0) foundN = true;
    }
    if (foundA && foundN) { it++; break; }
  }
  out[0] = it; out[1] = foundA; out[2] = foundN;
}
}

// ---- displacement.ApplyDisplacementMap (internal/displacement/displacement.go:36-280) ------------------------------
namespace {
struct MinimalTriangle {  // displacement.go:19-33
  Vec3 vertex0, vertex1, vertex2;
  int32_t material;
  double u0, u1, u2, v0, v1, v2;
};
void tessellate(const MinimalTriangle& in, MinimalTriangle out[4]) {  // displacement.go:36-103
  Vec3 a = ScalarDiv(Add(in.vertex0, in.vertex1), 2.0), b = ScalarDiv(Add(in.vertex1, in.vertex2), 2.0), c = ScalarDiv(Add(in.vertex2, in.vertex0), 2.0);
  double ua = (in.u0 + in.u1) / 2.0, va = (in.v0 + in.v1) / 2.0, ub = (in.u1 + in.u2) / 2.0, vb = (in.v1 + in.v2) / 2.0;
  double uc = (in.u2 + in.u0) / 2.0, vc = (in.v2 + in.v0) / 2.0;
  out[0] = {in.vertex0, a, c, in.material, in.u0, ua, uc, in.v0, va, vc};
  out[1] = {a, b, c, in.material, ua, ub, uc, va, vb, vc};
  out[2] = {a, in.vertex1, b, in.material, ua, in.u1, ub, va, in.v1, vb};
  out[3] = {c, b, in.vertex2, in.material, uc, ub, in.u2, vc, vb, in.v2};
}
bool isTessellatedEnough(const MinimalTriangle& t, double maxDeltaU, double maxDeltaV, const Texture& map, double mn, double mx,
                         double adaptiveThreshold) {  // displacement.go:120-141
  bool uv = std::fabs(t.u1 - t.u0) <= maxDeltaU && std::fabs(t.u2 - t.u1) <= maxDeltaU && std::fabs(t.u0 - t.u2) <= maxDeltaU &&
            std::fabs(t.v1 - t.v0) <= maxDeltaV && std::fabs(t.v2 - t.v1) <= maxDeltaV && std::fabs(t.v0 - t.v2) <= maxDeltaV;
  if (!uv) return false;
  double d0 = map.Value(t.u0, t.v0).Z, d1 = map.Value(t.u1, t.v1).Z, d2 = map.Value(t.u2, t.v2).Z;  // :106-118
  double variation = gomax(d0, gomax(d1, d2)) - gomin(d0, gomin(d1, d2));
  return variation * std::fabs(mx - mn) <= adaptiveThreshold;
}
void displaceOne(const MinimalTriangle& t, const Texture& map, double mn, double mx, double out15[15]) {  // displacement.go:201-280
  Vec3 edge1 = Sub(t.vertex1, t.vertex0), edge2 = Sub(t.vertex2, t.vertex0);
  Vec3 normal = MakeUnitVector(Cross(edge1, edge2));
  double dU1 = t.u1 - t.u0, dU2 = t.u2 - t.u0, dV1 = t.v1 - t.v0, dV2 = t.v2 - t.v0;
  double f = 1.0 / (dU1 * dV2 - dU2 * dV1);
  Vec3 tangent = MakeUnitVector(V(f * (dV2 * edge1.X - dV1 * edge2.X), f * (dV2 * edge1.Y - dV1 * edge2.Y), f * (dV2 * edge1.Z - dV1 * edge2.Z)));
  Vec3 bitangent = MakeUnitVector(V(f * (-dU2 * edge1.X + dU1 * edge2.X), f * (-dU2 * edge1.Y + dU1 * edge2.Y), f * (-dU2 * edge1.Z + dU1 * edge2.Z)));
  auto disp = [&](const Vec3& vtx, double u, double v) {
    double z = mn + ((mx - mn) * map.Value(u, v).Z);
    // mat3.MatrixVectorMul(tbn, (0, 0, z)) (mat3.go:34): every row keeps its three products
    Vec3 d = V(tangent.X * 0.0 + bitangent.X * 0.0 + normal.X * z, tangent.Y * 0.0 + bitangent.Y * 0.0 + normal.Y * z,
               tangent.Z * 0.0 + bitangent.Z * 0.0 + normal.Z * z);
    return Add(vtx, d);
  };
  Vec3 p0 = disp(t.vertex0, t.u0, t.v0), p1 = disp(t.vertex1, t.u1, t.v1), p2 = disp(t.vertex2, t.u2, t.v2);
  double o[15] = {p0.X, p0.Y, p0.Z, p1.X, p1.Y, p1.Z, p2.X, p2.Y, p2.Z, t.u0, t.v0, t.u1, t.v1, t.u2, t.v2};
  std::memcpy(out15, o, sizeof(o));
}
// applyTessellation (displacement.go:188-218)
std::vector<MinimalTriangle> applyTessellation(std::vector<MinimalTriangle> in, double maxDeltaU, double maxDeltaV, const Texture& map, double mn,
                                               double mx, double adaptiveThreshold) {
  std::vector<MinimalTriangle> done;
  int level = 0;
  while (!in.empty()) {
    // (the reference loops forever when two ADJACENT texels differ by more than threshold/|max-min|: a triangle
    // straddling them never passes; such maps are invalid inputs and are cut off here)
    if (++level > 40 || in.size() > (1u << 25)) { in.clear(); break; }
    std::vector<MinimalTriangle> toIn;
    for (const MinimalTriangle& t : in) {
      MinimalTriangle ch[4];
      tessellate(t, ch);
      for (int k = 0; k < 4; k++) (isTessellatedEnough(ch[k], maxDeltaU, maxDeltaV, map, mn, mx, adaptiveThreshold) ? done : toIn).push_back(ch[k]);
    }
    in.swap(toIn);
  }
  return done;
}
void applyOne(std::vector<MinimalTriangle> in, const Texture& map, double mn, double mx, std::vector<double>& out, std::vector<int32_t>& mats) {
  double maxDeltaU = 4.0 / (double)(map.sizeX - 1), maxDeltaV = 4.0 / (double)(map.sizeY - 1);  // displacement.go:176-178
  const double adaptiveThreshold = 2.0;
  for (const MinimalTriangle& t : applyTessellation(std::move(in), maxDeltaU, maxDeltaV, map, mn, mx, adaptiveThreshold)) {
    double o[15];
    displaceOne(t, map, mn, mx, o);
    out.insert(out.end(), o, o + 15);
    mats.push_back(t.material);
  }
}
MinimalTriangle mt_from15(const double* p, int32_t material) {
  return MinimalTriangle{V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), material, p[9], p[11], p[13], p[10], p[12], p[14]};
}
void mt_to15(const MinimalTriangle& t, double* o) {
  const double v[15] = {t.vertex0.X, t.vertex0.Y, t.vertex0.Z, t.vertex1.X, t.vertex1.Y, t.vertex1.Z, t.vertex2.X, t.vertex2.Y, t.vertex2.Z,
                        t.u0, t.v0, t.u1, t.v1, t.u2, t.v2};
  std::memcpy(o, v, sizeof(v));
}
}  // namespace

extern "C" {
// tris: n x 15 doubles (v0 v1 v2 u0 v0 u1 v1 u2 v2).  per_triangle != 0 = the transport call shape: one
// ApplyDisplacementMap call per input triangle, results concatenated (transport.go:633-646).
// Returns the number of output triangles; *out_tris / *out_mats are malloc'ed (free with oracle_free).
int64_t oracle_apply_displacement(int64_t n, const double* tris, const int32_t* mats, int32_t w, int32_t h, const double* pixels,
                                  double mn, double mx, int per_triangle, double** out_tris, int32_t** out_mats) {
  Texture map; map.type = IZPI_TEX_IMAGE; map.sizeX = w; map.sizeY = h; map.pixels = pixels;
  std::vector<double> out; std::vector<int32_t> om;
  auto mk1 = [&](int64_t i) {
    const double* p = tris + 15 * i;
    return MinimalTriangle{V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), mats ? mats[i] : 0, p[9], p[11], p[13], p[10], p[12], p[14]};
  };
  if (per_triangle) { for (int64_t i = 0; i < n; i++) applyOne({mk1(i)}, map, mn, mx, out, om); }
  else { std::vector<MinimalTriangle> in; for (int64_t i = 0; i < n; i++) in.push_back(mk1(i)); applyOne(in, map, mn, mx, out, om); }
  *out_tris = (double*)std::malloc(out.size() * sizeof(double) + 8);
  *out_mats = (int32_t*)std::malloc(om.size() * sizeof(int32_t) + 8);
  std::memcpy(*out_tris, out.data(), out.size() * sizeof(double));
  std::memcpy(*out_mats, om.data(), om.size() * sizeof(int32_t));
  return (int64_t)om.size();
}
void oracle_free(void* p) { std::free(p); }

// ---- the internal steps, exposed so that the reference's own unit vectors (displacement_test.go) can be replayed -----------
// tessellate() (displacement.go:36-103): one triangle (15 doubles) -> four
void oracle_displacement_tessellate(const double* in15, double* out60) {
  MinimalTriangle ch[4];
  tessellate(mt_from15(in15, 0), ch);
  for (int k = 0; k < 4; k++) mt_to15(ch[k], out60 + 15 * k);
}
// applyTessellation() with explicit limits over a CONSTANT texture (displacement_test.go:84-157); returns the triangle count
int64_t oracle_displacement_apply_tessellation(int64_t n, const double* tris15, double maxDeltaU, double maxDeltaV, const double* rgb,
                                               double mn, double mx, double adaptiveThreshold) {
  Texture map; map.type = IZPI_TEX_CONSTANT; map.color = V(rgb[0], rgb[1], rgb[2]);
  std::vector<MinimalTriangle> in;
  for (int64_t i = 0; i < n; i++) in.push_back(mt_from15(tris15 + 15 * i, 0));
  return (int64_t)applyTessellation(std::move(in), maxDeltaU, maxDeltaV, map, mn, mx, adaptiveThreshold).size();
}
// applyDisplacement() over a CONSTANT texture, no tessellation (displacement_test.go:159-213): n triangles in, n out
void oracle_displacement_apply_displacement(int64_t n, const double* tris15, const double* rgb, double mn, double mx, double* out15) {
  Texture map; map.type = IZPI_TEX_CONSTANT; map.color = V(rgb[0], rgb[1], rgb[2]);
  for (int64_t i = 0; i < n; i++) displaceOne(mt_from15(tris15 + 15 * i, 0), map, mn, mx, out15 + 15 * i);
}
}
