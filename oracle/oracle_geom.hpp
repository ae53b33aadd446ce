// oracle_geom.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of Izpi's geometry layer (reference = /root/reference, pure Go).  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this; the product (izpi_b200/) never does.
//
// Every function cites the reference file:line it follows.  Arithmetic is IEEE fp64 with
// the reference's operation order and NO fused multiply-add (build with
// -ffp-contract=off; Go/amd64 never fuses).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <limits>
#include <memory>
#include <vector>

#include "../include/izpi_scene.h"

// ---- libm sensitivity probe (tests only) -------------------------------------------------------------------------------
// The device's sin / cos / pow / exp / atan2 / asin are not correctly rounded (CUDA documents 1-2 ulp), glibc's mostly are.
// A last-bit difference in a scatter direction is invisible at 1e-7 -- unless the path is chaotic (a chain of bounces
// between glass spheres multiplies a perturbation by ~10-100x per bounce) or a comparison sits within an ulp of its
// threshold.  To CLASSIFY the pixels on which the device and the oracle may legitimately differ, the oracle can perturb
// every libm result on the path by up to +-2 ulp, pseudo-randomly: pixels whose value survives that are stable, and on
// those the device must agree (tests/test_render_gpu.py).  0 = off (the default, and the only mode that restates the reference).
extern int g_oracle_libm_jitter;
inline double LM(double v) {
  if (g_oracle_libm_jitter == 0 || !(v == v) || v == 0.0 || std::isinf(v)) return v;
  uint64_t b;
  std::memcpy(&b, &v, 8);
  uint64_t h = (b ^ (uint64_t)g_oracle_libm_jitter * 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
  h ^= h >> 29;
  int k = (int)(h % 5) - 2;  // -2 .. +2 ulp
  b += (uint64_t)(int64_t)k;
  std::memcpy(&v, &b, 8);
  return v;
}

namespace orc {

// ---------------------------------------------------------------- vec3 (vec3/vec3.go)
struct Vec3 {
  double X = 0, Y = 0, Z = 0;
};
inline Vec3 V(double x, double y, double z) { Vec3 v; v.X = x; v.Y = y; v.Z = z; return v; }
inline double Length(const Vec3& v) { return std::sqrt((v.X * v.X) + (v.Y * v.Y) + (v.Z * v.Z)); }   // vec3.go:17
inline double SquaredLength(const Vec3& v) { return (v.X * v.X) + (v.Y * v.Y) + (v.Z * v.Z); }        // vec3.go:22
inline Vec3 MakeUnitVector(Vec3 v) { double l = Length(v); v.X = v.X / l; v.Y = v.Y / l; v.Z = v.Z / l; return v; }  // vec3.go:27
inline Vec3 Add(const Vec3& a, const Vec3& b) { return V(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }          // vec3.go:37
inline Vec3 Add(const Vec3& a, const Vec3& b, const Vec3& c) { return Add(Add(a, b), c); }
inline Vec3 Sub(const Vec3& a, const Vec3& b) { return V(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }          // vec3.go:50
inline Vec3 Sub(const Vec3& a, const Vec3& b, const Vec3& c) { return Sub(Sub(a, b), c); }
inline Vec3 Sub(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d) { return Sub(Sub(Sub(a, b), c), d); }
inline Vec3 Mul(const Vec3& a, const Vec3& b) { return V(a.X * b.X, a.Y * b.Y, a.Z * b.Z); }          // vec3.go:63
inline Vec3 ScalarMul(const Vec3& a, double t) { return V(a.X * t, a.Y * t, a.Z * t); }               // vec3.go:81
inline Vec3 ScalarDiv(const Vec3& a, double t) { return V(a.X / t, a.Y / t, a.Z / t); }               // vec3.go:90
inline double Dot(const Vec3& a, const Vec3& b) { return (a.X * b.X) + (a.Y * b.Y) + (a.Z * b.Z); }   // vec3.go:99
inline Vec3 Cross(const Vec3& a, const Vec3& b) {                                                     // vec3.go:104
  return V((a.Y * b.Z) - (a.Z * b.Y), -((a.X * b.Z) - (a.Z * b.X)), (a.X * b.Y) - (a.Y * b.X));
}
inline Vec3 UnitVector(const Vec3& v) { return ScalarDiv(v, Length(v)); }                             // vec3.go:113
inline Vec3 Lerp(const Vec3& a, const Vec3& b, double t) {                                            // vec3.go:254
  return V((1 - t) * a.X + t * b.X, (1 - t) * a.Y + t * b.Y, (1 - t) * a.Z + t * b.Z);
}
inline Vec3 DeNAN(const Vec3& v) {                                                                     // vec3.go:141
  auto f = [](double x) { return (std::isnan(x) || std::isinf(x)) ? 0.0 : x; };
  return V(f(v.X), f(v.Y), f(v.Z));
}
inline Vec3 Min3(const Vec3& a, const Vec3& b, const Vec3& c) {                                        // vec3.go:162
  double x = DBL_MAX, y = DBL_MAX, z = DBL_MAX;
  if (a.X < x) x = a.X; if (b.X < x) x = b.X; if (c.X < x) x = c.X;
  if (a.Y < y) y = a.Y; if (b.Y < y) y = b.Y; if (c.Y < y) y = c.Y;
  if (a.Z < z) z = a.Z; if (b.Z < z) z = b.Z; if (c.Z < z) z = c.Z;
  return V(x, y, z);
}
inline Vec3 Max3(const Vec3& a, const Vec3& b, const Vec3& c) {                                        // vec3.go:207
  double x = -DBL_MAX, y = -DBL_MAX, z = -DBL_MAX;
  if (a.X > x) x = a.X; if (b.X > x) x = b.X; if (c.X > x) x = c.X;
  if (a.Y > y) y = a.Y; if (b.Y > y) y = b.Y; if (c.Y > y) y = c.Y;
  if (a.Z > z) z = a.Z; if (b.Z > z) z = b.Z; if (c.Z > z) z = c.Z;
  return V(x, y, z);
}
// Go's math.Min / math.Max for the non-NaN, non-signed-zero cases that occur here.
inline double gomin(double a, double b) { return a < b ? a : b; }
inline double gomax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------- RNG
// mode 0: the reference's fastrandom.LCG (fastrandom.go:41-47).
// mode 1: the device's counter generator (izpi_b200/csrc/device/rng.cuh), so that oracle
//         and device walk the SAME paths on the same (pixel, sample) key.
struct Rng {
  int mode = 0;
  uint64_t state = 0;  // LCG state
  uint64_t key = 0;    // counter mode: stream key
  uint32_t ctr = 0;    // counter mode: draw index
  static uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  static uint64_t stream_key(uint64_t seed, uint64_t pixel, uint64_t sample) {
    return mix64(mix64(seed ^ 0x9E3779B97F4A7C15ull) + pixel * 0xD1B54A32D192ED03ull + sample * 0x8CB92BA72F3D8DD7ull);
  }
  double Float64() {
    if (mode == 0) {
      state = (1664525ull * state + 1013904223ull) % 4294967296ull;
      return (double)state / 4294967296.0;
    }
    uint64_t z = mix64(key + (uint64_t)(++ctr) * 0x9E3779B97F4A7C15ull);
    return (double)(uint32_t)(z >> 32) / 4294967296.0;  // same 2^-32 granularity as the LCG
  }
};

// ---------------------------------------------------------------- ray / hit record
struct Ray {  // ray/ray.go:9
  Vec3 origin, direction;
  double lambda = 0, time = 0;
  Vec3 PointAtParameter(double t) const { return Add(origin, ScalarMul(direction, t)); }  // ray.go:45
};
inline Ray NewRay(const Vec3& o, const Vec3& d, double time, double lambda = 0) {
  Ray r; r.origin = o; r.direction = d; r.time = time; r.lambda = lambda; return r;
}

struct Material;  // oracle_shade.hpp

struct HitRecord {  // hitrecord/hitrecord.go:6
  double u = 0, v = 0, t = 0;
  Vec3 p, normal;
  int32_t prim = -1;  // index of the top-level hitable in construction order (oracle bookkeeping)
};

struct AABB {  // aabb/aabb.go:12
  Vec3 min, max;
};
inline AABB SurroundingBox(const AABB& a, const AABB& b) {  // aabb.go:26
  AABB r;
  r.min = V(gomin(a.min.X, b.min.X), gomin(a.min.Y, b.min.Y), gomin(a.min.Z, b.min.Z));
  r.max = V(gomax(a.max.X, b.max.X), gomax(a.max.Y, b.max.Y), gomax(a.max.Z, b.max.Z));
  return r;
}

// per-thread traversal counters (algorithmic bytes per ray)
struct Stats {
  uint64_t nodes = 0, tris = 0, spheres = 0, others = 0, rays = 0;
};
extern thread_local Stats* g_stats;

// ---------------------------------------------------------------- Hitable (hitable/api.go:14)
struct Texture;
struct Hitable {
  int32_t prim_id = -1;
  virtual ~Hitable() {}
  virtual bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const = 0;
  virtual bool BoundingBox(AABB& box) const = 0;
  virtual double PDFValue(const Vec3& o, const Vec3& v) const { return 0.0; }
  virtual Vec3 Random(const Vec3& o, Rng& rng) const { return V(1, 0, 0); }
  virtual bool IsEmitter() const = 0;
};

bool MaterialIsEmitter(const Material* m);
const Texture* MaterialNormalMap(const Material* m);
Vec3 TextureValue(const Texture* t, double u, double v, const Vec3& p);

// ---------------------------------------------------------------- Triangle (hitable/triangle.go)
struct Triangle : Hitable {
  Vec3 vertex0, vertex1, vertex2, edge1, edge2, normal, tangent, bitangent;
  double area = 0;
  const Material* material = nullptr;
  double u0 = 0, u1 = 0, u2 = 0, v0 = 0, v1 = 0, v2 = 0;
  AABB bb;

  // NewTriangleWithUV (triangle.go:60-70) -> NewTriangleWithUVAndNormal (:73-134)
  Triangle(const Vec3& a, const Vec3& b, const Vec3& c, double u0_, double v0_, double u1_, double v1_,
           double u2_, double v2_, const Material* m) {
    Vec3 e1 = Sub(b, a), e2 = Sub(c, a);
    normal = MakeUnitVector(Cross(e1, e2));
    double deltaU1 = u1_ - u0_, deltaU2 = u2_ - u0_, deltaV1 = v1_ - v0_, deltaV2 = v2_ - v0_;
    edge1 = Sub(b, a);
    edge2 = Sub(c, a);
    Vec3 n = Cross(edge1, edge2);
    area = Length(n) / 2.0;
    double f = 1.0 / (deltaU1 * deltaV2 - deltaU2 * deltaV1);
    tangent = MakeUnitVector(V(f * (deltaV2 * edge1.X - deltaV1 * edge2.X), f * (deltaV2 * edge1.Y - deltaV1 * edge2.Y),
                               f * (deltaV2 * edge1.Z - deltaV1 * edge2.Z)));
    bitangent = MakeUnitVector(V(f * (-deltaU2 * edge1.X + deltaU1 * edge2.X), f * (-deltaU2 * edge1.Y + deltaU1 * edge2.Y),
                                 f * (-deltaU2 * edge1.Z + deltaU1 * edge2.Z)));
    Vec3 mn = Min3(a, b, c), mx = Max3(a, b, c);
    Vec3 size = Sub(mx, mn);
    double maxDim = gomax(size.X, gomax(size.Y, size.Z));
    double epsilon = gomax(maxDim * 1e-4, 1e-6);
    Vec3 delta = V(epsilon, epsilon, epsilon);
    bb.min = Sub(mn, delta);
    bb.max = Add(mx, delta);
    vertex0 = a; vertex1 = b; vertex2 = c;
    material = m;
    u0 = u0_; u1 = u1_; u2 = u2_; v0 = v0_; v1 = v1_; v2 = v2_;
  }

  // triangle.go:193-265
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {
    if (g_stats) g_stats->tris++;
    const double epsilon = 1e-8;
    Vec3 h = Cross(r.direction, edge2);
    double a = Dot(edge1, h);
    if (std::fabs(a) < epsilon) return false;
    double f = 1.0 / a;
    Vec3 s = Sub(r.origin, vertex0);
    double u = f * Dot(s, h);
    if (u < -epsilon || u > 1.0 + epsilon) return false;
    Vec3 q = Cross(s, edge1);
    double v = f * Dot(r.direction, q);
    if (v < -epsilon || u + v > 1.0 + epsilon) return false;
    double t = f * Dot(edge2, q);
    if (t < tMin || t > tMax) return false;
    double w = 1.0 - u - v;
    double sum = u + v + w;
    if (std::fabs(sum - 1.0) > epsilon) { u /= sum; v /= sum; w /= sum; }
    double uu = w * u0 + u * u1 + v * u2;
    double vv = w * v0 + u * v1 + v * v2;
    Vec3 nrm = normal;  // the protobuf path never sets per-vertex normals (transport.go:631)
    rec.t = t; rec.u = uu; rec.v = vv; rec.p = r.PointAtParameter(t); rec.prim = prim_id;
    mat = material;
    const Texture* normalMap = MaterialNormalMap(material);
    if (!normalMap) { rec.normal = nrm; return true; }
    Vec3 nts = TextureValue(normalMap, uu, vv, Vec3());   // triangle.go:250-264
    nts.X = 2 * nts.X - 1.0; nts.Y = 2 * nts.Y - 1.0; nts.Z = 2 * nts.Z - 1.0;
    // mat3.NewTBN / MatrixVectorMul (mat3/mat3.go:19,34)
    Vec3 nn = V(tangent.X * nts.X + bitangent.X * nts.Y + nrm.X * nts.Z,
                tangent.Y * nts.X + bitangent.Y * nts.Y + nrm.Y * nts.Z,
                tangent.Z * nts.X + bitangent.Z * nts.Y + nrm.Z * nts.Z);
    rec.normal = MakeUnitVector(nn);
    return true;
  }
  bool BoundingBox(AABB& box) const override { box = bb; return true; }
  double PDFValue(const Vec3& o, const Vec3& v) const override {  // triangle.go:271-280
    Ray r = NewRay(o, v, 0);
    HitRecord rec; const Material* m;
    if (Hit(r, 0.001, DBL_MAX, rec, m)) {
      double distanceSquared = rec.t * rec.t * SquaredLength(v);
      double cosine = std::fabs(Dot(v, ScalarDiv(rec.normal, Length(v))));
      return distanceSquared / (cosine * area);
    }
    return 0;
  }
  Vec3 Random(const Vec3& o, Rng& rng) const override {  // triangle.go:317-326
    double t1 = rng.Float64();
    Vec3 p01 = Lerp(vertex0, vertex1, t1);
    double t2 = rng.Float64();
    Vec3 p02 = Lerp(vertex0, vertex2, t2);
    double t3 = rng.Float64();
    return Sub(Lerp(p01, p02, t3), o);
  }
  bool IsEmitter() const override { return MaterialIsEmitter(material); }
};

// ---------------------------------------------------------------- Sphere (hitable/sphere.go)
inline void getSphereUV(const Vec3& p, double& u, double& v) {  // sphere.go:29-35
  double phi = LM(std::atan2(p.Z, p.X));
  double theta = LM(std::asin(p.Y));
  u = 1.0 - (phi + M_PI) / (2.0 * M_PI);
  v = (theta + M_PI / 2.0) / M_PI;
}
struct ONB {  // onb/onb.go
  Vec3 u, v, w;
  void BuildFromW(const Vec3& n) {  // onb.go:38-52
    w = UnitVector(n);
    Vec3 a = (std::fabs(w.X) > 0.9) ? V(0, 1, 0) : V(1, 0, 0);
    v = UnitVector(Cross(w, a));
    u = Cross(w, v);
  }
  Vec3 Local(const Vec3& a) const { return Add(ScalarMul(u, a.X), ScalarMul(v, a.Y), ScalarMul(w, a.Z)); }  // onb.go:63
};
inline Vec3 RandomToSphere(double radius, double distanceSquared, Rng& rng) {  // vec3.go:130-138
  double r1 = rng.Float64(), r2 = rng.Float64();
  double z = 1 + r2 * (std::sqrt(1 - radius * radius / distanceSquared) - 1);
  double phi = 2 * M_PI * r1;
  double x = LM(std::cos(phi)) * std::sqrt(1 - z * z);
  double y = LM(std::sin(phi)) * std::sqrt(1 - z * z);
  return V(x, y, z);
}
struct Sphere : Hitable {
  Vec3 center0, center1;
  double time0 = 0, time1 = 1, radius = 0;
  const Material* material = nullptr;
  Vec3 center(double time) const {  // sphere.go:125-127
    return Add(center0, ScalarMul(Sub(center1, center0), ((time - time0) / (time1 - time0))));
  }
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {  // sphere.go:63-95
    if (g_stats) g_stats->spheres++;
    Vec3 oc = Sub(r.origin, center(r.time));
    double a = Dot(r.direction, r.direction);
    double b = Dot(oc, r.direction);
    double c = Dot(oc, oc) - (radius * radius);
    double discriminant = (b * b) - (a * c);
    if (discriminant > 0) {
      double temp = (-b - std::sqrt(b * b - a * c)) / a;
      if (temp < tMax && temp > tMin) {
        Vec3 on = ScalarDiv(Sub(r.PointAtParameter(temp), center(r.time)), radius);
        if (Dot(r.direction, on) >= 0) on = ScalarMul(on, -1);
        getSphereUV(on, rec.u, rec.v);
        rec.t = temp; rec.p = r.PointAtParameter(temp); rec.normal = on; rec.prim = prim_id;
        mat = material;
        return true;
      }
      temp = (-b + std::sqrt(b * b - a * c)) / a;
      if (temp < tMax && temp > tMin) {
        Vec3 on = ScalarDiv(Sub(r.PointAtParameter(temp), center(r.time)), radius);
        if (Dot(r.direction, on) >= 0) on = ScalarMul(on, -1);
        getSphereUV(on, rec.u, rec.v);
        rec.t = temp; rec.p = r.PointAtParameter(temp);
        // second root: the record carries the UNFLIPPED outward normal (sphere.go:88-91)
        rec.normal = ScalarDiv(Sub(r.PointAtParameter(temp), center(r.time)), radius);
        rec.prim = prim_id;
        mat = material;
        return true;
      }
    }
    return false;
  }
  bool BoundingBox(AABB& box) const override {  // sphere.go:114-123
    Vec3 rr = V(radius, radius, radius);
    AABB b0{Sub(center0, rr), Add(center0, rr)}, b1{Sub(center1, rr), Add(center1, rr)};
    box = SurroundingBox(b0, b1);
    return true;
  }
  double PDFValue(const Vec3& o, const Vec3& v) const override {  // sphere.go:129-137
    HitRecord rec; const Material* m;
    if (Hit(NewRay(o, v, 0), 0.001, DBL_MAX, rec, m)) {
      double cosThetaMax = std::sqrt(1 - radius * radius / SquaredLength(Sub(center0, o)));
      double solidAngle = 2 * M_PI * (1 - cosThetaMax);
      return 1 / solidAngle;
    }
    return 0.0;
  }
  Vec3 Random(const Vec3& o, Rng& rng) const override {  // sphere.go:139-145
    Vec3 direction = Sub(center0, o);
    double distanceSquared = SquaredLength(direction);
    ONB uvw; uvw.BuildFromW(direction);
    return uvw.Local(RandomToSphere(radius, distanceSquared, rng));
  }
  bool IsEmitter() const override { return MaterialIsEmitter(material); }
};

// ---------------------------------------------------------------- axis-aligned rects
// axis: 0 = YZRect (plane x=k, a=y, b=z), 1 = XZRect (y=k, a=x, b=z), 2 = XYRect (z=k, a=x, b=y)
struct Rect : Hitable {
  int axis = 0;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;
  const Material* material = nullptr;
  static double comp(const Vec3& v, int i) { return i == 0 ? v.X : (i == 1 ? v.Y : v.Z); }
  int ia() const { return axis == 0 ? 1 : 0; }
  int ib() const { return axis == 2 ? 1 : 2; }
  // xyrect.go:38-53, xzrect.go:40-55, yzrect.go:38-53
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {
    if (g_stats) g_stats->others++;
    double t = (k - comp(r.origin, axis)) / comp(r.direction, axis);
    if (t < tMin || t > tMax) return false;
    double a = comp(r.origin, ia()) + (t * comp(r.direction, ia()));
    double b = comp(r.origin, ib()) + (t * comp(r.direction, ib()));
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    rec.u = (a - a0) / (a1 - a0);
    rec.v = (b - b0) / (b1 - b0);
    rec.t = t; rec.p = r.PointAtParameter(t);
    rec.normal = V(axis == 0 ? 1 : 0, axis == 1 ? 1 : 0, axis == 2 ? 1 : 0);
    rec.prim = prim_id;
    mat = material;
    return true;
  }
  bool BoundingBox(AABB& box) const override {  // xyrect.go:90-101 etc. (asymmetric -1e-4 / +1e-3)
    double lo[3], hi[3];
    lo[axis] = k - 0.0001; hi[axis] = k + 0.001;
    lo[ia()] = a0; hi[ia()] = a1; lo[ib()] = b0; hi[ib()] = b1;
    box.min = V(lo[0], lo[1], lo[2]); box.max = V(hi[0], hi[1], hi[2]);
    return true;
  }
  double PDFValue(const Vec3& o, const Vec3& v) const override {  // only XZRect: xzrect.go:106-116
    if (axis != 1) return 0.0;
    HitRecord rec; const Material* m;
    if (Hit(NewRay(o, v, 0), 0.001, DBL_MAX, rec, m)) {
      double area = (a1 - a0) * (b1 - b0);
      double distanceSquared = rec.t * rec.t * SquaredLength(v);
      double cosine = std::fabs(Dot(v, ScalarDiv(rec.normal, Length(v))));
      return distanceSquared / (cosine * area);
    }
    return 0;
  }
  Vec3 Random(const Vec3& o, Rng& rng) const override {  // xzrect.go:118-126; XY/YZ return (1,0,0)
    if (axis != 1) return V(1, 0, 0);
    double x = a0 + rng.Float64() * (a1 - a0);
    double z = b0 + rng.Float64() * (b1 - b0);
    return Sub(V(x, k, z), o);
  }
  bool IsEmitter() const override { return MaterialIsEmitter(material); }
};

// ---------------------------------------------------------------- wrappers
struct FlipNormals : Hitable {  // flip_normals.go
  std::unique_ptr<Hitable> h;
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {
    if (h->Hit(r, tMin, tMax, rec, mat)) { rec.normal = ScalarMul(rec.normal, -1); rec.prim = prim_id; return true; }
    return false;
  }
  bool BoundingBox(AABB& box) const override { return h->BoundingBox(box); }
  double PDFValue(const Vec3& o, const Vec3& v) const override { return h->PDFValue(o, v); }
  Vec3 Random(const Vec3& o, Rng& rng) const override { return h->Random(o, rng); }
  bool IsEmitter() const override { return h->IsEmitter(); }
};
struct Translate : Hitable {  // translate.go
  std::unique_ptr<Hitable> h;
  Vec3 offset;
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {
    Ray moved = NewRay(Sub(r.origin, offset), r.direction, r.time);
    if (h->Hit(moved, tMin, tMax, rec, mat)) { rec.p = Add(rec.p, offset); rec.prim = prim_id; return true; }
    return false;
  }
  bool BoundingBox(AABB& box) const override {
    AABB b;
    if (!h->BoundingBox(b)) return false;
    box.min = Add(b.min, offset); box.max = Add(b.max, offset);
    return true;
  }
  double PDFValue(const Vec3& o, const Vec3& v) const override { return h->PDFValue(o, v); }
  Vec3 Random(const Vec3& o, Rng& rng) const override { return h->Random(o, rng); }
  bool IsEmitter() const override { return h->IsEmitter(); }
};
struct RotateY : Hitable {  // rotate_y.go
  std::unique_ptr<Hitable> h;
  double sinTheta = 0, cosTheta = 1;
  AABB bbox; bool hasBox = false;
  void init(double angle) {  // rotate_y.go:27-79
    double radians = (M_PI / 180.0) * angle;
    sinTheta = std::sin(radians); cosTheta = std::cos(radians);
    AABB b; hasBox = h->BoundingBox(b);
    Vec3 mn = V(DBL_MAX, DBL_MAX, DBL_MAX), mx = V(-DBL_MAX, -DBL_MAX, -DBL_MAX);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
      double x = (double)i * b.max.X + (1.0 - (double)i) * b.min.X;
      double y = (double)j * b.max.Y + (1.0 - (double)j) * b.min.Y;
      double z = (double)k * b.max.Z + (1.0 - (double)k) * b.min.Z;
      double newx = cosTheta * x + sinTheta * z;
      double newz = -sinTheta * x + cosTheta * z;
      if (newx > mx.X) mx.X = newx; if (y > mx.Y) mx.Y = y; if (newz > mx.Z) mx.Z = newz;
      if (newx < mn.X) mn.X = newx; if (y < mn.Y) mn.Y = y; if (newz < mn.Z) mn.Z = newz;
    }
    bbox.min = mn; bbox.max = mx;
  }
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {  // :81-111
    Vec3 o = V(cosTheta * r.origin.X - sinTheta * r.origin.Z, r.origin.Y, sinTheta * r.origin.X + cosTheta * r.origin.Z);
    Vec3 d = V(cosTheta * r.direction.X - sinTheta * r.direction.Z, r.direction.Y,
               sinTheta * r.direction.X + cosTheta * r.direction.Z);
    Ray rot = NewRay(o, d, r.time);
    if (h->Hit(rot, tMin, tMax, rec, mat)) {
      Vec3 p = V(cosTheta * rec.p.X + sinTheta * rec.p.Z, rec.p.Y, -sinTheta * rec.p.X + cosTheta * rec.p.Z);
      Vec3 n = V(cosTheta * rec.normal.X + sinTheta * rec.normal.Z, rec.normal.Y,
                 -sinTheta * rec.normal.X + cosTheta * rec.normal.Z);
      rec.p = p; rec.normal = n; rec.prim = prim_id;
      return true;
    }
    return false;
  }
  bool BoundingBox(AABB& box) const override { box = bbox; return hasBox; }
  double PDFValue(const Vec3& o, const Vec3& v) const override { return h->PDFValue(o, v); }
  Vec3 Random(const Vec3& o, Rng& rng) const override { return h->Random(o, rng); }
  bool IsEmitter() const override { return h->IsEmitter(); }
};

// ---------------------------------------------------------------- HitableSlice (hitable_slice.go)
struct HitableSlice : Hitable {
  std::vector<Hitable*> hitables;  // not owned
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {  // :30-45
    bool hitAnything = false;
    double closestSoFar = tMax;
    HitRecord tmp; const Material* tm;
    for (Hitable* h : hitables) {
      if (h->Hit(r, tMin, closestSoFar, tmp, tm)) {
        rec = tmp; mat = tm; hitAnything = true; closestSoFar = rec.t;
      }
    }
    return hitAnything;
  }
  bool BoundingBox(AABB& box) const override {  // :70-96
    if (hitables.empty()) return false;
    if (!hitables[0]->BoundingBox(box)) return false;
    for (size_t i = 1; i < hitables.size(); i++) {
      AABB t;
      if (!hitables[i]->BoundingBox(t)) return false;
      box = SurroundingBox(box, t);
    }
    return true;
  }
  double PDFValue(const Vec3& o, const Vec3& v) const override {  // :98-105
    double weight = 1.0 / (double)hitables.size();
    double sum = 0;
    for (Hitable* h : hitables) sum += weight * h->PDFValue(o, v);
    return sum;
  }
  Vec3 Random(const Vec3& o, Rng& rng) const override {  // :107-110
    int index = (int)(rng.Float64() * (double)hitables.size());
    return hitables[index]->Random(o, rng);
  }
  bool IsEmitter() const override { return false; }
};

// Box = HitableSlice of six rects (box.go:23-46); PDFValue 0, Random (1,0,0), IsEmitter = material's.
struct Box : Hitable {
  std::vector<std::unique_ptr<Hitable>> owned;
  HitableSlice sides;
  const Material* mat_ = nullptr;
  Box(const Vec3& p0, const Vec3& p1, const Material* m) {
    mat_ = m;
    auto rect = [&](int axis, double a0, double a1, double b0, double b1, double k, bool flip) {
      auto r = std::make_unique<Rect>();
      r->axis = axis; r->a0 = a0; r->a1 = a1; r->b0 = b0; r->b1 = b1; r->k = k; r->material = m;
      if (flip) { auto f = std::make_unique<FlipNormals>(); f->h = std::move(r); owned.push_back(std::move(f)); }
      else owned.push_back(std::move(r));
      sides.hitables.push_back(owned.back().get());
    };
    rect(2, p0.X, p1.X, p0.Y, p1.Y, p1.Z, false);
    rect(2, p0.X, p1.X, p0.Y, p1.Y, p0.Z, true);
    rect(1, p0.X, p1.X, p0.Z, p1.Z, p1.Y, false);
    rect(1, p0.X, p1.X, p0.Z, p1.Z, p0.Y, true);
    rect(0, p0.Y, p1.Y, p0.Z, p1.Z, p1.X, false);
    rect(0, p0.Y, p1.Y, p0.Z, p1.Z, p0.X, true);
  }
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override {
    if (sides.Hit(r, tMin, tMax, rec, mat)) { rec.prim = prim_id; return true; }
    return false;
  }
  bool BoundingBox(AABB& box) const override { return sides.BoundingBox(box); }
  bool IsEmitter() const override { return MaterialIsEmitter(mat_); }
};

// ---------------------------------------------------------------- BVH4 (hitable/bvh4.go)
// box-test flavours
enum { BOX_SSE = 0, BOX_SCALAR = 1 };
uint8_t RayAABB4(int flavour, const float org[3], const float inv[3], const izpi_bvh4_node& n, float tMax);
float conservativeFloat32Min(double v);  // bvh4.go:494
float conservativeFloat32Max(double v);  // bvh4.go:506

struct BVH4 : Hitable {
  std::vector<izpi_bvh4_node> Nodes;
  std::vector<Hitable*> Primitives;     // reordered (leaf DFS order), not owned
  std::vector<int32_t> PrimitiveIndices; // Primitives[i] == original hitables[PrimitiveIndices[i]]
  int flavour = BOX_SSE;
  bool Hit(const Ray& r, double tMin, double tMax, HitRecord& rec, const Material*& mat) const override;
  bool BoundingBox(AABB& box) const override;
  bool IsEmitter() const override { return false; }
};
// newBVH4 (bvh4.go:558): randomFunc is the injected LCG / zero function.
std::unique_ptr<BVH4> newBVH4(const std::vector<Hitable*>& hitables, Rng* lcg, bool rand_zero);

}  // namespace orc
