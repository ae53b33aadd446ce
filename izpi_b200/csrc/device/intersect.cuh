// intersect.cuh -- closest-hit on the device: fp32 4-wide slab test, fp64 primitive tests and the
// BVH4 traversal in the reference's exact order.
//
//   box4()            <- RayAABB4_SIMD, SSE flavour   internal/hitable/bvh4_simd_amd64.go:27-110
//   tri_test()        <- (*Triangle).Hit              internal/hitable/triangle.go:193-226
//   sphere_test()     <- (*Sphere).Hit                internal/hitable/sphere.go:63-95
//   rect_test()       <- X?Rect.Hit                   internal/hitable/xyrect.go:38, xzrect.go:40, yzrect.go:38
//   prim_hit<FULL>()  <- FlipNormals/Translate/RotateY/Box wrappers (flip_normals.go:27, translate.go:29,
//                        rotate_y.go:81, box.go:48) around the above
//   world_closest()   <- HitableSlice.Hit (hitable_slice.go:30-45) over (*BVH4).Hit (bvh4.go:49-164)
//
// Compiled with --fmad=false: every product and sum rounds separately, like Go on amd64.
#pragma once
#include "dscene.cuh"
#include "shade_textures.cuh"

namespace izpi {

struct DRay {
  d3 o, d;
  double time, lambda;
};
struct DHit {  // hitrecord.HitRecord (hitrecord.go:6-12)
  double t, u, v;
  d3 p, n;
};

__device__ __forceinline__ d3 point_at(const DRay& r, double t) { return r.o + r.d * t; }  // ray.go:45

// MINPS / MAXPS of `x.Min(y)`, `x.Max(y)`: the second operand wins on NaN and on equality.
__device__ __forceinline__ float sse_min(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float sse_max(float x, float y) { return x > y ? x : y; }

// One lane of RayAABB4_SIMD.  Returns hit; tnear is the slab entry distance (t_min of the test).
__device__ __forceinline__ bool box1(float ox, float oy, float oz, float ix, float iy, float iz, float mnx, float mny,
                                     float mnz, float mxx, float mxy, float mxz, float tmaxf) {
  float t0x = __fmul_rn(__fsub_rn(mnx, ox), ix), t1x = __fmul_rn(__fsub_rn(mxx, ox), ix);
  float tmn = sse_min(t0x, t1x), tmx = sse_max(t0x, t1x);
  float t0y = __fmul_rn(__fsub_rn(mny, oy), iy), t1y = __fmul_rn(__fsub_rn(mxy, oy), iy);
  tmn = sse_max(tmn, sse_min(t0y, t1y));
  tmx = sse_min(tmx, sse_max(t0y, t1y));
  float t0z = __fmul_rn(__fsub_rn(mnz, oz), iz), t1z = __fmul_rn(__fsub_rn(mxz, oz), iz);
  tmn = sse_max(tmn, sse_min(t0z, t1z));
  tmx = sse_min(tmx, sse_max(t0z, t1z));
  return (tmx >= tmn) && (tmx >= 0.0f) && (tmaxf >= tmn);
}

// ---- primitives --------------------------------------------------------------------------
__device__ __forceinline__ bool tri_test(d3 v0, d3 e1, d3 e2, const DRay& r, double tmin, double tmax, double& t,
                                         double& u, double& v) {
  const double epsilon = 1e-8;
  d3 h = cross(r.d, e2);
  double a = dot(e1, h);
  if (fabs(a) < epsilon) return false;
  double f = 1.0 / a;
  d3 s = r.o - v0;
  u = f * dot(s, h);
  if (u < -epsilon || u > 1.0 + epsilon) return false;
  d3 q = cross(s, e1);
  v = f * dot(r.d, q);
  if (v < -epsilon || u + v > 1.0 + epsilon) return false;
  t = f * dot(e2, q);
  if (t < tmin || t > tmax) return false;  // inclusive at both ends
  return true;
}

// root: 1 = first root accepted, 2 = second root accepted
__device__ __forceinline__ int sphere_test(d3 c, double radius, const DRay& r, double tmin, double tmax, double& t) {
  d3 oc = r.o - c;
  double a = dot(r.d, r.d);
  double b = dot(oc, r.d);
  double cc = dot(oc, oc) - (radius * radius);
  double disc = (b * b) - (a * cc);
  if (disc > 0) {
    double temp = (-b - sqrt(b * b - a * cc)) / a;
    if (temp < tmax && temp > tmin) { t = temp; return 1; }  // strict
    temp = (-b + sqrt(b * b - a * cc)) / a;
    if (temp < tmax && temp > tmin) { t = temp; return 2; }
  }
  return 0;
}

// axis: 0 = YZRect (x = k), 1 = XZRect (y = k), 2 = XYRect (z = k)
__device__ __forceinline__ bool rect_test(int axis, double a0, double a1, double b0, double b1, double k, const DRay& r,
                                          double tmin, double tmax, double& t, double& u, double& v) {
  int ia = axis == 0 ? 1 : 0, ib = axis == 2 ? 1 : 2;
  t = (k - comp(r.o, axis)) / comp(r.d, axis);
  if (t < tmin || t > tmax) return false;
  double a = comp(r.o, ia) + (t * comp(r.d, ia));
  double b = comp(r.o, ib) + (t * comp(r.d, ib));
  if (a < a0 || a > a1 || b < b0 || b > b1) return false;
  u = (a - a0) / (a1 - a0);
  v = (b - b0) / (b1 - b0);
  return true;
}

__device__ __forceinline__ d3 axis_normal(int axis) { return mk(axis == 0 ? 1.0 : 0.0, axis == 1 ? 1.0 : 0.0, axis == 2 ? 1.0 : 0.0); }

__device__ __forceinline__ void sphere_uv(d3 p, double& u, double& v) {  // sphere.go:29-35
  double phi = atan2(p.z, p.x);
  double theta = asin(p.y);
  u = 1.0 - (phi + M_PI) / (2.0 * M_PI);
  v = (theta + M_PI / 2.0) / M_PI;
}

struct PrimRec {
  double a[9];
  int32_t orig_id;
  uint32_t tag;
};
__device__ __forceinline__ PrimRec load_rec(const izpi_prim_rec* p) {  // 5 x 128-bit loads
  const double2* q = reinterpret_cast<const double2*>(p);
  PrimRec r;
  double2 x0 = __ldg(q), x1 = __ldg(q + 1), x2 = __ldg(q + 2), x3 = __ldg(q + 3);
  int4 x4 = __ldg(reinterpret_cast<const int4*>(q + 4));
  r.a[0] = x0.x; r.a[1] = x0.y; r.a[2] = x1.x; r.a[3] = x1.y; r.a[4] = x2.x; r.a[5] = x2.y; r.a[6] = x3.x; r.a[7] = x3.y;
  r.a[8] = __hiloint2double(x4.y, x4.x);
  r.orig_id = x4.z; r.tag = (uint32_t)x4.w;
  return r;
}
__device__ __forceinline__ int tag_type(uint32_t tag) { return tag & 7u; }
__device__ __forceinline__ bool tag_flip(uint32_t tag) { return (tag >> 3) & 1u; }
__device__ __forceinline__ int tag_material(uint32_t tag) { return (tag >> 4) & 0x3fffu; }
__device__ __forceinline__ int tag_xform(uint32_t tag) { return (int)(tag >> 18); }

// Primitive.Hit for the record `idx`.  FULL also fills the HitRecord (u, v, p, normal) the way
// the reference's wrappers compose it.
// need_uv = false skips the sphere's (u, v) = atan2 / asin (sphere.go:29-35) when the caller never reads them; u = v = 0 then.
template <bool FULL>
__device__ __forceinline__ bool prim_hit(const DScene& sc, int idx, const PrimRec& pr, const DRay& ray, double tmin,
                                         double tmax, DHit& h, bool need_uv = true) {
  DRay r = ray;
  const izpi_xform* xf = nullptr;
  int xi = tag_xform(pr.tag);
  if (xi) {
    xf = sc.xforms + (xi - 1);
    if (xf->has_translate) r.o = r.o - mk(xf->offset[0], xf->offset[1], xf->offset[2]);  // translate.go:30
    if (xf->has_rotate) {                                                                // rotate_y.go:82-91
      double c = xf->cos_theta, s = xf->sin_theta;
      r.o = mk(c * r.o.x - s * r.o.z, r.o.y, s * r.o.x + c * r.o.z);
      r.d = mk(c * r.d.x - s * r.d.z, r.d.y, s * r.d.x + c * r.d.z);
    }
  }
  bool ok = false;
  double t = 0, u = 0, v = 0;
  d3 n = mk(0, 0, 0);
  switch (tag_type(pr.tag)) {
    case IZPI_PRIM_TRIANGLE: {
      ok = tri_test(mk(pr.a[0], pr.a[1], pr.a[2]), mk(pr.a[3], pr.a[4], pr.a[5]), mk(pr.a[6], pr.a[7], pr.a[8]), r, tmin,
                    tmax, t, u, v);
      if (FULL && ok) {  // triangle.go:228-264
        const double epsilon = 1e-8;
        const izpi_tri_attr& at = sc.attrs[idx];
        double w = 1.0 - u - v;
        double sum = u + v + w;
        if (fabs(sum - 1.0) > epsilon) { u /= sum; v /= sum; w /= sum; }
        double uu = w * at.uv[0] + u * at.uv[2] + v * at.uv[4];
        double vv = w * at.uv[1] + u * at.uv[3] + v * at.uv[5];
        n = mk(at.normal[0], at.normal[1], at.normal[2]);
        const izpi_material_spec& m = sc.materials[tag_material(pr.tag)];
        if (m.type == IZPI_MAT_PBR && m.normal_tex >= 0) {
          d3 nt = texture_value(sc, m.normal_tex, uu, vv);
          nt = mk(2 * nt.x - 1.0, 2 * nt.y - 1.0, 2 * nt.z - 1.0);
          d3 tg = mk(at.tangent[0], at.tangent[1], at.tangent[2]), bt = mk(at.bitangent[0], at.bitangent[1], at.bitangent[2]);
          n = unit(mk(tg.x * nt.x + bt.x * nt.y + n.x * nt.z, tg.y * nt.x + bt.y * nt.y + n.y * nt.z,
                      tg.z * nt.x + bt.z * nt.y + n.z * nt.z));
        }
        u = uu; v = vv;
      }
      break;
    }
    case IZPI_PRIM_SPHERE: {
      d3 c = mk(pr.a[0], pr.a[1], pr.a[2]);
      int root = sphere_test(c, pr.a[3], r, tmin, tmax, t);
      ok = root != 0;
      if (FULL && ok) {
        d3 on = (point_at(r, t) - c) / pr.a[3];
        d3 outward = on;
        if (dot(r.d, on) >= 0) on = on * -1.0;
        if (need_uv) sphere_uv(on, u, v);
        n = root == 1 ? on : outward;  // second root keeps the unflipped normal (sphere.go:88-91)
      }
      break;
    }
    case IZPI_PRIM_XYRECT: case IZPI_PRIM_XZRECT: case IZPI_PRIM_YZRECT: {
      int axis = tag_type(pr.tag) == IZPI_PRIM_YZRECT ? 0 : (tag_type(pr.tag) == IZPI_PRIM_XZRECT ? 1 : 2);
      ok = rect_test(axis, pr.a[0], pr.a[1], pr.a[2], pr.a[3], pr.a[4], r, tmin, tmax, t, u, v);
      if (FULL) n = axis_normal(axis);
      break;
    }
    case IZPI_PRIM_BOX: {  // six sides, HitableSlice.Hit with shrinking closestSoFar (box.go:27-34)
      double closest = tmax;
      const double *p0 = pr.a, *p1 = pr.a + 3;
#pragma unroll
      for (int sgn = 0; sgn < 6; sgn++) {
        int axis = sgn < 2 ? 2 : (sgn < 4 ? 1 : 0);
        bool flip = sgn & 1;
        int ia = axis == 0 ? 1 : 0, ib = axis == 2 ? 1 : 2;
        double k = flip ? p0[axis] : p1[axis];
        double tt, uu, vv;
        if (rect_test(axis, p0[ia], p1[ia], p0[ib], p1[ib], k, r, tmin, closest, tt, uu, vv)) {
          ok = true; closest = tt; t = tt; u = uu; v = vv;
          if (FULL) { n = axis_normal(axis); if (flip) n = n * -1.0; }
        }
      }
      break;
    }
  }
  if (!ok) return false;
  h.t = t;
  if (FULL) {
    d3 p = point_at(r, t);
    if (xf) {
      if (xf->has_rotate) {  // rotate_y.go:96-105
        double c = xf->cos_theta, s = xf->sin_theta;
        p = mk(c * p.x + s * p.z, p.y, -s * p.x + c * p.z);
        n = mk(c * n.x + s * n.z, n.y, -s * n.x + c * n.z);
      }
      if (xf->has_translate) p = p + mk(xf->offset[0], xf->offset[1], xf->offset[2]);  // translate.go:32
    }
    if (tag_flip(pr.tag)) n = n * -1.0;  // flip_normals.go:29
    h.u = u; h.v = v; h.p = p; h.n = n;
  }
  return true;
}

// ---- traversal ----------------------------------------------------------------------------
// stack: int32 slots `stack[depth * stride]`, >= 64 deep (bvh4.go:71-72).
// Returns the record index of the closest primitive (or -1) and its t in best_t.
template <bool COUNT>
__device__ __forceinline__ int world_closest(const DScene& sc, const DRay& r, double tmin, double tmax, double& best_t,
                                             int32_t* stack, int stride, uint32_t& n_nodes, uint32_t& n_prims) {
  int best = -1;
  DHit h;
  if (sc.world_kind == IZPI_WORLD_SLICE) {  // hitable_slice.go:30-45
    for (int i = 0; i < sc.n_prims; i++) {
      PrimRec pr = load_rec(sc.prims + i);
      if (COUNT) n_prims++;
      if (prim_hit<false>(sc, i, pr, r, tmin, tmax, h)) { tmax = h.t; best = i; }
    }
    best_t = tmax;
    return best;
  }
  if (sc.n_nodes == 0) return -1;
  // ray -> float32 (bvh4.go:61-67): fp64 divide, then round
  const float ix = (float)(1.0 / r.d.x), iy = (float)(1.0 / r.d.y), iz = (float)(1.0 / r.d.z);
  const float ox = (float)r.o.x, oy = (float)r.o.y, oz = (float)r.o.z;
  int sp = 0;
  int cur = 0;
  while (cur != -1) {
    const float4* np = sc.nodes + (size_t)cur * 8;
    const float4 mnx = __ldg(np), mny = __ldg(np + 1), mnz = __ldg(np + 2);
    const float4 mxx = __ldg(np + 3), mxy = __ldg(np + 4), mxz = __ldg(np + 5);
    const int4 child = __ldg(reinterpret_cast<const int4*>(np + 6));
    const int4 count = __ldg(reinterpret_cast<const int4*>(np + 7));
    if (COUNT) n_nodes++;
    const float tmaxf = (float)tmax;  // float32(tMax) at node entry (bvh4.go:100); DBL_MAX -> +Inf
    const bool m0 = box1(ox, oy, oz, ix, iy, iz, mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, tmaxf);
    const bool m1 = box1(ox, oy, oz, ix, iy, iz, mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, tmaxf);
    const bool m2 = box1(ox, oy, oz, ix, iy, iz, mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, tmaxf);
    const bool m3 = box1(ox, oy, oz, ix, iy, iz, mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, tmaxf);
    int next = -1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool m = i == 0 ? m0 : (i == 1 ? m1 : (i == 2 ? m2 : m3));
      const int ci = i == 0 ? child.x : (i == 1 ? child.y : (i == 2 ? child.z : child.w));
      const int pc = i == 0 ? count.x : (i == 1 ? count.y : (i == 2 ? count.z : count.w));
      if (!m || ci == -1) continue;
      if (pc > 0) {
        for (int p = 0; p < pc; p++) {
          PrimRec pr = load_rec(sc.prims + ci + p);
          if (COUNT) n_prims++;
          if (prim_hit<false>(sc, ci + p, pr, r, tmin, tmax, h)) { tmax = h.t; best = ci + p; }
        }
      } else if (next == -1) {
        next = ci;
      } else {
        stack[sp * stride] = ci;
        sp++;
      }
    }
    if (next != -1) cur = next;
    else if (sp > 0) { sp--; cur = stack[sp * stride]; }
    else cur = -1;
  }
  best_t = tmax;
  return best;
}

}  // namespace izpi
