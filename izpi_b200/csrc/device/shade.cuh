// shade.cuh -- light transport on the device: counter RNG, spectral tables, materials, PDFs and
// light sampling.  Every function restates one reference function with the same operation order:
//
//   Rng                       replaces fastrandom.LCG (fastrandom.go:41-47) with a counter generator
//                             keyed (seed, pixel, sample); same 2^-32 granularity of Float64()
//   sample_wavelength()       <- spectral.SampleWavelength (spectral/spectral.go:184-224)
//   cie_values()              <- spectral.GetCIEValues (spectral.go:227-253)
//   random_cosine_direction() <- vec3.RandomCosineDirection (vec3/vec3.go:119-127, factor 2 kept)
//   Onb                       <- onb.BuildFromW / Local (onb/onb.go:38-52,63)
//   reflect/refract/schlick   <- material/material.go:20-43
//   lights_random()/lights_pdf_value() <- HitableSlice.Random/PDFValue (hitable_slice.go:98-110) over
//                             XZRect (xzrect.go:106-126), Sphere (sphere.go:129-145), Triangle
//                             (triangle.go:271-280,317-326); every other hitable: 0 and (1,0,0)
//   scatter_*()               <- Lambertian/Metal/Dielectric/PBR Scatter + SpectralScatter
#pragma once
#include "intersect.cuh"

namespace izpi {

// ---- RNG ------------------------------------------------------------------------------------
struct Rng {
  uint64_t key;
  uint32_t ctr;
};
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t pixel, uint64_t sample) {
  return mix64(mix64(seed ^ 0x9E3779B97F4A7C15ull) + pixel * 0xD1B54A32D192ED03ull + sample * 0x8CB92BA72F3D8DD7ull);
}
__device__ __forceinline__ double rnd(Rng& r) {  // Float64(): uniform in [0,1) on a 2^-32 lattice
  uint64_t z = mix64(r.key + (uint64_t)(++r.ctr) * 0x9E3779B97F4A7C15ull);
  return (double)(uint32_t)(z >> 32) / 4294967296.0;
}

// ---- CIE tables (spectral.go:16-59), 380..750 nm @ 5 nm ---------------------------------------
__constant__ double c_cie[3][75];
constexpr double kCieYIntegral = 21.3768;  // spectral.go:64
__device__ __forceinline__ double cie_wavelength(int i) { return 380.0 + 5.0 * (double)i; }

__device__ __forceinline__ void sample_wavelength(double random, double& lambda, double& pdf) {
  double target = random * kCieYIntegral;
  double current = 0.0;
  for (int i = 0; i < 75; i++) {
    double y = c_cie[1][i];
    if (current + y >= target) {
      if (i > 0) {
        double t = (target - current) / y;
        lambda = cie_wavelength(i - 1) + t * (cie_wavelength(i) - cie_wavelength(i - 1));
        double iy = c_cie[1][i - 1] + t * (c_cie[1][i] - c_cie[1][i - 1]);
        pdf = iy / kCieYIntegral;
        return;
      }
      lambda = cie_wavelength(i);
      pdf = y / kCieYIntegral;
      return;
    }
    current += y;
  }
  lambda = 750.0;
  pdf = c_cie[1][74] / kCieYIntegral;
}

__device__ __forceinline__ d3 cie_values(double wl) {
  if (wl <= 380.0) return mk(c_cie[0][0], c_cie[1][0], c_cie[2][0]);
  if (wl >= 750.0) return mk(c_cie[0][74], c_cie[1][74], c_cie[2][74]);
  int index = cie_bucket_at_or_above(wl);  // first i with cieWavelengths[i] >= wl (spectral.go:235-241); 380 < wl < 750 here, so index >= 1
  double w1 = cie_wavelength(index - 1), w2 = cie_wavelength(index);
  double t = (wl - w1) / (w2 - w1);
  return mk(c_cie[0][index - 1] + t * (c_cie[0][index] - c_cie[0][index - 1]),
            c_cie[1][index - 1] + t * (c_cie[1][index] - c_cie[1][index - 1]),
            c_cie[2][index - 1] + t * (c_cie[2][index] - c_cie[2][index - 1]));
}

// SpectralPowerDistribution.Value (spectral.go:151-181): inclusive clamps, unlike interpolateSPD
__device__ __forceinline__ double spd_value(const double* w, const double* val, int n, double wl) {
  if (n == 0) return 0.0;
  if (wl <= w[0]) return val[0];
  if (wl >= w[n - 1]) return val[n - 1];
  for (int i = 0; i < n - 1; i++) {
    double w1 = w[i], w2 = w[i + 1];
    if (wl >= w1 && wl <= w2) {
      double t = (wl - w1) / (w2 - w1);
      return val[i] + t * (val[i + 1] - val[i]);
    }
  }
  return 0.0;
}

// ---- sampling helpers ------------------------------------------------------------------------
struct Onb {
  d3 u, v, w;
};
__device__ __forceinline__ Onb onb_from_w(d3 n) {
  Onb o;
  o.w = unit(n);
  d3 a = fabs(o.w.x) > 0.9 ? mk(0, 1, 0) : mk(1, 0, 0);
  o.v = unit(cross(o.w, a));
  o.u = cross(o.w, o.v);
  return o;
}
__device__ __forceinline__ d3 onb_local(const Onb& o, d3 a) { return (o.u * a.x + o.v * a.y) + o.w * a.z; }

__device__ __forceinline__ d3 random_cosine_direction(Rng& rng) {
  double r1 = rnd(rng);
  double r2 = rnd(rng);
  double z = sqrt(1 - r2);
  double phi = 2 * M_PI * r1;
  double x = cos(phi) * 2 * sqrt(r2);
  double y = sin(phi) * 2 * sqrt(r2);
  return mk(x, y, z);
}
__device__ __forceinline__ d3 random_to_sphere(double radius, double dist2, Rng& rng) {  // vec3.go:130-138
  double r1 = rnd(rng);
  double r2 = rnd(rng);
  double z = 1 + r2 * (sqrt(1 - radius * radius / dist2) - 1);
  double phi = 2 * M_PI * r1;
  double x = cos(phi) * sqrt(1 - z * z);
  double y = sin(phi) * sqrt(1 - z * z);
  return mk(x, y, z);
}
__device__ __forceinline__ d3 random_in_unit_sphere(Rng& rng) {  // material.go:10-18
  for (;;) {
    double x = rnd(rng);
    double y = rnd(rng);
    double z = rnd(rng);
    d3 p = mk(x, y, z) * 2.0 - mk(1.0, 1.0, 1.0);
    if (sqlen(p) < 1.0) return p;
  }
}
__device__ __forceinline__ d3 reflect(d3 v, d3 n) { return v - n * (2 * dot(v, n)); }
__device__ __forceinline__ bool refract(d3 v, d3 n, double ni_over_nt, d3& out) {
  d3 uv = unit(v);
  double dt = dot(uv, n);
  double disc = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt);
  if (disc > 0) {
    out = (uv - n * dt) * ni_over_nt - n * sqrt(disc);
    return true;
  }
  return false;
}
__device__ __forceinline__ double schlick(double cosine, double ref_idx) {
  double r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
  r0 = r0 * r0;
  return r0 + (1.0 - r0) * pow((1.0 - cosine), 5.0);
}
__device__ __forceinline__ double cosine_pdf_value(d3 w, d3 direction) {  // pdf/cosine.go:27-34
  double c = dot(unit(direction), w);
  return c > 0 ? c / M_PI : 0.0;
}

// ---- lights ----------------------------------------------------------------------------------
// Wrappers delegate Random/PDFValue to the wrapped hitable with the SAME arguments
// (flip_normals.go:45-51, translate.go:58-64, rotate_y.go:147-153), so both run on the bare primitive.
__device__ __forceinline__ PrimRec bare(PrimRec pr) {
  pr.tag &= 0x3fff7u;  // keep type + material, drop flip and xform
  return pr;
}

__device__ __forceinline__ d3 light_random(const DScene& sc, int rec, d3 o, Rng& rng) {
  PrimRec pr = load_rec(sc.prims + rec);
  switch (tag_type(pr.tag)) {
    case IZPI_PRIM_XZRECT: {
      double x = pr.a[0] + rnd(rng) * (pr.a[1] - pr.a[0]);
      double z = pr.a[2] + rnd(rng) * (pr.a[3] - pr.a[2]);
      return mk(x, pr.a[4], z) - o;
    }
    case IZPI_PRIM_SPHERE: {
      d3 direction = mk(pr.a[0], pr.a[1], pr.a[2]) - o;
      double dist2 = sqlen(direction);
      Onb uvw = onb_from_w(direction);
      return onb_local(uvw, random_to_sphere(pr.a[3], dist2, rng));
    }
    case IZPI_PRIM_TRIANGLE: {
      // Triangle.Random lerps between the ORIGINAL vertices (v0 + e1 would differ from vertex1 by a
      // rounding), so the attribute record keeps vertex1 / vertex2.
      d3 v0 = mk(pr.a[0], pr.a[1], pr.a[2]);
      const izpi_tri_attr& at = sc.attrs[rec];
      d3 p1 = mk(at.vertex1[0], at.vertex1[1], at.vertex1[2]), p2 = mk(at.vertex2[0], at.vertex2[1], at.vertex2[2]);
      double t1 = rnd(rng);
      d3 p01 = lerp(v0, p1, t1);
      double t2 = rnd(rng);
      d3 p02 = lerp(v0, p2, t2);
      double t3 = rnd(rng);
      return lerp(p01, p02, t3) - o;
    }
    default:
      return mk(1, 0, 0);
  }
}

__device__ __forceinline__ double light_pdf_value(const DScene& sc, int rec, d3 o, d3 v) {
  PrimRec pr = bare(load_rec(sc.prims + rec));
  int type = tag_type(pr.tag);
  if (type != IZPI_PRIM_XZRECT && type != IZPI_PRIM_SPHERE && type != IZPI_PRIM_TRIANGLE) return 0.0;
  DRay r;
  r.o = o; r.d = v; r.time = 0; r.lambda = 0;
  DHit h;
  if (!prim_hit<true>(sc, rec, pr, r, 0.001, DBL_MAX, h, false)) return 0.0;  // PDFValue reads t and the normal, never u, v
  if (type == IZPI_PRIM_SPHERE) {
    double cos_theta_max = sqrt(1 - pr.a[3] * pr.a[3] / sqlen(mk(pr.a[0], pr.a[1], pr.a[2]) - o));
    double solid_angle = 2 * M_PI * (1 - cos_theta_max);
    return 1 / solid_angle;
  }
  double area = type == IZPI_PRIM_XZRECT ? (pr.a[1] - pr.a[0]) * (pr.a[3] - pr.a[2]) : sc.attrs[rec].area;
  double dist2 = h.t * h.t * sqlen(v);
  double cosine = fabs(dot(v, h.n / len(v)));
  return dist2 / (cosine * area);
}

__device__ __forceinline__ d3 lights_random(const DScene& sc, d3 o, Rng& rng) {
  int index = (int)(rnd(rng) * (double)sc.n_lights);
  return light_random(sc, sc.lights[index], o, rng);
}
// HitableSlice.PDFValue (hitable_slice.go:98-105): sum over the lights of weight * PDFValue, in list order.
// A direction hits few of the lights, and the expensive part of a light's PDFValue (square roots, divisions, the hit record)
// runs only for those -- under SIMT that tail would execute once per light with the two or three lanes that hit it.  So the
// loop is split: a cheap, lane-uniform pass marks each lane's CANDIDATE lights (spheres: the discriminant test of
// sphere.go:68-72, exactly the reference's; rects and triangles: always), then every lane walks its own candidates, so
// different lanes evaluate different lights in the same iteration.  A light that is not a candidate contributes
// weight * 0 = +0, which leaves the running sum unchanged, and each lane still adds its terms in list order: same value.
__device__ __forceinline__ double lights_pdf_value(const DScene& sc, d3 o, d3 v) {
  double weight = 1.0 / (double)sc.n_lights;
  double sum = 0.0;
  const double vv = dot(v, v);
  for (int base = 0; base < sc.n_lights; base += 32) {
    const int m = sc.n_lights - base < 32 ? sc.n_lights - base : 32;
    unsigned cand = 0;
    for (int k = 0; k < m; k++) {
      const izpi_prim_rec* p = sc.prims + sc.lights[base + k];
      const int type = tag_type(__ldg(&p->tag));
      bool c = type == IZPI_PRIM_XZRECT || type == IZPI_PRIM_TRIANGLE;
      if (type == IZPI_PRIM_SPHERE) {  // Sphere.Hit's `discriminant > 0` (sphere.go:68-72), same operation order as sphere_test
        const double2 c01 = __ldg(reinterpret_cast<const double2*>(p->a));
        const double2 c2r = __ldg(reinterpret_cast<const double2*>(p->a) + 1);
        d3 oc = o - mk(c01.x, c01.y, c2r.x);
        double b = dot(oc, v);
        double cc = dot(oc, oc) - (c2r.y * c2r.y);
        c = (b * b) - (vv * cc) > 0;
      }
      cand |= (c ? 1u : 0u) << k;
    }
    while (cand) {
      const int k = __ffs(cand) - 1;
      cand &= cand - 1;
      sum += weight * light_pdf_value(sc, sc.lights[base + k], o, v);
    }
  }
  return sum;
}

}  // namespace izpi
