// vecmath.h -- fp64 3-vectors with the reference's operation order (internal/vec3/vec3.go).
// Shared by the host flattening code and the device kernels.  Everything is compiled with
// multiply-add fusion OFF (nvcc --fmad=false, g++ -ffp-contract=off): Go on amd64 never fuses,
// and bit-exact closest-hit parity depends on it.
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define IZ_HD __host__ __device__ __forceinline__
#else
#define IZ_HD inline
#endif

namespace izpi {

struct d3 {
  double x, y, z;
};

IZ_HD d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
IZ_HD d3 operator+(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }   // vec3.Add
IZ_HD d3 operator-(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }   // vec3.Sub
IZ_HD d3 operator*(d3 a, double t) { return mk(a.x * t, a.y * t, a.z * t); }     // vec3.ScalarMul
IZ_HD d3 operator/(d3 a, double t) { return mk(a.x / t, a.y / t, a.z / t); }     // vec3.ScalarDiv
IZ_HD d3 hadamard(d3 a, d3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }    // vec3.Mul
IZ_HD double dot(d3 a, d3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); } // vec3.Dot (left to right)
IZ_HD d3 cross(d3 a, d3 b) {                                                     // vec3.Cross (note the negated middle term)
  return mk((a.y * b.z) - (a.z * b.y), -((a.x * b.z) - (a.z * b.x)), (a.x * b.y) - (a.y * b.x));
}
IZ_HD double sqlen(d3 v) { return (v.x * v.x) + (v.y * v.y) + (v.z * v.z); }
IZ_HD double len(d3 v) { return sqrt(sqlen(v)); }
IZ_HD d3 unit(d3 v) { return v / len(v); }  // vec3.UnitVector == MakeUnitVector
IZ_HD d3 lerp(d3 a, d3 b, double t) {       // vec3.Lerp
  return mk((1 - t) * a.x + t * b.x, (1 - t) * a.y + t * b.y, (1 - t) * a.z + t * b.z);
}
IZ_HD double comp(d3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

}  // namespace izpi
