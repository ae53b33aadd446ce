// intersect_g4.cuh -- cooperative closest hit: FOUR LANES PER RAY, eight rays per warp.
//
// Same results as (*BVH4).Hit (internal/hitable/bvh4.go:49-164) in the same order, organised for
// the SIMT machine instead of one scalar loop per ray:
//   * lane j of a ray group owns child slot j of the node being visited: one 32-byte child record
//     (two 128-bit loads), one fp32 slab test, then a warp ballot (4 bits per group) decides
//     next / pushes;
//   * the reference's leaf-nodes (own node, slot 0 only, bvh4.go:737-760) are folded into their
//     parent's slot at upload time: the slot carries (first primitive, count) and the entry pushed on
//     the stack keeps the slab-entry distance, so "visiting the leaf node" becomes the single
//     comparison float32(tMax_now) >= tNear -- the only one of the three mask conditions
//     (bvh4_simd_amd64.go:88-101) that can change between the parent's test and the visit, because
//     the leaf-node's box is bit-identical to the parent's slot box (bvh4.go:750-755 vs :782-787);
//   * the <= 4 primitives of a leaf are tested by the 4 lanes in parallel against the tMax at leaf
//     entry and then resolved in array order with the shrinking-tMax rule (a primitive accepted
//     sequentially is exactly one accepted against the entry tMax whose t also passes the running
//     tMax);
//   * node phase and leaf phase alternate warp-wide.  Both are warp-uniform loops with predicated
//     bodies, so every collective uses the full mask (sub-warp masks make the compiler serialise one
//     group at a time).  The node phase ends as soon as at most kNodeStragglers groups still stand on
//     an inner node while leaves are pending (profiles/sim_schedule.py: waiting for the last group
//     costs 1.4x more warp instructions).
#pragma once
#include "intersect.cuh"

namespace izpi {

constexpr int kG4Stack = 64;        // entries per ray (bvh4.go:71)
constexpr int kG4Slab = kG4Stack + 10;  // int2 slots of shared memory per ray: the stack, the fp64 ray (6 doubles), the fp32 ray (6 floats), 1 pad
constexpr int kNodeStragglers = 4;  // default: leave the node phase when <= this many groups are still in it

// Per-lane registers hold only what the node phase needs; the fp64 origin/direction (used by the primitive
// tests alone) live in the ray's shared-memory slab behind its stack, which buys resident warps: the kernel is
// latency-bound and its throughput is linear in them (scripts/sweep_blocks.sh).
struct G4State {
  double tmax;
  int best;   // record index of the closest primitive so far
  int sp;     // stack pointer
  int cur;    // >= 0: inner node to visit; kIdle: no ray; <= kLeaf: a pending leaf, see g4_leaf_ref / g4_leaf_of
  bool fast;  // no NaN can arise in the slab test: FMNMX min/max equal the SSE selects
};
constexpr int kLeaf = -2, kIdle = -1;
// a pending leaf L = (first primitive << 2) | (count - 1) >= 0 is kept in `cur` as kLeaf - L (one register for both)
__device__ __forceinline__ int g4_leaf_ref(int leaf) { return kLeaf - leaf; }
__device__ __forceinline__ int g4_leaf_of(int cur) { return kLeaf - cur; }
__device__ __forceinline__ bool g4_in_leaf(int cur) { return cur <= kLeaf; }

// stack entry: ref >= 0 inner node; ref < 0 leaf: ~ref = (first primitive << 2) | (count - 1)
__device__ __forceinline__ void g4_begin(G4State& s, const DScene& sc, const DRay& r, double tmax, int2* slab, int j) {
  const float ix = (float)(1.0 / r.d.x), iy = (float)(1.0 / r.d.y), iz = (float)(1.0 / r.d.z);  // bvh4.go:61-66
  const float ox = (float)r.o.x, oy = (float)r.o.y, oz = (float)r.o.z;                            // bvh4.go:67
  if (j == 0) {
    double* rs = reinterpret_cast<double*>(slab + kG4Stack);
    rs[0] = r.o.x; rs[1] = r.o.y; rs[2] = r.o.z; rs[3] = r.d.x; rs[4] = r.d.y; rs[5] = r.d.z;
    // the fp32 ray of the slab test is node-phase state only: it lives behind the fp64 ray and is reloaded at phase entry,
    // so that it holds no registers during the (register-hungry) primitive tests
    float* rf = reinterpret_cast<float*>(slab + kG4Stack + 6);
    rf[0] = ox; rf[1] = oy; rf[2] = oz; rf[3] = ix; rf[4] = iy; rf[5] = iz;
  }
  s.tmax = tmax; s.best = -1; s.sp = 0;
  s.cur = sc.n_nodes > 0 ? 0 : kIdle;
  // (bound - o) * inv is NaN only for 0 * Inf or Inf * 0: impossible when the origin is small enough for the
  // subtraction not to overflow and 1/d is finite and non-zero.  MINPS/MAXPS then agree with FMNMX (up to the
  // sign of zero, which no comparison below sees).
  const float big = 1e30f;
  s.fast = fabsf(ox) < big && fabsf(oy) < big && fabsf(oz) < big && fabsf(ix) <= 3.0e38f && fabsf(iy) <= 3.0e38f &&
           fabsf(iz) <= 3.0e38f && ix != 0.0f && iy != 0.0f && iz != 0.0f;
}

// One 32-byte child record of the child-major node copy in a single 256-bit load (sm_100: LDG.E.256): half the load
// instructions and half the L1 tag look-ups of two 128-bit loads.  p must be 32-byte aligned (it is: 128-byte nodes).
__device__ __forceinline__ void g4_load_child(const float4* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}

// Pop until something to do is found; leaves the ray idle when the stack is empty.
template <bool COUNT>
__device__ __forceinline__ void g4_pop(G4State& s, const DScene& sc, const int2* stack, int j, uint32_t& n_nodes) {
  while (s.sp > 0) {
    s.sp--;
    int2 e = stack[s.sp];
    if (e.x >= 0) { s.cur = e.x; return; }
    if (COUNT) n_nodes++;  // the reference loads the leaf node before its box test can fail
    if ((float)s.tmax >= __int_as_float(e.y)) { s.cur = g4_leaf_ref(~e.x); return; }
  }
  s.cur = kIdle;
}

// Node phase for the whole warp.  gshift = 4 * group, j = lane within group.
template <bool COUNT>
__device__ __forceinline__ void g4_node_phase(G4State& s, const DScene& sc, int2* stack, unsigned lane, int gshift, int j,
                                              uint32_t& n_nodes, int stragglers = kNodeStragglers) {
  const unsigned full = 0xffffffffu;
  __syncwarp();  // the slab written by lane 0 of the group in g4_begin
  const float4 ro = *reinterpret_cast<const float4*>(stack + kG4Stack + 6);
  const float2 ri = *reinterpret_cast<const float2*>(stack + kG4Stack + 8);
  const float ox = ro.x, oy = ro.y, oz = ro.z, ix = ro.w, iy = ri.x, iz = ri.y;
  for (;;) {
    const bool in_node = s.cur >= 0;
    const unsigned nm = __ballot_sync(full, in_node);
    if (nm == 0) break;
    if (__popc(nm) <= 4 * stragglers && __any_sync(full, g4_in_leaf(s.cur))) break;
    bool hit = false;
    float tmn = 0.0f;
    int ref = 0;
    if (in_node) {
      const float4* np = sc.nodes_t + (size_t)s.cur * 8 + 2 * j;
      float4 a, b;
      g4_load_child(np, a, b);
      ref = __float_as_int(b.z);  // inner child: node index; leaf child: ~((first primitive << 2) | (count - 1)), built at upload
      if (COUNT) n_nodes++;
      const float tmaxf = (float)s.tmax;  // float32(tMax) at node entry (bvh4.go:100)
      // one lane of RayAABB4_SIMD (bvh4_simd_amd64.go:52-101); tmn is its t_min, tmx its t_max
      const float t0x = __fmul_rn(__fsub_rn(a.x, ox), ix), t1x = __fmul_rn(__fsub_rn(a.w, ox), ix);
      const float t0y = __fmul_rn(__fsub_rn(a.y, oy), iy), t1y = __fmul_rn(__fsub_rn(b.x, oy), iy);
      const float t0z = __fmul_rn(__fsub_rn(a.z, oz), iz), t1z = __fmul_rn(__fsub_rn(b.y, oz), iz);
      float tmx;
      if (s.fast) {
        tmn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
        tmx = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
      } else {
        tmn = sse_min(t0x, t1x); tmx = sse_max(t0x, t1x);
        tmn = sse_max(tmn, sse_min(t0y, t1y)); tmx = sse_min(tmx, sse_max(t0y, t1y));
        tmn = sse_max(tmn, sse_min(t0z, t1z)); tmx = sse_min(tmx, sse_max(t0z, t1z));
      }
      // an empty slot (ChildIndex == -1, bvh4.go:108-110) carries NaN bounds in this copy: every comparison is false
      hit = (tmx >= tmn) && (tmx >= 0.0f) && (tmaxf >= tmn);
    }
    const unsigned m = (__ballot_sync(full, hit) >> gshift) & 0xfu;
    const int first = m ? __ffs(m) - 1 : 0;
    const int nref = __shfl_sync(full, ref, (lane & ~3u) + first);
    if (in_node) {
      if (m == 0) {
        g4_pop<COUNT>(s, sc, stack, j, n_nodes);
      } else {
        if (hit && j != first)  // later hit children are pushed in slot order (bvh4.go:141-145)
          stack[s.sp + __popc(m & ((1u << j) - 1u)) - 1] = make_int2(ref, __float_as_int(tmn));
        s.sp += __popc(m) - 1;
        if (nref >= 0) {
          s.cur = nref;  // first hit child is visited next (bvh4.go:137-140)
        } else {         // ... and when it is a leaf its box test repeats with the same tMax: it passes
          if (COUNT && !sc.root_is_leaf) n_nodes++;  // (a root that is itself the leaf-node was already counted)
          s.cur = g4_leaf_ref(~nref);
        }
      }
    }
    __syncwarp();  // pushes visible to the group before any pop
  }
}

// Leaf phase: lane j tests primitive j; the group then replays the reference's sequential
// `if hit { tMax = rec.T() }` loop (bvh4.go:125-134) over the four candidates.
// fp32 Moeller-Trumbore for the optional IZPI_TRACE_FP32 mode (reported separately, not bit-exact by design)
__device__ __forceinline__ bool tri_test_f32(const PrimRec& pr, const double* rs, float tmin, float tmax, float& t) {
  const float eps = 1e-8f;
  float dx = (float)rs[3], dy = (float)rs[4], dz = (float)rs[5];
  float e1x = (float)pr.a[3], e1y = (float)pr.a[4], e1z = (float)pr.a[5], e2x = (float)pr.a[6], e2y = (float)pr.a[7], e2z = (float)pr.a[8];
  float hx = dy * e2z - dz * e2y, hy = -(dx * e2z - dz * e2x), hz = dx * e2y - dy * e2x;
  float a = e1x * hx + e1y * hy + e1z * hz;
  if (fabsf(a) < eps) return false;
  float f = 1.0f / a;
  // the origin difference is formed in fp64 first: both terms can be large and close
  float sx = (float)(rs[0] - pr.a[0]), sy = (float)(rs[1] - pr.a[1]), sz = (float)(rs[2] - pr.a[2]);
  float u = f * (sx * hx + sy * hy + sz * hz);
  if (u < -eps || u > 1.0f + eps) return false;
  float qx = sy * e1z - sz * e1y, qy = -(sx * e1z - sz * e1x), qz = sx * e1y - sy * e1x;
  float v = f * (dx * qx + dy * qy + dz * qz);
  if (v < -eps || u + v > 1.0f + eps) return false;
  t = f * (e2x * qx + e2y * qy + e2z * qz);
  return !(t < tmin || t > tmax);
}

template <bool COUNT, bool F32 = false>
__device__ __forceinline__ void g4_leaf_phase(G4State& s, const DScene& sc, const int2* stack, unsigned lane, int gshift, int j,
                                              uint32_t& n_nodes, uint32_t& n_prims, double tmin) {
  const unsigned full = 0xffffffffu;
  const bool in_leaf = g4_in_leaf(s.cur);
  if (!__any_sync(full, in_leaf)) return;
  const int leaf = g4_leaf_of(s.cur);  // meaningful when in_leaf
  const int start = leaf >> 2, cnt = (leaf & 3) + 1;
  bool ok = false, strict = false;
  double t = 0;
  if (in_leaf && j < cnt) {
    PrimRec pr = load_rec(sc.prims + start + j);
    const double* rs = reinterpret_cast<const double*>(stack + kG4Stack);
    if (F32 && tag_type(pr.tag) == IZPI_PRIM_TRIANGLE && tag_xform(pr.tag) == 0) {
      float tf = 0.0f;
      ok = tri_test_f32(pr, rs, (float)tmin, (float)s.tmax, tf);
      t = (double)tf;
    } else {
      DRay r;
      r.o = mk(rs[0], rs[1], rs[2]); r.d = mk(rs[3], rs[4], rs[5]); r.time = 0; r.lambda = 0;
      DHit h;
      ok = prim_hit<false>(sc, start + j, pr, r, tmin, s.tmax, h);
      t = h.t;
    }
    strict = tag_type(pr.tag) == IZPI_PRIM_SPHERE;  // Sphere.Hit compares strictly (sphere.go:73,84)
    if (COUNT) n_prims++;
  }
  const unsigned okw = __ballot_sync(full, ok);
  if (okw) {  // a hit anywhere in the warp is rare (about one per ray): resolve only then
    const unsigned stw = __ballot_sync(full, strict);
    const unsigned okm = (okw >> gshift) & 0xfu, stm = (stw >> gshift) & 0xfu;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double tk = __shfl_sync(full, t, (lane & ~3u) + k);
      if ((okm >> k) & 1u) {
        bool acc = ((stm >> k) & 1u) ? (tk < s.tmax) : (tk <= s.tmax);
        if (acc) { s.tmax = tk; s.best = start + k; }
      }
    }
  }
  if (in_leaf) g4_pop<COUNT>(s, sc, stack, j, n_nodes);
}

}  // namespace izpi
