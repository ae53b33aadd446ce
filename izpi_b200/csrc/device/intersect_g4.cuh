// intersect_g4.cuh -- cooperative closest hit: FOUR LANES PER RAY, eight rays per warp.
//
// Same results as (*BVH4).Hit (internal/hitable/bvh4.go:49-164) in the same order, organised for
// the SIMT machine instead of one scalar loop per ray:
//   * lane j of a ray group owns child slot j of the node being visited: one 32-byte child record
//     (two 128-bit loads), one fp32 slab test, then a 4-bit group ballot decides next / pushes;
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
//   * node phase and leaf phase alternate warp-wide (while-while traversal), so the 8 groups of a warp
//     execute box tests together and primitive tests together.
#pragma once
#include "intersect.cuh"

namespace izpi {

constexpr int kG4Stack = 64;  // entries per ray (bvh4.go:71)

struct G4State {
  DRay r;
  float ox, oy, oz, ix, iy, iz;
  double tmin, tmax;
  int best;        // record index of the closest primitive so far
  int sp;          // stack pointer
  int cur;         // >= 0: inner node to visit; kLeaf: leaf pending; kIdle: no ray
  int leaf_start, leaf_cnt;
};
constexpr int kLeaf = -2, kIdle = -1;

// stack entry: ref >= 0 inner node; ref < 0 leaf: ~ref = (first primitive << 2) | (count - 1)
__device__ __forceinline__ int leaf_ref(int start, int cnt) { return ~((start << 2) | (cnt - 1)); }

__device__ __forceinline__ void g4_begin(G4State& s, const DScene& sc, const DRay& r, double tmin, double tmax) {
  s.r = r;
  s.ix = (float)(1.0 / r.d.x); s.iy = (float)(1.0 / r.d.y); s.iz = (float)(1.0 / r.d.z);  // bvh4.go:61-66
  s.ox = (float)r.o.x; s.oy = (float)r.o.y; s.oz = (float)r.o.z;                            // bvh4.go:67
  s.tmin = tmin; s.tmax = tmax; s.best = -1; s.sp = 0;
  s.cur = sc.n_nodes > 0 ? 0 : kIdle;
  s.leaf_start = 0; s.leaf_cnt = 0;
}

// Pop until something to do is found.  Returns false when the stack is empty (ray finished).
template <bool COUNT>
__device__ __forceinline__ bool g4_pop(G4State& s, const int2* stack, uint32_t& n_nodes) {
  while (s.sp > 0) {
    s.sp--;
    int2 e = stack[s.sp];
    if (e.x >= 0) { s.cur = e.x; return true; }
    if (COUNT) n_nodes++;  // the reference loads the leaf node before its box test can fail
    if ((float)s.tmax >= __int_as_float(e.y)) {
      int v = ~e.x;
      s.cur = kLeaf; s.leaf_start = v >> 2; s.leaf_cnt = (v & 3) + 1;
      return true;
    }
  }
  return false;
}

// One inner-node visit by the 4 lanes of a group.  gmask = this group's lanes, j = lane within group.
template <bool COUNT>
__device__ __forceinline__ void g4_node(G4State& s, const DScene& sc, int2* stack, unsigned gmask, int gshift, int j,
                                        uint32_t& n_nodes) {
  const float4* np = sc.nodes_t + (size_t)s.cur * 8 + 2 * j;
  const float4 a = __ldg(np);
  const float4 b = __ldg(np + 1);
  const int idx = __float_as_int(b.z), cnt = __float_as_int(b.w);
  if (COUNT) n_nodes++;
  const float tmaxf = (float)s.tmax;  // float32(tMax) at node entry (bvh4.go:100)
  // one lane of RayAABB4_SIMD (bvh4_simd_amd64.go:52-101); tnear is its t_min
  float t0 = __fmul_rn(__fsub_rn(a.x, s.ox), s.ix), t1 = __fmul_rn(__fsub_rn(a.w, s.ox), s.ix);
  float tmn = sse_min(t0, t1), tmx = sse_max(t0, t1);
  t0 = __fmul_rn(__fsub_rn(a.y, s.oy), s.iy); t1 = __fmul_rn(__fsub_rn(b.x, s.oy), s.iy);
  tmn = sse_max(tmn, sse_min(t0, t1)); tmx = sse_min(tmx, sse_max(t0, t1));
  t0 = __fmul_rn(__fsub_rn(a.z, s.oz), s.iz); t1 = __fmul_rn(__fsub_rn(b.y, s.oz), s.iz);
  tmn = sse_max(tmn, sse_min(t0, t1)); tmx = sse_min(tmx, sse_max(t0, t1));
  const bool hit = (tmx >= tmn) && (tmx >= 0.0f) && (tmaxf >= tmn) && (idx != -1);
  const unsigned m = (__ballot_sync(gmask, hit) >> gshift) & 0xfu;
  if (m == 0) {
    if (!g4_pop<COUNT>(s, stack, n_nodes)) s.cur = kIdle;
    return;
  }
  const int first = __ffs(m) - 1;
  const int ref = cnt > 0 ? leaf_ref(idx, cnt) : idx;
  if (hit && j != first) {  // later hit children are pushed in slot order (bvh4.go:141-145)
    int rank = __popc(m & ((1u << j) - 1u)) - 1;
    stack[s.sp + rank] = make_int2(ref, __float_as_int(tmn));
  }
  s.sp += __popc(m) - 1;
  const int nref = __shfl_sync(gmask, ref, first, 4);
  if (nref >= 0) {
    s.cur = nref;  // first hit child is visited next (bvh4.go:137-140)
  } else {         // ... and when it is a leaf its box test repeats with the same tMax: it passes
    if (COUNT && !sc.root_is_leaf) n_nodes++;  // (a root that is itself the leaf-node was already counted)
    int v = ~nref;
    s.cur = kLeaf; s.leaf_start = v >> 2; s.leaf_cnt = (v & 3) + 1;
  }
  __syncwarp(gmask);  // pushes visible to the group before any pop
}

// Leaf visit: lane j tests primitive j; the group then replays the reference's sequential
// `if hit { tMax = rec.T() }` loop (bvh4.go:125-134) over the four candidates.
template <bool COUNT>
__device__ __forceinline__ void g4_leaf(G4State& s, const DScene& sc, const int2* stack, unsigned gmask, int gshift, int j,
                                        uint32_t& n_nodes, uint32_t& n_prims) {
  bool ok = false, strict = false;
  double t = 0;
  if (j < s.leaf_cnt) {
    PrimRec pr = load_rec(sc.prims + s.leaf_start + j);
    DHit h;
    ok = prim_hit<false>(sc, s.leaf_start + j, pr, s.r, s.tmin, s.tmax, h);
    t = h.t;
    strict = tag_type(pr.tag) == IZPI_PRIM_SPHERE;  // Sphere.Hit compares strictly (sphere.go:73,84)
  }
  if (COUNT) n_prims += (uint32_t)(j < s.leaf_cnt);
  const unsigned okm = (__ballot_sync(gmask, ok) >> gshift) & 0xfu;
  const unsigned stm = (__ballot_sync(gmask, strict) >> gshift) & 0xfu;
  if (okm) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double tk = __shfl_sync(gmask, t, k, 4);
      if ((okm >> k) & 1u) {
        bool acc = ((stm >> k) & 1u) ? (tk < s.tmax) : (tk <= s.tmax);
        if (acc) { s.tmax = tk; s.best = s.leaf_start + k; }
      }
    }
  }
  if (!g4_pop<COUNT>(s, stack, n_nodes)) s.cur = kIdle;
}

}  // namespace izpi
