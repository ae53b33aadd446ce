// intersect_g2.cuh -- cooperative closest hit with TWO LANES PER RAY, sixteen rays per warp.
//
// Same results, in the same order, as the four-lanes-per-ray traversal (intersect_g4.cuh) and therefore as (*BVH4).Hit
// (internal/hitable/bvh4.go:49-164).  The difference is how the work of one ray is laid over the lanes:
//   * lane j of a pair owns child slots 2j and 2j+1 of the node being visited (64 contiguous bytes of the child-major
//     node copy: four 128-bit loads) and runs the two fp32 slab tests back to back -- two independent dependency chains
//     per lane instead of one -- and two ballots assemble the pair's 4-bit hit mask;
//   * lane j tests primitives j and j+2 of a leaf, one after the other, both against the tMax at leaf entry; the pair then
//     replays the reference's sequential accept rule over the four candidates in array order;
//   * a warp therefore keeps 16 rays in flight: the fixed per-iteration cost of a phase (ballots, loop control, stack
//     bookkeeping) is shared by twice as many rays, and each lane carries twice the memory-level parallelism.
// The per-ray slab in shared memory is the one of intersect_g4.cuh with a compile-time stack depth: STACK entries, then
// the fp64 ray, the fp32 ray and one pad slot.
#pragma once
#include "intersect_g4.cuh"

namespace izpi {

// Two slab sizes are compiled: 48 entries (14 two-warp blocks per SM; every NewBVH4 tree up to ~10^9 primitives, 3 entries
// per level) and 52 entries (13 blocks, about 4 % slower on config 2) for somewhat deeper trees such as the device-built LBVH
// of the 11.5 M-triangle mesh (51 entries); beyond that the 4-lane kernel with the reference's full 64 entries takes over.
#ifndef IZPI_G2_STACK
#define IZPI_G2_STACK 48
#endif
constexpr int kG2Stack = IZPI_G2_STACK;
constexpr int kG2StackDeep = 52;      // stack entries of the 2-lane kernel's slab; deeper trees use the 4-lane kernel (64 entries)
constexpr int kG2Stragglers = 7;  // leave the node phase when <= this many PAIRS are still in it while leaves are pending

template <int STACK>
struct G2Slab {
  static constexpr int kSlots = STACK + 10;  // int2 slots per ray; (kSlots / 2) odd keeps the 16 slabs of a warp on different banks
};

template <int STACK>
__device__ __forceinline__ void g2_begin(G4State& s, const DScene& sc, const DRay& r, double tmax, int2* slab, int j) {
  const float ix = (float)(1.0 / r.d.x), iy = (float)(1.0 / r.d.y), iz = (float)(1.0 / r.d.z);  // bvh4.go:61-66
  const float ox = (float)r.o.x, oy = (float)r.o.y, oz = (float)r.o.z;                            // bvh4.go:67
  if (j == 0) {
    double* rs = reinterpret_cast<double*>(slab + STACK);
    rs[0] = r.o.x; rs[1] = r.o.y; rs[2] = r.o.z; rs[3] = r.d.x; rs[4] = r.d.y; rs[5] = r.d.z;
    float* rf = reinterpret_cast<float*>(slab + STACK + 6);
    rf[0] = ox; rf[1] = oy; rf[2] = oz; rf[3] = ix; rf[4] = iy; rf[5] = iz;
  }
  s.tmax = tmax; s.best = -1; s.sp = 0;
  s.cur = sc.n_nodes > 0 ? 0 : kIdle;
  const float big = 1e30f;  // see g4_begin: no NaN can arise in the slab test, FMNMX equals the SSE selects
  s.fast = fabsf(ox) < big && fabsf(oy) < big && fabsf(oz) < big && fabsf(ix) <= 3.0e38f && fabsf(iy) <= 3.0e38f &&
           fabsf(iz) <= 3.0e38f && ix != 0.0f && iy != 0.0f && iz != 0.0f;
}

template <bool COUNT>
__device__ __forceinline__ void g2_pop(G4State& s, const int2* stack, uint32_t& n_nodes) {
  while (s.sp > 0) {
    s.sp--;
    int2 e = stack[s.sp];
    if (e.x >= 0) { s.cur = e.x; return; }
    if (COUNT) n_nodes++;  // the reference loads the leaf node before its box test can fail
    if ((float)s.tmax >= __int_as_float(e.y)) { s.cur = g4_leaf_ref(~e.x); return; }
  }
  s.cur = kIdle;
}

// one lane of RayAABB4_SIMD (bvh4_simd_amd64.go:52-101) on the child record {a, b}; returns hit, tmn = slab entry distance
__device__ __forceinline__ bool g2_slab(const float4 a, const float4 b, float ox, float oy, float oz, float ix, float iy, float iz,
                                        float tmaxf, bool fast, float& tmn) {
  const float t0x = __fmul_rn(__fsub_rn(a.x, ox), ix), t1x = __fmul_rn(__fsub_rn(a.w, ox), ix);
  const float t0y = __fmul_rn(__fsub_rn(a.y, oy), iy), t1y = __fmul_rn(__fsub_rn(b.x, oy), iy);
  const float t0z = __fmul_rn(__fsub_rn(a.z, oz), iz), t1z = __fmul_rn(__fsub_rn(b.y, oz), iz);
  float tmx;
  if (fast) {
    tmn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    tmx = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  } else {
    tmn = sse_min(t0x, t1x); tmx = sse_max(t0x, t1x);
    tmn = sse_max(tmn, sse_min(t0y, t1y)); tmx = sse_min(tmx, sse_max(t0y, t1y));
    tmn = sse_max(tmn, sse_min(t0z, t1z)); tmx = sse_min(tmx, sse_max(t0z, t1z));
  }
  return (tmx >= tmn) && (tmx >= 0.0f) && (tmaxf >= tmn);  // empty slots carry NaN bounds (context.cu): always false
}

// Node phase for the whole warp.  pshift = lane & ~1 (first lane of the pair), j = lane & 1.
template <bool COUNT, int STACK>
__device__ __forceinline__ void g2_node_phase(G4State& s, const DScene& sc, int2* stack, int pshift, int j, uint32_t& n_nodes, int stragglers) {
  const unsigned full = 0xffffffffu;
  __syncwarp();  // the slab written by lane 0 of the pair in g2_begin
  const float4 ro = *reinterpret_cast<const float4*>(stack + STACK + 6);
  const float2 ri = *reinterpret_cast<const float2*>(stack + STACK + 8);
  const float ox = ro.x, oy = ro.y, oz = ro.z, ix = ro.w, iy = ri.x, iz = ri.y;
  for (;;) {
    const bool in_node = s.cur >= 0;
    const unsigned nm = __ballot_sync(full, in_node);
    if (nm == 0) break;
    if (__popc(nm) <= 2 * stragglers && __any_sync(full, g4_in_leaf(s.cur))) break;
    bool hit0 = false, hit1 = false;
    float tmn0 = 0.0f, tmn1 = 0.0f;
    int ref0 = 0, ref1 = 0;
    if (in_node) {
      const float4* np = sc.nodes_t + (size_t)s.cur * 8 + 4 * j;
      float4 a0, b0, a1, b1;
      g4_load_child(np, a0, b0);
      g4_load_child(np + 2, a1, b1);
      ref0 = __float_as_int(b0.z); ref1 = __float_as_int(b1.z);
      if (COUNT) n_nodes++;
      const float tmaxf = (float)s.tmax;  // float32(tMax) at node entry (bvh4.go:100)
      hit0 = g2_slab(a0, b0, ox, oy, oz, ix, iy, iz, tmaxf, s.fast, tmn0);
      hit1 = g2_slab(a1, b1, ox, oy, oz, ix, iy, iz, tmaxf, s.fast, tmn1);
    }
    // child k of the pair's node: k = 2 * lane_in_pair + slot; bits of x0 / x1: lane 0's and lane 1's first / second slot
    const unsigned x0 = (__ballot_sync(full, hit0) >> pshift) & 3u, x1 = (__ballot_sync(full, hit1) >> pshift) & 3u;
    const unsigned m = (x0 & 1u) | ((x1 & 1u) << 1) | ((x0 & 2u) << 1) | ((x1 & 2u) << 2);
    const int first = m ? __ffs(m) - 1 : 0;
    const int nref = __shfl_sync(full, (first & 1) ? ref1 : ref0, pshift + (first >> 1));
    if (in_node) {
      if (m == 0) {
        g2_pop<COUNT>(s, stack, n_nodes);
      } else {
        const int k0 = 2 * j, k1 = 2 * j + 1;  // later hit children are pushed in slot order (bvh4.go:141-145)
        if (hit0 && k0 != first) stack[s.sp + __popc(m & ((1u << k0) - 1u)) - 1] = make_int2(ref0, __float_as_int(tmn0));
        if (hit1 && k1 != first) stack[s.sp + __popc(m & ((1u << k1) - 1u)) - 1] = make_int2(ref1, __float_as_int(tmn1));
        s.sp += __popc(m) - 1;
        if (nref >= 0) {
          s.cur = nref;  // first hit child is visited next (bvh4.go:137-140)
        } else {         // ... and when it is a leaf its box test repeats with the same tMax: it passes
          if (COUNT && !sc.root_is_leaf) n_nodes++;
          s.cur = g4_leaf_ref(~nref);
        }
      }
    }
    __syncwarp();  // pushes visible to the pair before any pop
  }
}

// Leaf phase: lane j tests primitives j and j + 2; the pair replays `if hit { tMax = rec.T() }` (bvh4.go:125-134) in array order.
template <bool COUNT, bool F32, int STACK>
__device__ __forceinline__ void g2_leaf_phase(G4State& s, const DScene& sc, const int2* stack, int pshift, int j, uint32_t& n_nodes,
                                              uint32_t& n_prims, double tmin) {
  const unsigned full = 0xffffffffu;
  const bool in_leaf = g4_in_leaf(s.cur);
  if (!__any_sync(full, in_leaf)) return;
  const int leaf = g4_leaf_of(s.cur);  // meaningful when in_leaf
  const int start = leaf >> 2, cnt = (leaf & 3) + 1;
  bool ok[2] = {false, false}, strict[2] = {false, false};
  double t[2] = {0.0, 0.0};
#pragma unroll  // both tests in one instruction stream: the second record's loads overlap the first test (668 -> 680 Mrays/s)
  for (int h = 0; h < 2; h++) {
    const int k = j + 2 * h;
    bool okh = false, sth = false;
    double th = 0.0;
    if (in_leaf && k < cnt) {
      PrimRec pr = load_rec(sc.prims + start + k);
      const double* rs = reinterpret_cast<const double*>(stack + STACK);
      if (F32 && tag_type(pr.tag) == IZPI_PRIM_TRIANGLE && tag_xform(pr.tag) == 0) {
        float tf = 0.0f;
        okh = tri_test_f32(pr, rs, (float)tmin, (float)s.tmax, tf);
        th = (double)tf;
      } else {
        DRay r;
        r.o = mk(rs[0], rs[1], rs[2]); r.d = mk(rs[3], rs[4], rs[5]); r.time = 0; r.lambda = 0;
        DHit hr;
        okh = prim_hit<false>(sc, start + k, pr, r, tmin, s.tmax, hr);
        th = hr.t;
      }
      sth = tag_type(pr.tag) == IZPI_PRIM_SPHERE;  // Sphere.Hit compares strictly (sphere.go:73,84)
      if (COUNT) n_prims++;
    }
    if (h == 0) { ok[0] = okh; strict[0] = sth; t[0] = th; } else { ok[1] = okh; strict[1] = sth; t[1] = th; }
  }
  const unsigned okw0 = __ballot_sync(full, ok[0]), okw1 = __ballot_sync(full, ok[1]);
  if (okw0 | okw1) {  // a hit anywhere in the warp is rare (about one per ray): resolve only then
    const unsigned stw0 = __ballot_sync(full, strict[0]), stw1 = __ballot_sync(full, strict[1]);
#pragma unroll
    for (int k = 0; k < 4; k++) {  // candidate k lives in lane (k & 1) of the pair, test slot (k >> 1)
      const double tk = __shfl_sync(full, (k >> 1) ? t[1] : t[0], pshift + (k & 1));
      const unsigned okb = (((k >> 1) ? okw1 : okw0) >> (pshift + (k & 1))) & 1u;
      const unsigned stb = (((k >> 1) ? stw1 : stw0) >> (pshift + (k & 1))) & 1u;
      if (okb) {
        const bool acc = stb ? (tk < s.tmax) : (tk <= s.tmax);
        if (acc) { s.tmax = tk; s.best = start + k; }
      }
    }
  }
  if (in_leaf) g2_pop<COUNT>(s, stack, n_nodes);
}

}  // namespace izpi
