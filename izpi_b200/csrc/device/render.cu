// render.cu -- wavefront path tracer behind the tile seam (renderRectRGB render/rgb.go:12,
// renderRectSpectral render/spectral.go:14, worker.RenderTile worker/render.go:17).
//
// One path = one (pixel, sample).  A batch of paths lives in HBM as 160-byte PathState records;
// each bounce runs
//   extend_kernel   closest hit for every live path (persistent warps, the trace kernel's traversal),
//                   misses terminate, hits are filed BY MATERIAL (dscene.cuh: one bin per image-textured material, one
//                   shared bin per class for the rest) with warp match/ballot/popc prefix ranks and one
//                   atomicAdd per (warp, bin): a counting sort by material ID into per-material bins;
//   shade_kernel<C> one launch per class over its bins, material after material: emission, BSDF sampling, light/BSDF mixture PDF
//                   (sampler/colour.go:33-65, sampler/spectral.go:47-80 in iterative form), survivors
//                   are compacted into the next bounce's queue the same way;
// then resolve_kernel sums each pixel's samples in sample order (the reference's loop order,
// rgb.go:30-39) into the fp64 canvas.  All arithmetic is fp64 without FMA.
#include <algorithm>
#include <cstring>
#include <deque>

#include <cub/device/device_radix_sort.cuh>

#include "../../../include/izpi_host.h"
#include "intersect_g2.cuh"
#include "shade.cuh"

using namespace izpi;

namespace {

constexpr int kThreads = 128;
constexpr int kStackDepth = 64;
constexpr int kClasses = 5;  // IZPI_MAT_* types
constexpr int kBinCounters = 16;  // first bin counter; kBinCounters + kMaxBins counters per queue set

struct alignas(16) PathState {
  double ox, oy, oz, dx, dy, dz, time, lambda;  // ray.RayImpl
  double bx, by, bz;                            // throughput (spectral: bx)
  double ax, ay, az;                            // accumulated radiance / final sample value
  double hit_t, lpdf;                           // closest hit t; wavelength pdf (spectral)
  uint64_t key;                                 // RNG stream key
  uint32_t ctr;                                 // RNG draw counter
  int32_t depth;
  int32_t hit_rec;
  int32_t alive;
  int32_t pad0, pad1;
};
static_assert(sizeof(PathState) == 160, "PathState layout");

struct RenderParams {
  int32_t width, height, spp, max_depth, sampler;
  int32_t n_bg;
  double background[3];
  const double* bg_w;
  const double* bg_v;
  uint64_t seed;
};

struct Queues {
  int32_t* cur;                 // live paths entering this bounce
  int32_t* next;                // survivors
  uint32_t* next_keys;          // coherence keys of the survivors, same positions as `next` (nullptr: the scene is not sorted)
  int32_t* bins;                // [n_bins][capacity]
  unsigned long long* counters; // [0] cur count, [1] next count, [7] work head, [8] rays traced, [9] nodes visited, [10] primitive tests (counting kernels), [kBinCounters + b] count of bin b
  int32_t capacity;
};

__device__ __forceinline__ DRay path_ray(const PathState& p) {
  DRay r;
  r.o = mk(p.ox, p.oy, p.oz); r.d = mk(p.dx, p.dy, p.dz); r.time = p.time; r.lambda = p.lambda;
  return r;
}

// Warp-aggregated append: lanes with `want` and the same `cls` share one atomicAdd.
// Returns the position written (within the bin), -1 for lanes that did not push.
__device__ __forceinline__ long long push_binned(bool want, int cls, int32_t value, int32_t* base, int stride,
                                                 unsigned long long* counts) {
  unsigned active = __ballot_sync(0xffffffffu, want);
  if (!want) return -1;
  unsigned peers = __match_any_sync(active, cls);
  int leader = __ffs(peers) - 1;
  unsigned lane = threadIdx.x & 31u;
  unsigned long long slot = 0;
  if ((int)lane == leader) slot = atomicAdd(&counts[cls], (unsigned long long)__popc(peers));
  slot = __shfl_sync(peers, slot, leader);
  int rank = __popc(peers & ((1u << lane) - 1u));
  base[(size_t)cls * stride + slot + rank] = value;
  return (long long)(slot + rank);
}

// Coherence key of a ray (see the sort below): low bits of the primitive the ray leaves, cell of the origin over the root box,
// cell of the direction (2 bits per axis of d / |d|_inf, sign included).
#ifndef IZPI_SORT_KEY
#define IZPI_SORT_KEY 0  // the three layouts measure the same on config 4 (127.6 / 128.8 / 129.0 Msamples/s); the 16-bit one sorts in two passes
#endif
#if IZPI_SORT_KEY == 0    // 1 bit of the primitive | 3 bits of origin per axis | 2 bits of direction per axis
constexpr int kSortKeyBits = 16, kKeyPrimBits = 1, kKeyPosBits = 3;
#elif IZPI_SORT_KEY == 1  // 4 | 2 | 2
constexpr int kSortKeyBits = 16, kKeyPrimBits = 4, kKeyPosBits = 2;
#else                     // 5 | 3 | 2: three radix passes
constexpr int kSortKeyBits = 20, kKeyPrimBits = 5, kKeyPosBits = 3;
#endif
__device__ __forceinline__ uint32_t coherence_key(const DScene& sc, int32_t prim, d3 o, d3 d) {
  uint32_t key = ((uint32_t)prim & ((1u << kKeyPrimBits) - 1u)) << (3 * kKeyPosBits + 6);
  const double ov[3] = {o.x, o.y, o.z}, dv[3] = {d.x, d.y, d.z};
  double dm = fmax(fabs(dv[0]), fmax(fabs(dv[1]), fabs(dv[2])));
  if (!(dm > 0)) dm = 1.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float f = ((float)ov[a] - sc.world_min[a]) / (sc.world_max[a] - sc.world_min[a]);
    int c = (int)(f * (float)(1 << kKeyPosBits));
    c = c < 0 ? 0 : (c > (1 << kKeyPosBits) - 1 ? (1 << kKeyPosBits) - 1 : c);
    int q = (int)((dv[a] / dm + 1.0) * 2.0);
    q = q < 0 ? 0 : (q > 3 ? 3 : q);
    key |= (uint32_t)c << (6 + kKeyPosBits * a);
    key |= (uint32_t)q << (2 * a);
  }
  return key;
}

// ---- ray generation -------------------------------------------------------------------------
// path index = pixel_local * s_count + s.  Draw order per sample (rgb.go:31-36 / render/spectral.go:77-88):
// [lambda] u v, camera disc (rejection), camera time.
__global__ void raygen_kernel(const __grid_constant__ DScene sc, RenderParams rp, PathState* paths,
                              const uint32_t* __restrict__ pixels, int n_pixels, int s_begin, int s_count, Queues q) {
  long long n = (long long)n_pixels * s_count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int pl = (int)(i / s_count), s = s_begin + (int)(i % s_count);
    uint32_t xy = pixels[pl];
    int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
    PathState p;
    Rng rng;
    rng.key = stream_key(rp.seed, (uint64_t)y * (uint64_t)rp.width + (uint64_t)x, (uint64_t)s);
    rng.ctr = 0;
    double lambda = 0, lpdf = 1;
    bool live = true;
    if (rp.sampler == IZPI_SAMPLER_SPECTRAL) {
      sample_wavelength(rnd(rng), lambda, lpdf);
      if (lpdf == 0) live = false;  // `continue` in RenderPixelSpectral (spectral.go:80-82)
    }
    const izpi_camera& c = sc.camera;
    if (live) {
      double u = ((double)x + rnd(rng)) / (double)rp.width;
      double v = ((double)y + rnd(rng)) / (double)rp.height;
      d3 disc;
      for (;;) {  // camera.randomInUnitDisc (camera.go:82-89)
        double a = rnd(rng);
        double b = rnd(rng);
        disc = mk(a, b, 0) * 2.0 - mk(1.0, 1.0, 0);
        if (dot(disc, disc) < 1.0) break;
      }
      d3 rd = disc * c.lens_radius;
      d3 cu = mk(c.u[0], c.u[1], c.u[2]), cv = mk(c.v[0], c.v[1], c.v[2]);
      d3 offset = cu * rd.x + cv * rd.y;
      double time = c.time0 + rnd(rng) * (c.time1 - c.time0);
      d3 org = mk(c.origin[0], c.origin[1], c.origin[2]);
      d3 llc = mk(c.lower_left_corner[0], c.lower_left_corner[1], c.lower_left_corner[2]);
      d3 hor = mk(c.horizontal[0], c.horizontal[1], c.horizontal[2]), ver = mk(c.vertical[0], c.vertical[1], c.vertical[2]);
      d3 dir = (((llc + hor * u) + ver * v) - org) - offset;  // camera.go:66-69
      d3 o = org + offset;
      p.ox = o.x; p.oy = o.y; p.oz = o.z; p.dx = dir.x; p.dy = dir.y; p.dz = dir.z; p.time = time;
    } else {
      p.ox = p.oy = p.oz = p.dx = p.dy = p.dz = p.time = 0;
    }
    p.lambda = lambda; p.lpdf = lpdf;
    p.bx = p.by = p.bz = 1.0;
    p.ax = p.ay = p.az = 0.0;
    p.hit_t = 0; p.key = rng.key; p.ctr = rng.ctr; p.depth = 0; p.hit_rec = -1; p.alive = live ? 1 : 0;
    p.pad0 = p.pad1 = 0;
    paths[i] = p;
    // every live path enters bounce 0; the queue is the identity (dead spectral samples are skipped)
    unsigned m = __ballot_sync(__activemask(), live);
    if (live) {
      unsigned lane = threadIdx.x & 31u;
      int leader = __ffs(m) - 1;
      unsigned long long slot = 0;
      if ((int)lane == leader) slot = atomicAdd(&q.counters[0], (unsigned long long)__popc(m));
      slot = __shfl_sync(m, slot, leader);
      q.cur[slot + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
    }
  }
}

// terminal value of a path: what the pixel loop adds for this sample
__device__ __forceinline__ void finish_path(const RenderParams& rp, PathState& p, d3 acc) {
  if (rp.sampler != IZPI_SAMPLER_SPECTRAL) {  // vec3.DeNAN (rgb.go:36, vec3.go:141-158)
    p.ax = isfinite(acc.x) ? acc.x : 0.0; p.ay = isfinite(acc.y) ? acc.y : 0.0; p.az = isfinite(acc.z) ? acc.z : 0.0;
  } else {  // render/spectral.go:91-95: XYZ += radiance * cmf / pdf, no DeNAN
    d3 cmf = cie_values(p.lambda);
    p.ax = (acc.x * cmf.x) / p.lpdf; p.ay = (acc.x * cmf.y) / p.lpdf; p.az = (acc.x * cmf.z) / p.lpdf;
  }
  p.alive = 0;
}

__device__ __forceinline__ d3 background_term(const RenderParams& rp, double lambda) {
  if (rp.sampler >= IZPI_SAMPLER_ALBEDO) return mk(0, 0, 0);  // albedo.go:35, normal.go:33: a miss is black
  if (rp.sampler == IZPI_SAMPLER_COLOUR) return mk(rp.background[0], rp.background[1], rp.background[2]);
  double v = rp.n_bg > 0 ? spd_value(rp.bg_w, rp.bg_v, rp.n_bg, lambda) : 0.0;  // colours.SpectralBlack
  return mk(v, 0, 0);
}

// ---- extend ---------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kThreads)
extend_kernel(const __grid_constant__ DScene sc, RenderParams rp, PathState* paths, Queues q) {
  extern __shared__ int32_t stack_smem[];
  int32_t* stack = stack_smem + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const long long n = (long long)q.counters[0];
  uint32_t nn = 0, np = 0;
  unsigned long long traced = 0;
  for (;;) {
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(&q.counters[7], 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    long long i = base + lane;
    bool hit = false;
    int cls = 0;
    int32_t pi = -1;
    if (i < n) {
      pi = q.cur[i];
      PathState& p = paths[pi];
      if (p.depth >= rp.max_depth && rp.sampler <= IZPI_SAMPLER_SPECTRAL) {
        // colour.go:34-36 returns (0,0,1); spectral.go:48-51 returns the background SPD
        d3 term = rp.sampler == IZPI_SAMPLER_COLOUR ? mk(0, 0, 1.0) : background_term(rp, p.lambda);
        d3 beta = mk(p.bx, p.by, p.bz), acc = mk(p.ax, p.ay, p.az);
        finish_path(rp, p, acc + hadamard(beta, term));
      } else {
        traced++;  // atomic.AddUint64(numRays, 1) (colour.go:38)
        DRay r = path_ray(p);
        double t = 0;
        int rec = world_closest<COUNT>(sc, r, 0.001, DBL_MAX, t, stack, kThreads, nn, np);
        if (rec < 0) {
          d3 beta = mk(p.bx, p.by, p.bz), acc = mk(p.ax, p.ay, p.az);
          finish_path(rp, p, acc + hadamard(beta, background_term(rp, p.lambda)));
        } else {
          p.hit_rec = rec; p.hit_t = t;
          hit = true;
          cls = sc.mat_bin[tag_material(sc.prims[rec].tag)];
        }
      }
    }
    __syncwarp();
    push_binned(hit, cls, pi, q.bins, q.capacity, q.counters + kBinCounters);
  }
  if (traced) atomicAdd(&q.counters[8], traced);
  if (COUNT) { atomicAdd(&q.counters[9], (unsigned long long)nn); atomicAdd(&q.counters[10], (unsigned long long)np); }
}

// Same stage with the 4-lanes-per-ray traversal (intersect_g4.cuh) for reference-shaped BVH4 worlds:
// 8 paths per warp, queue indices drawn in chunks, finished groups replaced immediately.
constexpr int kExtendChunk = 256;

template <bool COUNT>
__global__ void __launch_bounds__(kThreads, 7)  // 72 registers: seven resident blocks per SM (latency-bound kernel; 8 gains nothing, trace.cu)
extend_g4_kernel(const __grid_constant__ DScene sc, RenderParams rp, PathState* paths, Queues q, int stragglers) {
  extern __shared__ int2 g4_stack_smem[];
  const unsigned full = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const int j = lane & 3, gshift = (lane >> 2) * 4;
  int2* stack = g4_stack_smem + (size_t)(threadIdx.x >> 2) * kG4Slab;
  const int n = (int)q.counters[0];  // live paths of this bounce (a batch holds < 2^31 paths)
  // queue indices per atomicAdd: large when there is plenty of work, down to one packet of 8 when a late
  // bounce has only a few thousand live paths (otherwise a handful of warps would walk them serially)
  const int warps = (int)gridDim.x * (kThreads / 32);
  int chunk = (n / (warps * 4) + 7) & ~7;
  chunk = chunk < 8 ? 8 : (chunk > kExtendChunk ? kExtendChunk : chunk);
  uint32_t nn = 0, np = 0;
  unsigned traced = 0;
  int chunk_next = 0, chunk_end = 0;
  bool exhausted = false;
  G4State s;
  s.cur = kIdle;
  int32_t pi = -1;
  for (;;) {
    unsigned idle = __ballot_sync(full, s.cur == kIdle);
    if (idle) {
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&q.counters[7], (unsigned long long)chunk);
        b = __shfl_sync(full, b, 0);
        if (b >= (unsigned long long)n) { exhausted = true; chunk_next = chunk_end = 0; }
        else { chunk_next = (int)b; chunk_end = (int)b + chunk < n ? (int)b + chunk : n; }
      }
      int before = __popc(idle & ((1u << gshift) - 1u)) >> 2, total = __popc(idle) >> 2;
      if (s.cur == kIdle && chunk_next + before < chunk_end) {
        pi = q.cur[chunk_next + before];
        PathState& p = paths[pi];
        if (p.depth >= rp.max_depth && rp.sampler <= IZPI_SAMPLER_SPECTRAL) {  // colour.go:34-36 / spectral.go:48-51
          if (j == 0) {
            d3 term = rp.sampler == IZPI_SAMPLER_COLOUR ? mk(0, 0, 1.0) : background_term(rp, p.lambda);
            finish_path(rp, p, mk(p.ax, p.ay, p.az) + hadamard(mk(p.bx, p.by, p.bz), term));
          }
          pi = -1;
        } else {
          if (j == 0) traced++;  // atomic.AddUint64(numRays, 1) (colour.go:38)
          g4_begin(s, sc, path_ray(p), DBL_MAX, stack, j);
        }
      }
      int take = chunk_end - chunk_next;
      chunk_next += take < total ? take : total;
      if (exhausted && __ballot_sync(full, s.cur == kIdle && pi < 0) == full) break;
    }
    g4_node_phase<COUNT>(s, sc, stack, lane, gshift, j, nn, stragglers);
    g4_leaf_phase<COUNT>(s, sc, stack, lane, gshift, j, nn, np, 0.001);
    const bool done = s.cur == kIdle && pi >= 0;
    if (__any_sync(full, done)) {
      bool hit = false;
      int cls = 0;
      if (done && j == 0) {
        PathState& p = paths[pi];
        if (s.best < 0) {
          finish_path(rp, p, mk(p.ax, p.ay, p.az) + hadamard(mk(p.bx, p.by, p.bz), background_term(rp, p.lambda)));
        } else {
          p.hit_rec = s.best; p.hit_t = s.tmax;
          hit = true;
          cls = sc.mat_bin[tag_material(sc.prims[s.best].tag)];
        }
      }
      push_binned(hit, cls, pi, q.bins, q.capacity, q.counters + kBinCounters);
      if (done) pi = -1;
    }
  }
  if (traced) atomicAdd(&q.counters[8], (unsigned long long)traced);
  if (COUNT) {  // nodes are counted by the first lane of a ray's group, primitive tests by the lane that ran them (trace.cu)
    if (j == 0) atomicAdd(&q.counters[9], (unsigned long long)nn);
    atomicAdd(&q.counters[10], (unsigned long long)np);
  }
}


// The same stage with two lanes per path (intersect_g2.cuh): 16 paths per warp.  Used when the tree's worst-case stack fits
// the kG2Stack-entry slab (launch_cfg); trace.cu explains the trade.
constexpr int kExt2Threads = 64;  // two-warp blocks, 14 per SM at 72 registers (trace.cu explains the choice)

#ifndef IZPI_EXT2_MIN_BLOCKS
#define IZPI_EXT2_MIN_BLOCKS 14
#endif
template <int STACK, bool COUNT>
__global__ void __launch_bounds__(kExt2Threads, IZPI_EXT2_MIN_BLOCKS)
extend_g2_kernel(const __grid_constant__ DScene sc, RenderParams rp, PathState* paths, Queues q, int stragglers) {
  extern __shared__ int2 g4_stack_smem[];
  constexpr int kSlots = G2Slab<STACK>::kSlots;
  const unsigned full = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const int j = lane & 1, pshift = (int)(lane & ~1u);
  int2* stack = g4_stack_smem + (size_t)(threadIdx.x >> 1) * kSlots;
  const int n = (int)q.counters[0];  // live paths of this bounce (a batch holds < 2^31 paths)
  const int warps = (int)gridDim.x * (kExt2Threads / 32);
  // guided self-scheduling of queue chunks (trace.cu, trace_g2_kernel): short tails for the small late-bounce launches
  uint32_t nn = 0, np = 0;
  unsigned traced = 0;
  int chunk_next = 0, chunk_end = 0;
  bool exhausted = false;
  G4State s;
  s.cur = kIdle;
  int32_t pi = -1;
  for (;;) {
    unsigned idle = __ballot_sync(full, s.cur == kIdle);
    if (idle) {
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long b = 0;
        int chunk = 16;
        if (lane == 0) {
          const long long left = (long long)n - (long long)*reinterpret_cast<volatile unsigned long long*>(&q.counters[7]);
          long long want = left > 0 ? left / (2 * (long long)warps) : 0;
          chunk = want > kExtendChunk ? kExtendChunk : (want < 16 ? 16 : (int)((want + 15) & ~15ll));
          b = atomicAdd(&q.counters[7], (unsigned long long)chunk);
        }
        b = __shfl_sync(full, b, 0);
        chunk = __shfl_sync(full, chunk, 0);
        if (b >= (unsigned long long)n) { exhausted = true; chunk_next = chunk_end = 0; }
        else { chunk_next = (int)b; chunk_end = (int)b + chunk < n ? (int)b + chunk : n; }
      }
      int before = __popc(idle & ((1u << pshift) - 1u)) >> 1, total = __popc(idle) >> 1;
      if (s.cur == kIdle && chunk_next + before < chunk_end) {
        pi = q.cur[chunk_next + before];
        PathState& p = paths[pi];
        if (p.depth >= rp.max_depth && rp.sampler <= IZPI_SAMPLER_SPECTRAL) {  // colour.go:34-36 / spectral.go:48-51
          if (j == 0) {
            d3 term = rp.sampler == IZPI_SAMPLER_COLOUR ? mk(0, 0, 1.0) : background_term(rp, p.lambda);
            finish_path(rp, p, mk(p.ax, p.ay, p.az) + hadamard(mk(p.bx, p.by, p.bz), term));
          }
          pi = -1;
        } else {
          if (j == 0) traced++;  // atomic.AddUint64(numRays, 1) (colour.go:38)
          g2_begin<STACK>(s, sc, path_ray(p), DBL_MAX, stack, j);
        }
      }
      int take = chunk_end - chunk_next;
      chunk_next += take < total ? take : total;
      if (exhausted && __ballot_sync(full, s.cur == kIdle && pi < 0) == full) break;
    }
    g2_node_phase<COUNT, STACK>(s, sc, stack, pshift, j, nn, stragglers);
    g2_leaf_phase<COUNT, false, STACK>(s, sc, stack, pshift, j, nn, np, 0.001);
    const bool done = s.cur == kIdle && pi >= 0;
    if (__any_sync(full, done)) {
      bool hit = false;
      int cls = 0;
      if (done && j == 0) {
        PathState& p = paths[pi];
        if (s.best < 0) {
          finish_path(rp, p, mk(p.ax, p.ay, p.az) + hadamard(mk(p.bx, p.by, p.bz), background_term(rp, p.lambda)));
        } else {
          p.hit_rec = s.best; p.hit_t = s.tmax;
          hit = true;
          cls = sc.mat_bin[tag_material(sc.prims[s.best].tag)];
        }
      }
      push_binned(hit, cls, pi, q.bins, q.capacity, q.counters + kBinCounters);
      if (done) pi = -1;
    }
  }
  if (traced) atomicAdd(&q.counters[8], (unsigned long long)traced);
  if (COUNT) {  // nodes are counted by the first lane of a ray's group, primitive tests by the lane that ran them (trace.cu)
    if (j == 0) atomicAdd(&q.counters[9], (unsigned long long)nn);
    atomicAdd(&q.counters[10], (unsigned long long)np);
  }
}

// ---- shade ----------------------------------------------------------------------------------
// Dielectric.calculatePathLength (dielectric.go:119-153): nested closest hit from just inside the surface.
#ifndef IZPI_PATH_LENGTH_INLINE
#define IZPI_PATH_LENGTH_INLINE __forceinline__  // inlined: 72 B of spills instead of 112 B + call overhead (config 4: 104 -> 109 Msamples/s)
#endif
// `stack`: this thread's column of the block's shared-memory stack (stride kThreads), as deep as the uploaded tree can need
// (launch_cfg): the nested trace keeps no 64-entry array in local memory.
__device__ IZPI_PATH_LENGTH_INLINE double path_length(const DScene& sc, d3 hit_p, d3 sdir, double time, double lambda, int32_t* stack) {
  if (!sc.dielectric_has_world) return 10.0;
  DRay tr;
  tr.o = hit_p + sdir * 0.001; tr.d = sdir; tr.time = time; tr.lambda = lambda;
  uint32_t a = 0, b = 0;
  double t = 0;
  int rec = world_closest<false>(sc, tr, 0.0, 1000.0, t, stack, kThreads, a, b);
  if (rec < 0) return 10.0;
  PrimRec pr = load_rec(sc.prims + rec);
  DHit eh;
  prim_hit<true>(sc, rec, pr, tr, 0.0, DBL_MAX, eh, false);  // only the exit point is read
  double pl = len(eh.p - hit_p);
  if (pl < 0.1) pl = 0.1;
  if (pl > 100.0) pl = 100.0;
  return pl;
}

struct Scatter {
  bool ok, specular;
  d3 spec_dir;     // SpecularRay direction (origin = hit point)
  d3 atten;        // RGB attenuation, or (a,0,0) spectral
  d3 pdf_w;        // W axis of the Cosine pdf (already unit)
};

// pbr.go:59-156 / :158-263, the part shared by the RGB and the spectral scatter
__device__ __forceinline__ void pbr_common(const DScene& sc, const izpi_material_spec& m, const DRay& r, const DHit& h,
                                           Rng& rng, Scatter& s) {
  d3 normal;
  if (m.normal_tex >= 0) {
    d3 nuv = texture_value(sc, m.normal_tex, h.u, h.v);
    d3 tn = mk(2.0 * nuv.x - 1.0, 2.0 * nuv.y - 1.0, nuv.z);
    d3 n = h.n;
    d3 t = cross(n, mk(0, 1, 0));
    if (dot(t, t) < 0.001) t = cross(n, mk(1, 0, 0));
    t = unit(t);
    d3 b = unit(cross(n, t));
    normal = unit(mk(t.x * tn.x + b.x * tn.y + n.x * tn.z, t.y * tn.x + b.y * tn.y + n.y * tn.z,
                     t.z * tn.x + b.z * tn.y + n.z * tn.z));
  } else {
    normal = h.n;
  }
  d3 rough = m.roughness_tex >= 0 ? texture_value(sc, m.roughness_tex, h.u, h.v) : mk(0.5, 0.5, 0.5);
  d3 metal = m.metalness_tex >= 0 ? texture_value(sc, m.metalness_tex, h.u, h.v) : mk(0, 0, 0);
  double roughness = (rough.x + rough.y + rough.z) / 3.0;
  double metalness = (metal.x + metal.y + metal.z) / 3.0;
  Onb uvw = onb_from_w(normal);
  d3 ud = unit(r.d);
  d3 reflected = reflect(ud, normal);
  double cos_theta = fabs(dot(ud, normal));
  double fresnel = 0.04 + (1.0 - 0.04) * pow(1.0 - cos_theta, 5.0);
  fresnel = fresnel + (metalness * 0.5);
  double p_spec = fresnel * (1.0 - roughness);
  if (rnd(rng) < p_spec) {
    double rf = roughness * 0.3;
    if (!(0.01 < rf)) rf = 0.01;  // math.Max(0.01, rf)
    d3 rdir = random_in_unit_sphere(rng);
    s.spec_dir = unit(reflected + rdir * rf);
    s.specular = true;
  } else {
    (void)unit(onb_local(uvw, random_cosine_direction(rng)));  // finalDir: drawn, then ignored by the integrator
    s.specular = false;
  }
  s.pdf_w = uvw.w;
  s.ok = true;
}

// dielectric.go:66-102
__device__ __forceinline__ d3 dielectric_common(const DRay& r, const DHit& h, Rng& rng, double ref_idx, bool& reflected_out) {
  d3 outward;
  double ni_over_nt, cosine, reflect_prob;
  d3 reflected = reflect(r.d, h.n);
  double dn = dot(r.d, h.n);
  if (dn > 0) {
    outward = h.n * -1.0;
    ni_over_nt = ref_idx;
    cosine = ref_idx * dn / len(r.d);
  } else {
    outward = h.n;
    ni_over_nt = 1.0 / ref_idx;
    cosine = -dn / len(r.d);
  }
  d3 refracted = mk(0, 0, 0);
  if (refract(r.d, outward, ni_over_nt, refracted)) reflect_prob = schlick(cosine, ref_idx);
  else reflect_prob = 1.0;
  if (rnd(rng) < reflect_prob) { reflected_out = true; return reflected; }
  reflected_out = false;
  return refracted;
}

template <int CLS>
// Resident blocks per SM: without a bound the dielectric class takes 158 registers (3 blocks) and the others up to 118 (4);
// capping them at 128 / 96 registers spills a little but the extra warps win (config 4: 87 -> 96 Msamples/s with 4 blocks
// for the dielectric class; config 1: 401 -> 431 with 5 blocks for the others).
#ifndef IZPI_SHADE_MIN_BLOCKS_DIELECTRIC
#define IZPI_SHADE_MIN_BLOCKS_DIELECTRIC 4
#endif
#ifndef IZPI_SHADE_MIN_BLOCKS
#define IZPI_SHADE_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(kThreads, (CLS == IZPI_MAT_DIELECTRIC ? IZPI_SHADE_MIN_BLOCKS_DIELECTRIC : IZPI_SHADE_MIN_BLOCKS))
shade_kernel(const __grid_constant__ DScene sc, RenderParams rp, PathState* paths, Queues q) {
  // The bins of this class, one after the other: each bin's range is padded to whole warps, so a warp shades ONE material's
  // paths (one texture set) and stays in the loop for the ballots.
  extern __shared__ int32_t shade_stack[];  // dielectric class only: the nested path-length trace's stack (path_length)
  __shared__ long long s_begin[kMaxBins + 1];
  __shared__ int s_bin[kMaxBins];
  __shared__ int s_nb;
  if (threadIdx.x == 0) {
    int nb = 0;
    long long at = 0;
    for (int b = 0; b < sc.n_bins; b++) {
      if (sc.bin_class[b] != CLS) continue;
      s_bin[nb] = b; s_begin[nb] = at;
      at += ((long long)q.counters[kBinCounters + b] + 31) & ~31ll;
      nb++;
    }
    s_begin[nb] = at; s_nb = nb;
  }
  __syncthreads();
  const int nb = s_nb;
  const long long n_round = s_begin[nb];
  const bool spectral = rp.sampler == IZPI_SAMPLER_SPECTRAL;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_round; i += (long long)gridDim.x * blockDim.x) {
    bool survive = false;
    int32_t pi = -1;
    int k = 0;
    while (k + 1 < nb && i >= s_begin[k + 1]) k++;
    const int b = s_bin[k];
    const long long li = i - s_begin[k];
    if (li < (long long)q.counters[kBinCounters + b]) {
      pi = q.bins[(size_t)b * q.capacity + li];
      PathState& p = paths[pi];
      DRay r = path_ray(p);
      PrimRec pr = load_rec(sc.prims + p.hit_rec);
      const izpi_material_spec& m = sc.materials[tag_material(pr.tag)];
      DHit h;
      prim_hit<true>(sc, p.hit_rec, pr, r, 0.001, DBL_MAX, h, (sc.mat_flags[tag_material(pr.tag)] & kMatNeedsUV) != 0);
      h.t = p.hit_t;
      if (rp.sampler >= IZPI_SAMPLER_ALBEDO) {  // debug AOVs: one hit, no bounce (sampler/albedo.go:30-36, normal.go:28-34)
        d3 aov;
        if (rp.sampler == IZPI_SAMPLER_NORMAL) aov = h.n;
        else if (CLS == IZPI_MAT_METAL) aov = mk(m.v[0], m.v[1], m.v[2]);             // metal.go:50
        else if (CLS == IZPI_MAT_DIELECTRIC) aov = mk(1.0, 1.0, 1.0);                  // dielectric.go:223
        else aov = m.tex >= 0 ? texture_value(sc, m.tex, h.u, h.v) : mk(0, 0, 0);      // lambertian.go:84, diffuselight.go:70, pbr.go:281
        finish_path(rp, p, aov);
      } else {
      Rng rng;
      rng.key = p.key; rng.ctr = p.ctr;
      d3 beta = mk(p.bx, p.by, p.bz), acc = mk(p.ax, p.ay, p.az);
      Scatter s;
      s.ok = false; s.specular = false; s.spec_dir = mk(0, 0, 0); s.atten = mk(0, 0, 0); s.pdf_w = mk(0, 0, 1);
      d3 emitted = mk(0, 0, 0);
      if (CLS == IZPI_MAT_LAMBERT) {  // lambertian.go:44-71
        Onb uvw = onb_from_w(h.n);
        (void)onb_local(uvw, random_cosine_direction(rng));  // scatterCommon's ray: two draws, result unused
        s.pdf_w = uvw.w; s.ok = true; s.specular = false;
        if (spectral) s.atten = mk(m.spectral_tex >= 0 ? spectral_value(sc, m.spectral_tex, p.lambda, h.u, h.v) : 0.0, 0, 0);
        else s.atten = m.tex >= 0 ? texture_value(sc, m.tex, h.u, h.v) : mk(0, 0, 0);
      } else if (CLS == IZPI_MAT_METAL) {  // metal.go:34-41; spectral: non_spectral.go:18-20 -> no scatter
        if (!spectral) {
          d3 reflected = reflect(unit(r.d), h.n);
          s.spec_dir = reflected + random_in_unit_sphere(rng) * m.s;
          s.atten = mk(m.v[0], m.v[1], m.v[2]);
          s.ok = true; s.specular = true;
        }
      } else if (CLS == IZPI_MAT_DIELECTRIC) {  // dielectric.go:156-207
        double ref_idx = spectral ? spectral_value(sc, m.spectral_tex, p.lambda, h.u, h.v) : m.s;
        bool is_reflected;
        s.spec_dir = dielectric_common(r, h, rng, ref_idx, is_reflected);
        s.ok = true; s.specular = true;
        if (spectral) {
          double albedo = 1.0;
          if (!is_reflected) {
            double pl = path_length(sc, h.p, s.spec_dir, r.time, p.lambda, shade_stack + threadIdx.x);
            if (m.spectral_absorption_tex >= 0) albedo = exp(-spectral_value(sc, m.spectral_absorption_tex, p.lambda, h.u, h.v) * pl);
          }
          s.atten = mk(albedo, 0, 0);
        } else if (m.compute_beer_lambert && !(m.v[0] == 0 && m.v[1] == 0 && m.v[2] == 0) && !is_reflected) {
          double pl = path_length(sc, h.p, s.spec_dir, r.time, p.lambda, shade_stack + threadIdx.x);
          s.atten = mk(exp(-m.v[0] * pl), exp(-m.v[1] * pl), exp(-m.v[2] * pl));
        } else {
          s.atten = mk(1.0, 1.0, 1.0);
        }
      } else if (CLS == IZPI_MAT_DIFFUSE_LIGHT) {  // diffuselight.go:41-63
        if (dot(h.n, r.d) < 0.0) {
          if (spectral) emitted = mk(m.spectral_tex >= 0 ? spectral_value(sc, m.spectral_tex, p.lambda, h.u, h.v) : 0.0, 0, 0);
          else emitted = m.tex >= 0 ? texture_value(sc, m.tex, h.u, h.v) : mk(0, 0, 0);
        }
      } else {  // PBR
        if (spectral) {
          double albedo;
          if (m.spectral_tex >= 0) albedo = spectral_value(sc, m.spectral_tex, p.lambda, h.u, h.v);
          else { d3 rgb = texture_value(sc, m.tex, h.u, h.v); albedo = 0.299 * rgb.x + 0.587 * rgb.y + 0.114 * rgb.z; }
          pbr_common(sc, m, r, h, rng, s);
          s.atten = mk(s.specular ? albedo * 1.5 : albedo, 0, 0);
        } else {
          d3 albedo = texture_value(sc, m.tex, h.u, h.v);
          pbr_common(sc, m, r, h, rng, s);
          s.atten = albedo;
        }
      }

      // the integrator step (colour.go:41-61 / spectral.go:57-76) in iterative form:
      //   L = E + atten * L' * scatPDF / pdf   ->   acc += beta*E ; beta *= atten*scatPDF/pdf
      if (!s.ok) {
        finish_path(rp, p, acc + hadamard(beta, emitted));
      } else if (s.specular) {  // emitted is dropped on the specular branch
        beta = hadamard(beta, s.atten);
        p.ox = h.p.x; p.oy = h.p.y; p.oz = h.p.z; p.dx = s.spec_dir.x; p.dy = s.spec_dir.y; p.dz = s.spec_dir.z;
        survive = true;
      } else {
        d3 dir = rnd(rng) < 0.5 ? lights_random(sc, h.p, rng) : onb_local(onb_from_w(s.pdf_w), random_cosine_direction(rng));
        double pdf_val = 0.5 * lights_pdf_value(sc, h.p, dir) + 0.5 * cosine_pdf_value(s.pdf_w, dir);
        double cosine = dot(h.n, unit(dir));  // ScatteringPDF (lambertian.go:74-81, pbr.go:266-273)
        if (cosine < 0) cosine = 0;
        double spdf = cosine / M_PI;
        acc = acc + hadamard(beta, emitted);
        if (spectral) beta.x = beta.x * ((s.atten.x * spdf) / pdf_val);
        else beta = hadamard(beta, (s.atten * spdf) / pdf_val);
        p.ox = h.p.x; p.oy = h.p.y; p.oz = h.p.z; p.dx = dir.x; p.dy = dir.y; p.dz = dir.z;
        survive = true;
      }
      if (survive) {
        p.bx = beta.x; p.by = beta.y; p.bz = beta.z; p.ax = acc.x; p.ay = acc.y; p.az = acc.z;
        p.key = rng.key; p.ctr = rng.ctr; p.depth = p.depth + 1;
      }
      }  // integrators
    }
    const long long pos = push_binned(survive, 0, pi, q.next, q.capacity, q.counters + 1);
    if (pos >= 0 && q.next_keys) {
      const PathState& p = paths[pi];  // the scattered ray written above (still in L1)
      q.next_keys[pos] = coherence_key(sc, p.hit_rec, mk(p.ox, p.oy, p.oz), mk(p.dx, p.dy, p.dz));
    }
  }
}

// ---- coherence sort (tiny, specular scenes) ----------------------------------------------------
// After two bounces the live rays of a scene like config 4 are incoherent: the thread-per-ray traversal and the dielectric
// shade (whose nested path-length trace is a second traversal) then run at ~15 of 32 active lanes.  Sorting the queue by
// (surface the ray leaves, origin cell, direction cell) puts rays that walk the same nodes and hit the same primitive into
// the same warp.  Paths are independent and keyed by (pixel, sample), so the order in which a bounce processes them changes
// nothing in the image.  `n_bound` is the host's (stale, >=) count; entries past the device's own count get the last key.
__global__ void sort_pad_kernel(const unsigned long long* __restrict__ counters, int n_bound, uint32_t* __restrict__ keys) {
  const int i = (int)counters[0] + blockIdx.x * blockDim.x + threadIdx.x;  // entries past the device's own count sort to the end
  if (i < n_bound) keys[i] = 0xffffffffu;
}

// swap queues between bounces without a host round trip
__global__ void advance_kernel(Queues q, unsigned long long* host_visible_count) {
  q.counters[0] = q.counters[1];
  *host_visible_count = q.counters[1];
  q.counters[1] = 0;
  for (int b = 0; b < kMaxBins; b++) q.counters[kBinCounters + b] = 0;
  q.counters[7] = 0;
}

// ---- resolve ----------------------------------------------------------------------------------
// canvas holds running per-pixel SUMS (rgb.go:30-37 adds sample after sample); the row is ny - y and
// y == 0 falls outside the image (rgb.go:41; floatimage drops out-of-bounds Set calls).  The device canvas has one
// hidden row `height` that receives y == 0: the local path never copies it out, the worker path (izpi_render_tile_rows)
// streams it like any other row, as worker.RenderTile does (worker/render.go:33-69).
__global__ void resolve_kernel(RenderParams rp, const PathState* __restrict__ paths, const uint32_t* __restrict__ pixels,
                               int n_pixels, int s_count, double* canvas) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= n_pixels) return;
  uint32_t xy = pixels[pl];
  int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
  int row = rp.height - y;
  if (row < 0 || row > rp.height) return;
  double* px = canvas + ((size_t)row * rp.width + x) * 4;
  double a = px[0], b = px[1], c = px[2];
  const PathState* p = paths + (size_t)pl * s_count;
  for (int s = 0; s < s_count; s++) { a = a + p[s].ax; b = b + p[s].ay; c = c + p[s].az; }
  px[0] = a; px[1] = b; px[2] = c; px[3] = 1.0;
}

// mean over spp: RGB divides (rgb.go:39), spectral multiplies by 1/spp (render/spectral.go:99-103)
__global__ void mean_kernel(RenderParams rp, const double* __restrict__ sums, double* __restrict__ out, long long n_px) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  double a = sums[4 * i], b = sums[4 * i + 1], c = sums[4 * i + 2];
  if (rp.sampler != IZPI_SAMPLER_SPECTRAL) { a = a / (double)rp.spp; b = b / (double)rp.spp; c = c / (double)rp.spp; }
  else { double inv = 1.0 / (double)rp.spp; a = a * inv; b = b * inv; c = c * inv; }
  out[4 * i] = a; out[4 * i + 1] = b; out[4 * i + 2] = c; out[4 * i + 3] = sums[4 * i + 3];
}

// worker.RenderTile's reply rows (worker/render.go:33-69): row r = image row y0 + r (no flip), `stride` doubles per row of
// which the first 4 * width are the pixel means (RGB: sum / spp, worker/render.go:86; spectral: sum * (1/spp),
// render/spectral.go:99-103), alpha 1; the rest stays zero like the reference's make([]float64, stripSize).
__global__ void tile_rows_kernel(RenderParams rp, const double* __restrict__ sums, int x0, int y0, int w, int h, long long stride,
                                 double* __restrict__ rows) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)w * h) return;
  int r = (int)(i / w), c = (int)(i % w);
  const double* px = sums + ((size_t)(rp.height - (y0 + r)) * rp.width + (x0 + c)) * 4;
  double a = px[0], b = px[1], d = px[2];
  if (rp.sampler != IZPI_SAMPLER_SPECTRAL) { a = a / (double)rp.spp; b = b / (double)rp.spp; d = d / (double)rp.spp; }
  else { double inv = 1.0 / (double)rp.spp; a = a * inv; b = b * inv; d = d * inv; }
  double* o = rows + (size_t)r * stride + (size_t)c * 4;
  o[0] = a; o[1] = b; o[2] = d; o[3] = 1.0;
}

// spectral.FireflyRejection (firefly_rejection.go:12-113): 3x3 neighbourhood of a SNAPSHOT of Y
__global__ void firefly_kernel(const double* __restrict__ snap, double* __restrict__ pix, int width, int height) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= width || y >= height) return;
  size_t idx = ((size_t)y * width + x) * 4;
  double cur = snap[idx + 1];
  if (cur <= 0) return;
  double nb[8];
  int n = 0;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++) {
      if (dx == 0 && dy == 0) continue;
      int nx = x + dx, ny = y + dy;
      if (nx >= 0 && nx < width && ny >= 0 && ny < height) {
        double v = snap[((size_t)ny * width + nx) * 4 + 1];
        if (v > 0) nb[n++] = v;
      }
    }
  if (n < 3) return;
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += nb[i];
  double mean = sum / (double)n;
  double var = 0.0;
  for (int i = 0; i < n; i++) { double d = nb[i] - mean; var += d * d; }
  double stddev = sqrt(var / (double)n);
  double threshold = mean + 2.5 * stddev;
  if (cur > threshold && threshold > 0) {
    double ratio = threshold / cur;
    pix[idx] *= ratio; pix[idx + 1] *= ratio; pix[idx + 2] *= ratio;
  }
}

// spectral.XYZToRGB (rgb_image.go:13-17,28-67)
__global__ void xyz_to_acescg_kernel(double* pix, long long n_px, double exposure) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  double x = pix[4 * i] * exposure, y = pix[4 * i + 1] * exposure, z = pix[4 * i + 2] * exposure;
  pix[4 * i] = 1.6410234 * x + -0.3248033 * y + -0.2364247 * z;
  pix[4 * i + 1] = -0.6636629 * x + 1.6153316 * y + 0.0167563 * z;
  pix[4 * i + 2] = 0.0117219 * x + -0.0082845 * y + 0.9883949 * z;
}

}  // namespace

// ---- host side of the render entry points -------------------------------------------------------
// Work is cut into batches of (pixel block x sample block) paths.  kSlots batches are in flight at once,
// each on its own stream with its own path/queue buffers: while one batch runs its long tail of
// nearly empty bounces (paths caught between glass and metal live up to maxDepth = 50 bounces), the next
// one fills the SMs with its first, wide bounces.  Resolves are chained in batch order, so every pixel
// still adds its samples in the reference's order and the canvas is independent of the overlap.
//
// Where the tiles come from: a CURSOR over the caller's tile list.  A context that shares the list with other
// contexts (the members of a device group, or one-process-per-GPU ranks with the cursor in shared memory) claims the
// next run of tiles with one atomic fetch-add whenever a slot has room -- the reference's workers pull work units from
// one channel the same way (renderer.go:126-147) -- so nobody is handed a fixed share of cheap sky tiles or dear mesh tiles.
#ifndef IZPI_RENDER_SLOTS
#define IZPI_RENDER_SLOTS 2
#endif
constexpr int kSlots = IZPI_RENDER_SLOTS;
constexpr int kMaxBounces = 4096;

__global__ void accumulate_traced_kernel(const unsigned long long* counters, unsigned long long* total) {
  total[0] += counters[8]; total[1] += counters[9]; total[2] += counters[10];
}

struct BatchSlot {
  PathState* d_paths = nullptr;
  Queues q{};
  uint32_t* d_pixels = nullptr;
  int64_t pixel_capacity = 0;
  uint32_t *d_keys_cur = nullptr, *d_keys_alt = nullptr;  // coherence sort: keys of the `cur` queue, CUB's key output (q.next_keys is the third buffer)
  int32_t* d_sorted = nullptr;                            // third queue buffer: where the sorted order lands
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int64_t sort_capacity = 0;
  unsigned long long* h_count = nullptr;  // pinned + mapped: advance_kernel writes the live count of bounce b to [b & 7]
  unsigned long long* d_count_mapped = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t resolved = nullptr;
  // run state
  bool busy = false, drained = false;
  int batch = -1, n_pixels = 0, s_begin = 0, s_count = 0, bounce = 0;
  unsigned long long live = 0;
};

struct TileRun { uint32_t x0, y0, x1, y1; };  // horizontally merged tiles rendered by this context since the last setup

struct RenderState {
  izpi_render_config cfg{};
  RenderParams rp{};
  double* d_canvas = nullptr;   // running sums, 4*W*H
  double* d_out = nullptr;      // means / epilogue scratch
  double* d_snap = nullptr;
  size_t canvas_capacity = 0;
  double* d_bg = nullptr;
  unsigned long long* d_total_rays = nullptr;  // [0] rays traced, [1] nodes visited, [2] primitive tests (counting kernels)
  BatchSlot slot[kSlots];
  int64_t batch_paths_env = 0;  // IZPI_BATCH_PATHS: paths per batch (default: chosen per scene in setup_one)
  bool allocated = false;
  int bins_allocated = 0;
  std::vector<TileRun> rendered;  // what izpi_render_finish of a device group has to move
  long long rendered_pixels = 0;
  // IZPI_RENDER_STATS: per-stage CUDA-event time (batches serialised) and counting extend kernels
  bool stats_on = false;  // events around the stages, one batch in flight
  bool count_on = false;  // counting extend kernels
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { cudaEvent_t a, b; int kind; };  // kind 0 extend, 1 shade + advance, 2 raygen / resolve
  std::vector<Span> spans;
  izpi_render_stats stats{};
};

void render_state_free(izpi_ctx* ctx) {
  RenderState* r = ctx->render;
  if (!r) return;
  cudaFree(r->d_canvas); cudaFree(r->d_out); cudaFree(r->d_snap); cudaFree(r->d_bg); cudaFree(r->d_total_rays);
  for (BatchSlot& s : r->slot) {
    cudaFree(s.d_paths); cudaFree(s.d_pixels); cudaFree(s.q.cur); cudaFree(s.q.next); cudaFree(s.q.bins); cudaFree(s.q.counters);
    cudaFree(s.d_keys_cur); cudaFree(s.d_keys_alt); cudaFree(s.q.next_keys); cudaFree(s.d_sorted); cudaFree(s.d_sort_tmp);
    if (s.h_count) cudaFreeHost(s.h_count);
    for (cudaEvent_t e : s.ev) if (e) cudaEventDestroy(e);
    if (s.resolved) cudaEventDestroy(s.resolved);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (cudaEvent_t e : r->ev_pool) cudaEventDestroy(e);
  delete r;
  ctx->render = nullptr;
}

namespace {

#include "cie_tables.inc"

// c_cie is __constant__ memory: one copy per DEVICE, so the upload is tracked per context (a thread that owns contexts on
// two GPUs must fill both)
int upload_cie(izpi_ctx* ctx) {
  if (ctx->cie_uploaded) return IZPI_OK;
  IZ_CUDA(cudaMemcpyToSymbol(c_cie, kCieTable, sizeof(kCieTable)));
  ctx->cie_uploaded = true;
  return IZPI_OK;
}

template <typename K, typename... Args>
int launch(izpi_ctx* ctx, cudaStream_t st, K kern, dim3 grid, dim3 block, size_t smem, Args... args) {
  kern<<<grid, block, smem, st>>>(args...);
  IZ_CUDA(cudaGetLastError());
  ctx->launches++;
  return IZPI_OK;
}

struct LaunchCfg {
  bool use_g4, use_g2, g2_deep, count, sort_rays;
  long long sort_min;
  size_t smem, smem4, smem2;
  int ext_blocks, ext4_blocks, ext2_blocks;
};

int launch_cfg(izpi_ctx* ctx, LaunchCfg& lc) {
  // thread-per-ray stage: one int32 stack per thread, as deep as THIS tree can need (a 22-primitive scene needs 4 entries, not
  // the reference's 64: 2 KB of shared memory per block instead of 32 KB, so occupancy is set by registers alone)
  const int depth = ctx->scene.world_kind == IZPI_WORLD_BVH4 ? std::min(kStackDepth, std::max(1, ctx->scene.scalar_need + 1)) : 1;
  lc.smem = (size_t)depth * kThreads * sizeof(int32_t);
  lc.smem4 = (size_t)(kThreads / 4) * kG4Slab * sizeof(int2);
  lc.g2_deep = ctx->scene.g4_need > kG2Stack;  // the 52-entry slab (intersect_g2.cuh)
  const size_t smem2_a = (size_t)(kExt2Threads / 2) * G2Slab<kG2Stack>::kSlots * sizeof(int2);
  const size_t smem2_b = (size_t)(kExt2Threads / 2) * G2Slab<kG2StackDeep>::kSlots * sizeof(int2);
  lc.smem2 = lc.g2_deep ? smem2_b : smem2_a;
  OccupancyCache& oc = ctx->occ;  // per context: function attributes and occupancy belong to a device
  if (!oc.ext || oc.ext_smem != lc.smem) {
    oc.ext_smem = lc.smem;
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc.ext2[0], extend_g2_kernel<kG2Stack, false>, kExt2Threads, smem2_a));
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc.ext2[1], extend_g2_kernel<kG2StackDeep, false>, kExt2Threads, smem2_b));
    if (oc.ext2[0] < 1) oc.ext2[0] = 1;
    if (oc.ext2[1] < 1) oc.ext2[1] = 1;
    IZ_CUDA(cudaFuncSetAttribute(extend_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kStackDepth * kThreads * sizeof(int32_t))));
    IZ_CUDA(cudaFuncSetAttribute(extend_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kStackDepth * kThreads * sizeof(int32_t))));
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc.ext4, extend_g4_kernel<false>, kThreads, lc.smem4));
    if (oc.ext4 < 1) oc.ext4 = 1;
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc.ext, extend_kernel<false>, kThreads, lc.smem));
    if (oc.ext < 1) oc.ext = 1;
  }
  lc.ext_blocks = oc.ext; lc.ext4_blocks = oc.ext4; lc.ext2_blocks = oc.ext2[lc.g2_deep ? 1 : 0];
  // tiny trees (config 4 has 22 primitives) stay cache-resident and coherent: the thread-per-ray stage wins there
  int coop_min_nodes = 64;
  if (const char* e = getenv("IZPI_COOP_MIN_NODES")) coop_min_nodes = atoi(e);
  lc.use_g4 = ctx->scene.world_kind == IZPI_WORLD_BVH4 && ctx->scene.g4_ok && !ctx->force_scalar && ctx->scene.n_nodes >= coop_min_nodes;
  lc.use_g2 = lc.use_g4 && ctx->trace_lanes == 2 && ctx->scene.g4_need <= kG2StackDeep;
  lc.count = ctx->render && ctx->render->count_on;
  // coherence sort: thread-per-ray BVH worlds only (the cooperative kernels do not need it, slices have no origin cells)
  lc.sort_rays = !lc.use_g4 && ctx->scene.world_kind == IZPI_WORLD_BVH4;
  lc.sort_min = 1 << 18;
  if (const char* e = getenv("IZPI_SORT_RAYS")) { long v = atol(e); if (v <= 0) lc.sort_rays = false; else if (v > 1) lc.sort_min = v; }
  return IZPI_OK;
}

// IZPI_RENDER_STATS: events around the stages of a bounce
int span_begin(RenderState* r, cudaStream_t st, int kind) {
  if (!r->stats_on) return IZPI_OK;
  while (r->ev_used + 2 > r->ev_pool.size()) {
    cudaEvent_t e;
    IZ_CUDA(cudaEventCreate(&e));
    r->ev_pool.push_back(e);
  }
  RenderState::Span sp{r->ev_pool[r->ev_used], r->ev_pool[r->ev_used + 1], kind};
  r->ev_used += 2;
  IZ_CUDA(cudaEventRecord(sp.a, st));
  r->spans.push_back(sp);
  return IZPI_OK;
}
int span_end(RenderState* r, cudaStream_t st) {
  if (!r->stats_on) return IZPI_OK;
  IZ_CUDA(cudaEventRecord(r->spans.back().b, st));
  return IZPI_OK;
}

// first launches of a batch: pixel list, counters, ray generation
int batch_start(izpi_ctx* ctx, RenderState* r, BatchSlot& s, int batch, const uint32_t* px, int n_pixels, int s_begin, int s_count) {
  cudaStream_t st = s.stream;
  if (n_pixels > s.pixel_capacity) {
    IZ_CUDA(cudaStreamSynchronize(st));
    cudaFree(s.d_pixels); s.d_pixels = nullptr; s.pixel_capacity = 0;
    IZ_CUDA(cudaMalloc(&s.d_pixels, (size_t)n_pixels * 4));
    s.pixel_capacity = n_pixels;
  }
  IZ_CUDA(cudaMemcpyAsync(s.d_pixels, px, (size_t)n_pixels * 4, cudaMemcpyHostToDevice, st));
  IZ_CUDA(cudaMemsetAsync(s.q.counters, 0, (kBinCounters + kMaxBins) * sizeof(unsigned long long), st));
  long long n = (long long)n_pixels * s_count;
  int gen_grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
  int rc = span_begin(r, st, 2);
  if (rc != IZPI_OK) return rc;
  rc = launch(ctx, st, raygen_kernel, dim3(gen_grid), dim3(256), 0, ctx->scene, r->rp, s.d_paths, s.d_pixels, n_pixels, s_begin,
              s_count, s.q);
  if (rc != IZPI_OK) return rc;
  if ((rc = span_end(r, st)) != IZPI_OK) return rc;
  s.busy = true; s.drained = false; s.batch = batch; s.n_pixels = n_pixels; s.s_begin = s_begin; s.s_count = s_count;
  s.bounce = 0; s.live = (unsigned long long)n;
  return IZPI_OK;
}

template <int CLS>
int launch_shade(izpi_ctx* ctx, RenderState* r, BatchSlot& s, int grid) {
  if (!((ctx->scene.class_mask >> CLS) & 1)) return IZPI_OK;  // no primitive carries this class: its bin stays empty
  size_t smem = 0;
  if (CLS == IZPI_MAT_DIELECTRIC && ctx->scene.world_kind == IZPI_WORLD_BVH4)  // stack of the nested trace, sized like the extend stage's
    smem = (size_t)std::min(kStackDepth, std::max(1, ctx->scene.scalar_need + 1)) * kThreads * sizeof(int32_t);
  return launch(ctx, s.stream, shade_kernel<CLS>, dim3(grid), dim3(kThreads), smem, ctx->scene, r->rp, s.d_paths, s.q);
}

// One bounce of a batch.  The live count of bounce b-2 is harvested first (the host runs at most two bounces ahead
// of the device); counts only shrink, so the stale value is a valid upper bound for the grid size.
int batch_step(izpi_ctx* ctx, RenderState* r, BatchSlot& s, const LaunchCfg& lc) {
  cudaStream_t st = s.stream;
  const int sm = ctx->sm_count;
  if (s.bounce >= 2) {
    IZ_CUDA(cudaEventSynchronize(s.ev[(s.bounce - 2) & 7]));
    s.live = s.h_count[(s.bounce - 2) & 7];
  }
  if (s.live == 0 || s.bounce > r->rp.max_depth || s.bounce >= kMaxBounces) { s.drained = true; return IZPI_OK; }
  int rc;
  if (lc.sort_rays && s.d_sorted && s.bounce >= 2 && s.live >= (unsigned long long)lc.sort_min) {
    // Coherence sort of this bounce's queue.  Its keys were written by the previous bounce's shade kernels next to the queue
    // entries (Queues::next_keys, now s.d_keys_cur); entries between the device's own count and the host's stale bound get
    // the largest key.  Two onesweep passes over (16-bit key, path index) pairs.
    const int nb = (int)std::min<unsigned long long>(s.live, (unsigned long long)s.q.capacity);
    if ((rc = span_begin(r, st, 2)) != IZPI_OK) return rc;
    if ((rc = launch(ctx, st, sort_pad_kernel, dim3((nb + 255) / 256), dim3(256), 0, s.q.counters, nb, s.d_keys_cur)) != IZPI_OK) return rc;
    size_t tmp = s.sort_tmp_bytes;
    IZ_CUDA(cub::DeviceRadixSort::SortPairs(s.d_sort_tmp, tmp, s.d_keys_cur, s.d_keys_alt, s.q.cur, s.d_sorted, nb, 0, kSortKeyBits, st));
    ctx->launches += 3;  // histogram + two onesweep passes
    std::swap(s.q.cur, s.d_sorted);  // the sorted order is this bounce's queue; the old buffer becomes the next sort's output
    if ((rc = span_end(r, st)) != IZPI_OK) return rc;
  }
  if ((rc = span_begin(r, st, 0)) != IZPI_OK) return rc;
  long long want = ((long long)s.live + kThreads - 1) / kThreads;
  if (lc.use_g2) {
    long long want2 = ((long long)s.live + (kExt2Threads / 2) - 1) / (kExt2Threads / 2);
    int eg2 = (int)std::max<long long>(1, std::min<long long>(want2, (long long)sm * lc.ext2_blocks));
    auto k = lc.g2_deep ? (lc.count ? extend_g2_kernel<kG2StackDeep, true> : extend_g2_kernel<kG2StackDeep, false>)
                        : (lc.count ? extend_g2_kernel<kG2Stack, true> : extend_g2_kernel<kG2Stack, false>);
    if ((rc = launch(ctx, st, k, dim3(eg2), dim3(kExt2Threads), lc.smem2, ctx->scene, r->rp, s.d_paths, s.q, ctx->pair_stragglers)) != IZPI_OK) return rc;
  } else if (lc.use_g4) {
    long long want4 = ((long long)s.live + (kThreads / 4) - 1) / (kThreads / 4);
    int eg4 = (int)std::max<long long>(1, std::min<long long>(want4, (long long)sm * lc.ext4_blocks));
    if ((rc = launch(ctx, st, lc.count ? extend_g4_kernel<true> : extend_g4_kernel<false>, dim3(eg4), dim3(kThreads), lc.smem4, ctx->scene, r->rp,
                     s.d_paths, s.q, ctx->node_stragglers)) != IZPI_OK) return rc;
  } else {
    int eg = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm * lc.ext_blocks));
    if ((rc = launch(ctx, st, lc.count ? extend_kernel<true> : extend_kernel<false>, dim3(eg), dim3(kThreads), lc.smem, ctx->scene, r->rp, s.d_paths,
                     s.q)) != IZPI_OK) return rc;
  }
  if ((rc = span_end(r, st)) != IZPI_OK) return rc;
  if ((rc = span_begin(r, st, 1)) != IZPI_OK) return rc;
  int sg = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm * 8));
  if ((rc = launch_shade<IZPI_MAT_LAMBERT>(ctx, r, s, sg)) != IZPI_OK) return rc;
  if ((rc = launch_shade<IZPI_MAT_METAL>(ctx, r, s, sg)) != IZPI_OK) return rc;
  if ((rc = launch_shade<IZPI_MAT_DIELECTRIC>(ctx, r, s, sg)) != IZPI_OK) return rc;
  if ((rc = launch_shade<IZPI_MAT_DIFFUSE_LIGHT>(ctx, r, s, sg)) != IZPI_OK) return rc;
  if ((rc = launch_shade<IZPI_MAT_PBR>(ctx, r, s, sg)) != IZPI_OK) return rc;
  std::swap(s.q.cur, s.q.next);
  std::swap(s.d_keys_cur, s.q.next_keys);  // the survivors' keys follow their queue
  Queues qs = s.q;  // after the swap: cur = survivors; advance moves the count
  if ((rc = launch(ctx, st, advance_kernel, dim3(1), dim3(1), 0, qs, s.d_count_mapped + (s.bounce & 7))) != IZPI_OK) return rc;
  if ((rc = span_end(r, st)) != IZPI_OK) return rc;
  IZ_CUDA(cudaEventRecord(s.ev[s.bounce & 7], st));
  s.bounce++;
  return IZPI_OK;
}

// last launches of a batch: per-pixel sums in sample order, after the previous batch's resolve
int batch_finish(izpi_ctx* ctx, RenderState* r, BatchSlot& s, cudaEvent_t prev_resolved) {
  cudaStream_t st = s.stream;
  if (prev_resolved) IZ_CUDA(cudaStreamWaitEvent(st, prev_resolved, 0));
  int rc = span_begin(r, st, 2);
  if (rc != IZPI_OK) return rc;
  rc = launch(ctx, st, resolve_kernel, dim3((s.n_pixels + 127) / 128), dim3(128), 0, r->rp, s.d_paths, s.d_pixels, s.n_pixels,
              s.s_count, r->d_canvas);
  if (rc != IZPI_OK) return rc;
  if ((rc = launch(ctx, st, accumulate_traced_kernel, dim3(1), dim3(1), 0, s.q.counters, r->d_total_rays)) != IZPI_OK) return rc;
  if ((rc = span_end(r, st)) != IZPI_OK) return rc;
  IZ_CUDA(cudaEventRecord(s.resolved, st));
  s.busy = false;
  return IZPI_OK;
}

int sync_slots(RenderState* r) {
  for (BatchSlot& s : r->slot) IZ_CUDA(cudaStreamSynchronize(s.stream));
  return IZPI_OK;
}

// device side of the frame's tail: mean over spp into d_out, then the spectral epilogue (renderer.go:216-219)
int canvas_epilogue(izpi_ctx* ctx, RenderState* r, bool epilogue) {
  cudaStream_t st = ctx->stream;
  long long n_px = (long long)r->rp.width * r->rp.height;
  int rc;
  if ((rc = sync_slots(r)) != IZPI_OK) return rc;
  if ((rc = launch(ctx, st, mean_kernel, dim3((unsigned)((n_px + 255) / 256)), dim3(256), 0, r->rp, r->d_canvas, r->d_out, n_px)) != IZPI_OK) return rc;
  if (epilogue && r->rp.sampler == IZPI_SAMPLER_SPECTRAL) {
    IZ_CUDA(cudaMemcpyAsync(r->d_snap, r->d_out, (size_t)n_px * 32, cudaMemcpyDeviceToDevice, st));
    dim3 b(32, 8), g((r->rp.width + 31) / 32, (r->rp.height + 7) / 8);
    if ((rc = launch(ctx, st, firefly_kernel, g, b, 0, r->d_snap, r->d_out, r->rp.width, r->rp.height)) != IZPI_OK) return rc;
    if ((rc = launch(ctx, st, xyz_to_acescg_kernel, dim3((unsigned)((n_px + 255) / 256)), dim3(256), 0, r->d_out, n_px,
                     ctx->scene.camera.exposure)) != IZPI_OK) return rc;
  }
  return IZPI_OK;
}

int canvas_to_host(izpi_ctx* ctx, RenderState* r, double* host, bool epilogue) {
  int rc = canvas_epilogue(ctx, r, epilogue);
  if (rc != IZPI_OK) return rc;
  long long n_px = (long long)r->rp.width * r->rp.height;
  IZ_CUDA(cudaMemcpyAsync(host, r->d_out, (size_t)n_px * 32, cudaMemcpyDeviceToHost, ctx->stream));
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return IZPI_OK;
}

// Rows of the device / host canvas that hold the pixels of a run of tiles: image row y lives at canvas row H - y
// (rgb.go:41); y == 0 falls into the hidden row H, which never leaves the device on this path.
bool run_rows(const RenderState* r, const TileRun& t, int& row0, int& rows) {
  const int H = r->rp.height;
  int lo = H - (int)t.y1, hi = H - (int)t.y0;  // inclusive canvas rows
  if (hi > H - 1) hi = H - 1;
  if (lo < 0) lo = 0;
  row0 = lo; rows = hi - lo + 1;
  return rows > 0;
}

int setup_one(izpi_ctx* ctx, const izpi_render_config* cfg) {
  IZ_CUDA(cudaSetDevice(ctx->device));
  int rc = upload_cie(ctx);
  if (rc != IZPI_OK) return rc;
  // the big buffers (path states, queues, canvas) are allocated once per context and reused by later
  // setups: a render call then costs no cudaMalloc / cudaFree
  RenderState* r = ctx->render;
  if (!r) {
    r = new RenderState();
    ctx->render = r;
    const char* e = getenv("IZPI_BATCH_PATHS");
    if (e) { long v = atol(e); if (v >= 1024 && v <= (1l << 28)) r->batch_paths_env = v; }
  }
  r->cfg = *cfg;
  r->stats_on = (cfg->flags & (IZPI_RENDER_STATS | IZPI_RENDER_TIMING)) != 0;
  r->count_on = (cfg->flags & IZPI_RENDER_STATS) != 0;
  r->spans.clear(); r->ev_used = 0;
  std::memset(&r->stats, 0, sizeof(r->stats));
  r->rendered.clear(); r->rendered_pixels = 0;
  RenderParams& rp = r->rp;
  rp.width = cfg->width; rp.height = cfg->height; rp.spp = cfg->spp; rp.max_depth = cfg->max_depth; rp.sampler = cfg->sampler;
  for (int k = 0; k < 3; k++) rp.background[k] = cfg->background[k];
  rp.seed = cfg->seed; rp.n_bg = 0; rp.bg_w = rp.bg_v = nullptr;
  if (cfg->n_bg > 0 && cfg->bg_wavelengths && cfg->bg_values) {
    cudaFree(r->d_bg); r->d_bg = nullptr;
    IZ_CUDA(cudaMalloc(&r->d_bg, (size_t)cfg->n_bg * 16));
    IZ_CUDA(cudaMemcpy(r->d_bg, cfg->bg_wavelengths, (size_t)cfg->n_bg * 8, cudaMemcpyHostToDevice));
    IZ_CUDA(cudaMemcpy(r->d_bg + cfg->n_bg, cfg->bg_values, (size_t)cfg->n_bg * 8, cudaMemcpyHostToDevice));
    rp.n_bg = cfg->n_bg; rp.bg_w = r->d_bg; rp.bg_v = r->d_bg + cfg->n_bg;
  }
  size_t n_px = (size_t)cfg->width * cfg->height;
  const size_t n_px_hidden = n_px + (size_t)cfg->width;  // + the hidden row of y == 0 (resolve_kernel)
  if (n_px_hidden > r->canvas_capacity) {
    cudaFree(r->d_canvas); cudaFree(r->d_out); cudaFree(r->d_snap);
    r->d_canvas = r->d_out = r->d_snap = nullptr; r->canvas_capacity = 0;
    IZ_CUDA(cudaMalloc(&r->d_canvas, n_px_hidden * 32));
    IZ_CUDA(cudaMalloc(&r->d_out, n_px_hidden * 32));
    IZ_CUDA(cudaMalloc(&r->d_snap, n_px_hidden * 32));
    r->canvas_capacity = n_px_hidden;
  }
  if (!r->allocated) {
    IZ_CUDA(cudaMalloc(&r->d_total_rays, 4 * sizeof(unsigned long long)));
    for (BatchSlot& s : r->slot) {
      IZ_CUDA(cudaMalloc(&s.q.counters, (kBinCounters + kMaxBins) * sizeof(unsigned long long)));
      IZ_CUDA(cudaHostAlloc(&s.h_count, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
      IZ_CUDA(cudaHostGetDevicePointer(&s.d_count_mapped, s.h_count, 0));
      IZ_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      for (cudaEvent_t& e : s.ev) IZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      IZ_CUDA(cudaEventCreateWithFlags(&s.resolved, cudaEventDisableTiming));
    }
    r->allocated = true;
  }
  rc = sync_slots(r);
  if (rc != IZPI_OK) return rc;
  // Paths per batch.  Every batch ends in ~50 nearly empty bounces (the few paths caught between glass and metal), whose cost
  // is latency, not work: the fewer batches a frame has, the smaller their share.  Large trees, where one bounce of a full
  // batch takes tens of milliseconds, gain 6-12 % from 64 M paths per batch over 16 M (config 3: 214 -> 241 Msamples/s, config 5:
  // 169 -> 189, profiles/r02_tune_batch.log); the tiny scenes of the thread-per-ray stage do not (config 4: -1.5 %).  The
  // buffers are sized for the frame at hand (160 B of path state + queues per path: 11.6 GB per slot at 64 M) and only grow.
  LaunchCfg lc0;
  if ((rc = launch_cfg(ctx, lc0)) != IZPI_OK) return rc;
  int64_t want_cap = r->batch_paths_env ? r->batch_paths_env : (lc0.use_g4 ? (int64_t)1 << 26 : (int64_t)1 << 24);
  const int64_t frame_paths = (int64_t)n_px * (int64_t)std::max(1, cfg->spp);  // by spp, not sample_count: a one-sample warm-up of the frame allocates the frame's buffers
  if (frame_paths < want_cap) want_cap = std::max<int64_t>(frame_paths, 1 << 16);
  if (want_cap > r->slot[0].q.capacity || ctx->scene.n_bins > r->bins_allocated) {
    const int32_t cap = (int32_t)std::max<int64_t>(want_cap, r->slot[0].q.capacity);
    for (BatchSlot& s : r->slot) {
      if (cap > s.q.capacity) {
        cudaFree(s.d_paths); cudaFree(s.q.cur); cudaFree(s.q.next);
        s.d_paths = nullptr; s.q.cur = s.q.next = nullptr; s.q.capacity = 0;
        IZ_CUDA(cudaMalloc(&s.d_paths, (size_t)cap * sizeof(PathState)));
        IZ_CUDA(cudaMalloc(&s.q.cur, (size_t)cap * 4));
        IZ_CUDA(cudaMalloc(&s.q.next, (size_t)cap * 4));
        s.q.capacity = cap;
      }
      // one bin per image-textured material: the count belongs to the uploaded scene
      cudaFree(s.q.bins); s.q.bins = nullptr;
      IZ_CUDA(cudaMalloc(&s.q.bins, (size_t)s.q.capacity * 4 * (size_t)std::max(ctx->scene.n_bins, r->bins_allocated)));
    }
    r->bins_allocated = std::max(ctx->scene.n_bins, r->bins_allocated);
  }
  for (BatchSlot& s : r->slot) {  // coherence-sort buffers: only for scenes that are sorted, same capacity as the queues
    if (lc0.sort_rays && s.sort_capacity < (int64_t)s.q.capacity) {
      cudaFree(s.d_keys_cur); cudaFree(s.d_keys_alt); cudaFree(s.q.next_keys); cudaFree(s.d_sorted); cudaFree(s.d_sort_tmp);
      s.d_keys_cur = s.d_keys_alt = s.q.next_keys = nullptr; s.d_sorted = nullptr; s.d_sort_tmp = nullptr; s.sort_capacity = 0;
      const size_t cap = (size_t)s.q.capacity;
      IZ_CUDA(cudaMalloc(&s.d_keys_cur, cap * 4)); IZ_CUDA(cudaMalloc(&s.d_keys_alt, cap * 4)); IZ_CUDA(cudaMalloc(&s.q.next_keys, cap * 4));
      IZ_CUDA(cudaMalloc(&s.d_sorted, cap * 4));
      s.sort_tmp_bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, s.sort_tmp_bytes, s.d_keys_cur, s.d_keys_alt, s.q.cur, s.d_sorted, (int)cap, 0, kSortKeyBits, s.stream);
      IZ_CUDA(cudaMalloc(&s.d_sort_tmp, s.sort_tmp_bytes));
      s.sort_capacity = (int64_t)cap;
    }
    if (!lc0.sort_rays && s.q.next_keys) {  // a scene that is not sorted: the shade kernels must not write keys
      cudaFree(s.d_keys_cur); cudaFree(s.d_keys_alt); cudaFree(s.q.next_keys); cudaFree(s.d_sorted); cudaFree(s.d_sort_tmp);
      s.d_keys_cur = s.d_keys_alt = s.q.next_keys = nullptr; s.d_sorted = nullptr; s.d_sort_tmp = nullptr; s.sort_capacity = 0;
    }
  }
  IZ_CUDA(cudaMemsetAsync(r->d_canvas, 0, n_px_hidden * 32, ctx->stream));
  IZ_CUDA(cudaMemsetAsync(r->d_total_rays, 0, 4 * sizeof(unsigned long long), ctx->stream));
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return IZPI_OK;
}

// The render loop of ONE device over a tile list whose cursor may be shared with other contexts.
int tiles_one(izpi_ctx* ctx, int32_t n_tiles, const uint32_t* tiles, uint64_t* cursor, int sharers) {
  RenderState* r = ctx->render;
  IZ_CUDA(cudaSetDevice(ctx->device));
  const int s_total = r->cfg.sample_count, s_off = r->cfg.sample_offset;
  if (n_tiles == 0 || s_total <= 0) return IZPI_OK;
  uint64_t private_cursor = 0;
  if (!cursor) { cursor = &private_cursor; sharers = 1; }
  if (sharers < 1) sharers = 1;
  const int64_t cap = r->slot[0].q.capacity;
  struct Batch { std::vector<uint32_t>* px; int64_t p0; int np, s0, sc; };
  std::deque<std::vector<uint32_t>> claims;  // pixel lists stay alive until the frame is done
  std::deque<Batch> pending;
  bool exhausted = false;
  // Claim the next run of tiles (izpi_host_claim_tiles): everything at once for a private cursor (then split as before),
  // otherwise guided self-scheduling.
  auto claim = [&]() {
    const long long tile_paths = (long long)(tiles[2] - tiles[0] + 1) * (tiles[3] - tiles[1] + 1) * s_total;
    int32_t t0 = 0, t1 = 0;
    if (!izpi_host_claim_tiles(cursor, n_tiles, tile_paths, cap, sharers == 1 ? 1 : sharers * kSlots, &t0, &t1)) { exhausted = true; return; }
    // pixel list in the reference's loop order: tiles as given, rows y0..y1, columns x0..x1 (rgb.go:27-29)
    claims.emplace_back();
    std::vector<uint32_t>& px = claims.back();
    for (int t = t0; t < t1; t++) {
      const uint32_t x0 = tiles[4 * t], y0 = tiles[4 * t + 1], x1 = tiles[4 * t + 2], y1 = tiles[4 * t + 3];
      for (uint32_t y = y0; y <= y1; y++)
        for (uint32_t x = x0; x <= x1; x++) px.push_back(x | (y << 16));
      if (!r->rendered.empty() && r->rendered.back().y0 == y0 && r->rendered.back().y1 == y1 && r->rendered.back().x1 + 1 == x0) r->rendered.back().x1 = x1;
      else r->rendered.push_back({x0, y0, x1, y1});
      r->rendered_pixels += (long long)(x1 - x0 + 1) * (y1 - y0 + 1);
    }
    // batches: as many samples per pixel as fit, then as many pixels as fit; sample blocks in
    // increasing order so that every pixel's running sum adds its samples in the reference's order
    int s_block = (int)std::min<int64_t>(s_total, cap);
    int64_t px_block = std::max<int64_t>(1, cap / s_block);
    // keep both slots busy even when everything would fit one batch
    if (sharers == 1 && (int64_t)px.size() <= px_block && s_block == s_total && (int64_t)px.size() * s_total > (1 << 20))
      px_block = ((int64_t)px.size() + 1) / 2;
    for (int64_t p0 = 0; p0 < (int64_t)px.size(); p0 += px_block) {
      int np = (int)std::min<int64_t>(px_block, (int64_t)px.size() - p0);
      for (int s0 = 0; s0 < s_total; s0 += s_block) pending.push_back({&px, p0, np, s_off + s0, std::min(s_block, s_total - s0)});
    }
  };
  LaunchCfg lc;
  int rc = launch_cfg(ctx, lc);
  if (rc != IZPI_OK) return rc;
  int started = 0, resolved_upto = 0;  // batches [0, resolved_upto) have their resolve enqueued
  cudaEvent_t last_resolved = nullptr;
  for (;;) {
    bool any = false;
    for (BatchSlot& s : r->slot) {
      if (r->stats_on && &s != &r->slot[0]) continue;  // stage times are meaningful only without overlap
      if (!s.busy) {
        if (pending.empty() && !exhausted) claim();
        if (!pending.empty()) {
          const Batch b = pending.front();
          pending.pop_front();
          if ((rc = batch_start(ctx, r, s, started, b.px->data() + b.p0, b.np, b.s0, b.sc)) != IZPI_OK) return rc;
          started++;
        }
      }
      if (!s.busy) continue;
      any = true;
      if (!s.drained && (rc = batch_step(ctx, r, s, lc)) != IZPI_OK) return rc;
      if (s.drained && s.batch == resolved_upto) {  // resolves strictly in batch order
        if ((rc = batch_finish(ctx, r, s, last_resolved)) != IZPI_OK) return rc;
        last_resolved = s.resolved;
        resolved_upto++;
      }
    }
    if (!any && pending.empty() && exhausted) break;
  }
  if ((rc = sync_slots(r)) != IZPI_OK) return rc;
  if (r->stats_on) {
    for (const RenderState::Span& sp : r->spans) {
      float ms = 0;
      IZ_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
      if (sp.kind == 0) { r->stats.extend_ms += ms; r->stats.extend_launches++; }
      else if (sp.kind == 1) r->stats.shade_ms += ms;
      else r->stats.other_ms += ms;
    }
    r->spans.clear(); r->ev_used = 0;
  }
  return IZPI_OK;
}

int check_tiles(const RenderState* r, int32_t n_tiles, const uint32_t* tiles, const char* who) {
  for (int32_t t = 0; t < n_tiles; t++) {
    uint32_t x0 = tiles[4 * t], y0 = tiles[4 * t + 1], x1 = tiles[4 * t + 2], y1 = tiles[4 * t + 3];
    if (x1 < x0 || y1 < y0 || x1 >= (uint32_t)r->rp.width || y1 >= (uint32_t)r->rp.height) {
      set_error(std::string(who) + ": tile outside the image");
      return IZPI_EINVAL;
    }
  }
  return IZPI_OK;
}

int read_totals(izpi_ctx* ctx, unsigned long long v[3]) {
  RenderState* r = ctx->render;
  IZ_CUDA(cudaSetDevice(ctx->device));
  int rc = sync_slots(r);
  if (rc != IZPI_OK) return rc;
  IZ_CUDA(cudaMemcpy(v, r->d_total_rays, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return IZPI_OK;
}

}  // namespace

extern "C" {

int izpi_render_setup(izpi_ctx* ctx, const izpi_render_config* cfg) {
  if (!ctx || !cfg) { set_error("izpi_render_setup: bad argument"); return IZPI_EINVAL; }
  if (!ctx->has_scene) { set_error("izpi_render_setup: no scene uploaded"); return IZPI_ESTATE; }
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->width > 65535 || cfg->height > 65535 || cfg->spp <= 0 || cfg->max_depth < 0 ||
      cfg->sample_count < 0 || cfg->sample_offset < 0 || cfg->sample_offset + cfg->sample_count > cfg->spp ||
      cfg->sampler < IZPI_SAMPLER_COLOUR || cfg->sampler > IZPI_SAMPLER_NORMAL) {
    set_error("izpi_render_setup: invalid configuration");
    return IZPI_EINVAL;
  }
  if (ctx->scene.n_lights == 0 && cfg->sampler <= IZPI_SAMPLER_SPECTRAL) {  // HitableSlice.Random indexes an empty slice: the reference panics
    set_error("izpi_render_setup: scene has no emitters (scene.Lights is empty)");
    return IZPI_EINVAL;
  }
  if (!ctx->scene.attrs && ctx->scene.n_prims > 0) { set_error("izpi_render_setup: scene was uploaded without tri_attrs"); return IZPI_ESTATE; }
  return group_run(ctx, [&](izpi_ctx* m, int) { return setup_one(m, cfg); });
}

int izpi_render_tiles_shared(izpi_ctx* ctx, int32_t n_tiles, const uint32_t* tiles, uint64_t* cursor, int32_t sharers, double* canvas_rgba) {
  if (!ctx || n_tiles < 0 || (n_tiles > 0 && !tiles)) { set_error("izpi_render_tiles: bad argument"); return IZPI_EINVAL; }
  RenderState* r = ctx->render;
  if (!r) { set_error("izpi_render_tiles: izpi_render_setup has not been called"); return IZPI_ESTATE; }
  int rc = check_tiles(r, n_tiles, tiles, "izpi_render_tiles");
  if (rc != IZPI_OK) return rc;
  if (ctx->subs.empty()) {
    if ((rc = tiles_one(ctx, n_tiles, tiles, cursor, sharers)) != IZPI_OK) return rc;
  } else {
    // device group: one host thread per GPU, all claiming from one cursor (the caller's, when it shares the list further)
    const int members = 1 + (int)ctx->subs.size();
    ctx->group_cursor = 0;
    uint64_t* cur = cursor ? cursor : &ctx->group_cursor;
    const int all = cursor ? (sharers < 1 ? 1 : sharers) * members : members;
    if ((rc = group_run(ctx, [&](izpi_ctx* m, int) { return tiles_one(m, n_tiles, tiles, cur, all); })) != IZPI_OK) return rc;
  }
  if (canvas_rgba) {
    if (!ctx->subs.empty()) { set_error("izpi_render_tiles: a device group delivers the canvas through izpi_render_finish"); return IZPI_EINVAL; }
    return canvas_to_host(ctx, r, canvas_rgba, false);
  }
  return IZPI_OK;
}

int izpi_render_tiles(izpi_ctx* ctx, int32_t n_tiles, const uint32_t* tiles, double* canvas_rgba) {
  return izpi_render_tiles_shared(ctx, n_tiles, tiles, nullptr, 1, canvas_rgba);
}

int izpi_render_tile_rows(izpi_ctx* ctx, uint32_t strip_height, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, double* rows) {
  if (!ctx || !rows) { set_error("izpi_render_tile_rows: bad argument"); return IZPI_EINVAL; }
  RenderState* r = ctx->render;
  if (!r) { set_error("izpi_render_tile_rows: izpi_render_setup has not been called"); return IZPI_ESTATE; }
  if (strip_height == 0) { set_error("izpi_render_tile_rows: strip_height 0 (the reference indexes an empty slice here)"); return IZPI_EINVAL; }
  const uint32_t tile[4] = {x0, y0, x1, y1};
  int rc = check_tiles(r, 1, tile, "izpi_render_tile_rows");
  if (rc != IZPI_OK) return rc;
  if ((rc = tiles_one(ctx, 1, tile, nullptr, 1)) != IZPI_OK) return rc;  // a worker is one device: the group's first
  const int w = (int)(x1 - x0 + 1), h = (int)(y1 - y0 + 1);
  const long long stride = (long long)strip_height * 4 * w;
  double* d_rows = nullptr;
  IZ_CUDA(cudaMalloc(&d_rows, (size_t)stride * h * 8));
  cudaStream_t st = ctx->stream;
  cudaError_t e = cudaMemsetAsync(d_rows, 0, (size_t)stride * h * 8, st);
  if (e == cudaSuccess) {
    rc = launch(ctx, st, tile_rows_kernel, dim3((unsigned)(((long long)w * h + 127) / 128)), dim3(128), 0, r->rp, r->d_canvas, (int)x0, (int)y0, w, h,
                stride, d_rows);
    if (rc == IZPI_OK) e = cudaMemcpyAsync(rows, d_rows, (size_t)stride * h * 8, cudaMemcpyDeviceToHost, st);
    if (rc == IZPI_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(d_rows);
  if (rc != IZPI_OK) return rc;
  if (e != cudaSuccess) { set_error(std::string("izpi_render_tile_rows: ") + cudaGetErrorString(e)); return IZPI_ECUDA; }
  return IZPI_OK;
}

int izpi_render_canvas_device(izpi_ctx* ctx, double** d_canvas) {
  if (!ctx || !d_canvas) { set_error("izpi_render_canvas_device: bad argument"); return IZPI_EINVAL; }
  if (!ctx->render) { set_error("izpi_render_canvas_device: izpi_render_setup has not been called"); return IZPI_ESTATE; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  int rc = sync_slots(ctx->render);
  if (rc != IZPI_OK) return rc;
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  *d_canvas = ctx->render->d_canvas;
  return IZPI_OK;
}

int izpi_render_get_stats(izpi_ctx* ctx, izpi_render_stats* out) {
  if (!ctx || !out) { set_error("izpi_render_get_stats: bad argument"); return IZPI_EINVAL; }
  if (!ctx->render) { set_error("izpi_render_get_stats: izpi_render_setup has not been called"); return IZPI_ESTATE; }
  std::memset(out, 0, sizeof(*out));
  out->material_bins = (uint64_t)ctx->scene.n_bins;
  std::vector<izpi_ctx*> members{ctx};
  members.insert(members.end(), ctx->subs.begin(), ctx->subs.end());
  for (izpi_ctx* m : members) {
    unsigned long long v[3];
    int rc = read_totals(m, v);
    if (rc != IZPI_OK) return rc;
    const izpi_render_stats& s = m->render->stats;
    out->rays += v[0]; out->nodes_visited += v[1]; out->prim_tests += v[2];
    out->extend_launches += s.extend_launches;
    out->extend_ms = std::max(out->extend_ms, s.extend_ms); out->shade_ms = std::max(out->shade_ms, s.shade_ms);
    out->other_ms = std::max(out->other_ms, s.other_ms);  // members run concurrently
  }
  cudaSetDevice(ctx->device);
  return IZPI_OK;
}

int izpi_render_finish(izpi_ctx* ctx, double* canvas_rgba, uint64_t* total_rays) {
  if (!ctx) { set_error("izpi_render_finish: bad argument"); return IZPI_EINVAL; }
  RenderState* r = ctx->render;
  if (!r) { set_error("izpi_render_finish: izpi_render_setup has not been called"); return IZPI_ESTATE; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  if (total_rays) {
    izpi_render_stats st;
    int rc = izpi_render_get_stats(ctx, &st);
    if (rc != IZPI_OK) return rc;
    *total_rays = st.rays;
  }
  if (!canvas_rgba) return IZPI_OK;
  if (ctx->subs.empty()) return canvas_to_host(ctx, r, canvas_rgba, true);

  // ---- device group: pixels are disjoint, so every member moves only what it rendered ----
  const int W = r->rp.width, H = r->rp.height;
  const size_t pitch = (size_t)W * 32;
  if (r->rp.sampler == IZPI_SAMPLER_SPECTRAL) {
    // FireflyRejection reads 3x3 neighbourhoods across tile edges (firefly_rejection.go:40-73): the sums of the other
    // members' tiles go to the first device (peer copies of the rendered runs only), which runs the epilogue on the frame
    int rc = group_run(ctx, [&](izpi_ctx* m, int i) -> int {
      if (i == 0) return sync_slots(m->render);
      RenderState* mr = m->render;
      int rc2 = sync_slots(mr);
      if (rc2 != IZPI_OK) return rc2;
      for (const TileRun& t : mr->rendered) {
        int row0, rows;
        if (!run_rows(mr, t, row0, rows)) continue;
        const size_t off = ((size_t)row0 * W + t.x0) * 4;
        IZ_CUDA(cudaMemcpy2DAsync(r->d_canvas + off, pitch, mr->d_canvas + off, pitch, (size_t)(t.x1 - t.x0 + 1) * 32, (size_t)rows,
                                  cudaMemcpyDefault, m->stream));
      }
      IZ_CUDA(cudaStreamSynchronize(m->stream));
      return IZPI_OK;
    });
    if (rc != IZPI_OK) return rc;
    return canvas_to_host(ctx, r, canvas_rgba, true);
  }
  // RGB / AOV samplers: each member takes the mean of its own canvas and writes its runs straight into the caller's
  // canvas -- N device-to-host streams in parallel, nothing crosses the first device.  Pixels nobody rendered are zero,
  // as in the reference's fresh Float64NRGBA (row 0 always is: rgb.go:41 never writes it).
  long long covered = 0;
  for (const izpi_ctx* m : ctx->subs) covered += m->render->rendered_pixels;
  covered += r->rendered_pixels;
  const bool full = covered == (long long)W * H;  // the tile grid of the whole image: only canvas row 0 is left over
  const int members = 1 + (int)ctx->subs.size();
  auto copy_runs = [&](izpi_ctx* m) -> int {
    RenderState* mr = m->render;
    for (const TileRun& t : mr->rendered) {
      int row0, rows;
      if (!run_rows(mr, t, row0, rows)) continue;
      const size_t off = ((size_t)row0 * W + t.x0) * 4;
      IZ_CUDA(cudaMemcpy2DAsync(canvas_rgba + off, pitch, mr->d_out + off, pitch, (size_t)(t.x1 - t.x0 + 1) * 32, (size_t)rows,
                                cudaMemcpyDeviceToHost, m->stream));
    }
    IZ_CUDA(cudaStreamSynchronize(m->stream));
    return IZPI_OK;
  };
  int rc = group_run(ctx, [&](izpi_ctx* m, int i) -> int {
    int rc2 = canvas_epilogue(m, m->render, false);
    if (rc2 != IZPI_OK) return rc2;
    IZ_CUDA(cudaStreamSynchronize(m->stream));
    if (full) {
      if (i == 0) std::memset(canvas_rgba, 0, pitch);
      return copy_runs(m);
    }
    // partial frame: every member clears a share of the caller's canvas; the runs follow once all shares are clear
    const size_t total = (size_t)H * pitch, b0 = total / members * i, b1 = i == members - 1 ? total : total / members * (i + 1);
    std::memset(reinterpret_cast<char*>(canvas_rgba) + b0, 0, b1 - b0);
    return IZPI_OK;
  });
  if (rc == IZPI_OK && !full) rc = group_run(ctx, [&](izpi_ctx* m, int) -> int { return copy_runs(m); });
  return rc;
}

}  // extern "C"
