// trace.cu -- batched closest hit: the hitable.Hitable.Hit seam (internal/hitable/api.go:15) for a
// whole ray batch.  Persistent warps pull 32-ray packets from a global work queue; each lane walks
// the BVH4 in the reference's order with its own short stack in shared memory.
#include <cstdlib>
#include <cstring>

#include "intersect_g2.cuh"

using namespace izpi;

namespace {

constexpr int kTraceThreads = 128;
constexpr int kStackDepth = 64;  // bvh4.go:71

template <bool COUNT>
__global__ void __launch_bounds__(kTraceThreads)
trace_kernel(const __grid_constant__ DScene sc, long long n, const double* __restrict__ org,
             const double* __restrict__ dir, double tmin, double tmax, int32_t* __restrict__ ids,
             double* __restrict__ ts, unsigned long long* counters) {
  extern __shared__ int32_t stack_smem[];  // [kStackDepth][blockDim.x]: lane-interleaved, conflict-free
  int32_t* stack = stack_smem + threadIdx.x;
  const int stride = blockDim.x;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t n_nodes = 0, n_prims = 0;
  for (;;) {
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(&counters[0], 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    long long i = base + lane;
    if (i < n) {
      DRay r;
      r.o = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]);
      r.d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
      r.time = 0; r.lambda = 0;
      double t = 0;
      int rec = world_closest<COUNT>(sc, r, tmin, tmax, t, stack, stride, n_nodes, n_prims);
      ids[i] = rec >= 0 ? sc.prims[rec].orig_id : -1;
      ts[i] = rec >= 0 ? t : 0.0;
    }
  }
  if (COUNT) {
    atomicAdd(&counters[1], (unsigned long long)n_nodes);
    atomicAdd(&counters[2], (unsigned long long)n_prims);
  }
}

// ---- 4 lanes per ray -------------------------------------------------------------------------
// Each warp owns 8 ray slots.  Rays are drawn from the global queue in chunks of kChunk per warp (one
// atomicAdd per chunk), handed to idle groups as they finish (persistent threads with ray replacement),
// and the warp alternates between a node phase and a leaf phase (intersect_g4.cuh).
constexpr int kChunk = 256;
#ifndef IZPI_G4_MIN_BLOCKS
#define IZPI_G4_MIN_BLOCKS 7
#endif

template <bool COUNT, bool F32>
__global__ void __launch_bounds__(kTraceThreads, IZPI_G4_MIN_BLOCKS)
trace_g4_kernel(const __grid_constant__ DScene sc, long long n, const double* __restrict__ org,
                const double* __restrict__ dir, double tmin, double tmax, int32_t* __restrict__ ids,
                double* __restrict__ ts, unsigned long long* counters, int stragglers) {
  extern __shared__ int2 g4_stack_smem[];  // [rays per block][kG4Stack + 1]: one private stack per ray
  const unsigned lane = threadIdx.x & 31u;
  const int g = lane >> 2, j = lane & 3;
  const int gshift = g * 4;
  int2* stack = g4_stack_smem + (size_t)(threadIdx.x >> 2) * kG4Slab;
  // all ray indices are 32-bit here (the host entry points cut larger batches into launches below 2^31 rays)
  const int n32 = (int)n;
  const int warps = (int)gridDim.x * (kTraceThreads / 32);
  int chunk = (n32 / (warps * 4) + 7) & ~7;  // rays per atomicAdd: shrinks for small batches (tail balance)
  chunk = chunk < 8 ? 8 : (chunk > kChunk ? kChunk : chunk);
  uint32_t n_nodes = 0, n_prims = 0;
  // Cold state lives in shared memory so that it holds no registers across the traversal phases (the kernel is bound by
  // resident warps): the ray index of a group in the pad slot of its slab, the warp's position in its ray chunk in
  // the pad slot of the warp's first slab (.y of the same int2; only group 0's is used).
  int* ray_slot = reinterpret_cast<int*>(stack + kG4Stack + 9);                                          // .x = ray index (-1: none)
  int2* chunk_slot = g4_stack_smem + (size_t)((threadIdx.x & ~31u) >> 2) * kG4Slab + kG4Stack + 9;        // warp-wide
  int* chunk_pos = reinterpret_cast<int*>(chunk_slot) + 1;                                               // .y of group 0's pad: next ray
  int* chunk_lim = reinterpret_cast<int*>(chunk_slot + kG4Slab) + 1;                                     // .y of group 1's pad: chunk end (-1: exhausted)
  if (j == 0) *ray_slot = -1;
  if (lane == 0) { *chunk_pos = 0; *chunk_lim = 0; }
  __syncwarp();
  G4State s;
  s.cur = kIdle;
  for (;;) {
    // ---- hand rays to idle groups
    unsigned idle = __ballot_sync(0xffffffffu, s.cur == kIdle);
    if (idle) {
      int chunk_next = *chunk_pos, chunk_end = *chunk_lim;
      bool exhausted = chunk_end < 0;
      if (exhausted) chunk_end = chunk_next = 0;
      __syncwarp();
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&counters[0], (unsigned long long)chunk);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned long long)n32) { exhausted = true; chunk_next = chunk_end = 0; }
        else { chunk_next = (int)b; chunk_end = (int)b + chunk < n32 ? (int)b + chunk : n32; }
      }
      int before = __popc(idle & ((1u << gshift) - 1u)) >> 2;  // idle groups ahead of mine
      int total = __popc(idle) >> 2;
      if (s.cur == kIdle && chunk_next + before < chunk_end) {
        const int ray = chunk_next + before;
        const double* po = org + 3 * (size_t)ray;
        const double* pd = dir + 3 * (size_t)ray;
        DRay r;
        r.o = mk(po[0], po[1], po[2]);
        r.d = mk(pd[0], pd[1], pd[2]);
        r.time = 0; r.lambda = 0;
        g4_begin(s, sc, r, tmax, stack, j);
        if (j == 0) {
          if (s.cur == kIdle) { ids[ray] = -1; ts[ray] = 0.0; }
          else *ray_slot = ray;
        }
      }
      int take = chunk_end - chunk_next;
      chunk_next += take < total ? take : total;
      if (lane == 0) { *chunk_pos = chunk_next; *chunk_lim = exhausted ? -1 : chunk_end; }
      if (exhausted && __ballot_sync(0xffffffffu, s.cur == kIdle) == 0xffffffffu) break;
    }
    // ---- node phase, then leaf phase (both warp-uniform)
    g4_node_phase<COUNT>(s, sc, stack, lane, gshift, j, n_nodes, stragglers);
    g4_leaf_phase<COUNT, F32>(s, sc, stack, lane, gshift, j, n_nodes, n_prims, tmin);
    // ---- finished rays write their answer
    if (s.cur == kIdle && j == 0) {
      const int ray = *ray_slot;
      if (ray >= 0) {
        ids[ray] = s.best >= 0 ? sc.prims[s.best].orig_id : -1;
        ts[ray] = s.best >= 0 ? s.tmax : 0.0;
        *ray_slot = -1;
      }
    }
  }
  if (COUNT && j == 0) {
    atomicAdd(&counters[1], (unsigned long long)n_nodes);
    atomicAdd(&counters[2], (unsigned long long)n_prims);
  }
  if (COUNT && j != 0) atomicAdd(&counters[2], (unsigned long long)n_prims);
}


// ---- 2 lanes per ray --------------------------------------------------------------------------
// Sixteen ray slots per warp (intersect_g2.cuh); same queue, same replacement scheme as the 4-lane kernel.
// Blocks of two warps: 14 of them fit an SM at 72 registers without spills (28 warps; 13-14 blocks: 740 Mrays/s), where
// blocks of four warps stop at 6 x 4 = 24 warps with 78 registers (681) or spill at 7 x 4 (668); 16 x 2 warps at 64
// registers spill and fall to 616.
#ifndef IZPI_G2_THREADS
#define IZPI_G2_THREADS 64
#endif
#ifndef IZPI_G2_MIN_BLOCKS
#define IZPI_G2_MIN_BLOCKS 14
#endif
constexpr int kG2Threads = IZPI_G2_THREADS;

template <bool COUNT, bool F32, int STACK>
__global__ void __launch_bounds__(kG2Threads, IZPI_G2_MIN_BLOCKS)
trace_g2_kernel(const __grid_constant__ DScene sc, long long n, const double* __restrict__ org,
                const double* __restrict__ dir, double tmin, double tmax, int32_t* __restrict__ ids,
                double* __restrict__ ts, unsigned long long* counters, int stragglers) {
  extern __shared__ int2 g4_stack_smem[];  // [rays per block][STACK + 10]: one private slab per ray
  constexpr int kSlots = G2Slab<STACK>::kSlots;
  const unsigned lane = threadIdx.x & 31u;
  const int j = lane & 1, pshift = (int)(lane & ~1u);
  int2* stack = g4_stack_smem + (size_t)(threadIdx.x >> 1) * kSlots;
  const int n32 = (int)n;
  const int warps = (int)gridDim.x * (kG2Threads / 32);
  // Guided self-scheduling: a warp asks for (rays still in the queue) / (2 x warps), at most kChunk and at least one
  // full warp of pairs, so chunks shrink as the queue drains and the kernel's tail -- the time the last warps need for what
  // they still hold -- stays short whatever the batch size (it matters for the sliced host-buffer path and for the
  // renderer's many small launches).  The queue head is read without ordering: a stale value only changes a chunk's size.
  uint32_t n_nodes = 0, n_prims = 0;
  // cold state in shared memory (see trace_g4_kernel): ray index in the pad slot of the pair's slab, the warp's chunk cursor
  // in the pad slots of its first two slabs
  int* ray_slot = reinterpret_cast<int*>(stack + STACK + 9);
  int2* chunk_slot = g4_stack_smem + (size_t)((threadIdx.x & ~31u) >> 1) * kSlots + STACK + 9;
  int* chunk_pos = reinterpret_cast<int*>(chunk_slot) + 1;
  int* chunk_lim = reinterpret_cast<int*>(chunk_slot + kSlots) + 1;
  if (j == 0) *ray_slot = -1;
  if (lane == 0) { *chunk_pos = 0; *chunk_lim = 0; }
  __syncwarp();
  G4State s;
  s.cur = kIdle;
  for (;;) {
    unsigned idle = __ballot_sync(0xffffffffu, s.cur == kIdle);
    if (idle) {
      int chunk_next = *chunk_pos, chunk_end = *chunk_lim;
      bool exhausted = chunk_end < 0;
      if (exhausted) chunk_end = chunk_next = 0;
      __syncwarp();
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long b = 0;
        int chunk = 16;
        if (lane == 0) {
          const long long left = (long long)n32 - (long long)*reinterpret_cast<volatile unsigned long long*>(&counters[0]);
          long long want = left > 0 ? left / (2 * (long long)warps) : 0;
          chunk = want > kChunk ? kChunk : (want < 16 ? 16 : (int)((want + 15) & ~15ll));
          b = atomicAdd(&counters[0], (unsigned long long)chunk);
        }
        b = __shfl_sync(0xffffffffu, b, 0);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (b >= (unsigned long long)n32) { exhausted = true; chunk_next = chunk_end = 0; }
        else { chunk_next = (int)b; chunk_end = (int)b + chunk < n32 ? (int)b + chunk : n32; }
      }
      int before = __popc(idle & ((1u << pshift) - 1u)) >> 1;  // idle pairs ahead of mine
      int total = __popc(idle) >> 1;
      if (s.cur == kIdle && chunk_next + before < chunk_end) {
        const int ray = chunk_next + before;
        const double* po = org + 3 * (size_t)ray;
        const double* pd = dir + 3 * (size_t)ray;
        DRay r;
        r.o = mk(po[0], po[1], po[2]);
        r.d = mk(pd[0], pd[1], pd[2]);
        r.time = 0; r.lambda = 0;
        g2_begin<STACK>(s, sc, r, tmax, stack, j);
        if (j == 0) {
          if (s.cur == kIdle) { ids[ray] = -1; ts[ray] = 0.0; }
          else *ray_slot = ray;
        }
      }
      int take = chunk_end - chunk_next;
      chunk_next += take < total ? take : total;
      if (lane == 0) { *chunk_pos = chunk_next; *chunk_lim = exhausted ? -1 : chunk_end; }
      if (exhausted && __ballot_sync(0xffffffffu, s.cur == kIdle) == 0xffffffffu) break;
    }
    g2_node_phase<COUNT, STACK>(s, sc, stack, pshift, j, n_nodes, stragglers);
    g2_leaf_phase<COUNT, F32, STACK>(s, sc, stack, pshift, j, n_nodes, n_prims, tmin);
    if (s.cur == kIdle && j == 0) {
      const int ray = *ray_slot;
      if (ray >= 0) {
        ids[ray] = s.best >= 0 ? sc.prims[s.best].orig_id : -1;
        ts[ray] = s.best >= 0 ? s.tmax : 0.0;
        *ray_slot = -1;
      }
    }
  }
  if (COUNT && j == 0) atomicAdd(&counters[1], (unsigned long long)n_nodes);
  if (COUNT) atomicAdd(&counters[2], (unsigned long long)n_prims);
}

template <int STACK>
int launch_g2(izpi_ctx* ctx, int64_t n, const double* d_org, const double* d_dir, double tmin, double tmax, int mode, int32_t* d_ids,
              double* d_t, cudaStream_t st, bool count, unsigned long long* counters) {
  const size_t smem = (size_t)(kG2Threads / 2) * G2Slab<STACK>::kSlots * sizeof(int2);
  auto k = mode == IZPI_TRACE_FP32 ? trace_g2_kernel<false, true, STACK> : (count ? trace_g2_kernel<true, false, STACK> : trace_g2_kernel<false, false, STACK>);
  int& bps = ctx->occ.trace_g2[STACK == kG2Stack ? 0 : 1];  // per context: function attributes and occupancy belong to a device
  if (!bps) {
    IZ_CUDA(cudaFuncSetAttribute(trace_g2_kernel<false, false, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IZ_CUDA(cudaFuncSetAttribute(trace_g2_kernel<true, false, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IZ_CUDA(cudaFuncSetAttribute(trace_g2_kernel<false, true, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, trace_g2_kernel<false, false, STACK>, kG2Threads, smem));
    if (bps < 1) bps = 1;
  }
  int use_bps = bps;
  if (const char* e = getenv("IZPI_TRACE_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < use_bps) use_bps = v; }
  long long want = (n + (kG2Threads / 2) - 1) / (kG2Threads / 2);
  long long grid = (long long)ctx->sm_count * use_bps;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  k<<<(unsigned)grid, kG2Threads, smem, st>>>(ctx->scene, (long long)n, d_org, d_dir, tmin, tmax, d_ids, d_t, counters, ctx->pair_stragglers);
  IZ_CUDA(cudaGetLastError());
  ctx->launches++;
  return IZPI_OK;
}

// Diagnostic: the 4-wide slab test alone, so the reference's golden masks can be replayed on the device.
__global__ void box4_kernel(int n, const float* __restrict__ org, const float* __restrict__ inv,
                            const float* __restrict__ bounds, const float* __restrict__ tmax, uint8_t* __restrict__ masks) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* o = org + 3 * i; const float* v = inv + 3 * i; const float* b = bounds + 24 * i;
  uint8_t m = 0;
  for (int k = 0; k < 4; k++)
    if (box1(o[0], o[1], o[2], v[0], v[1], v[2], b[k], b[4 + k], b[8 + k], b[12 + k], b[16 + k], b[20 + k], tmax[i])) m |= (uint8_t)(1u << k);
  masks[i] = m;
}


// Diagnostic: world.Hit with the FULL hit record (hitrecord.HitRecord: t, u, v, p, normal), one thread per ray, so that the
// reference's Triangle.Hit records (triangle_test.go:69-134) can be replayed on the device field by field.
__global__ void debug_hit_kernel(const __grid_constant__ DScene sc, int n, const double* __restrict__ org, const double* __restrict__ dir,
                                 double tmin, double tmax, int32_t* __restrict__ ids, double* __restrict__ out9) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DRay r;
  r.o = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]);
  r.d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
  r.time = 0; r.lambda = 0;
  int32_t stack[kStackDepth];
  uint32_t a = 0, b = 0;
  double t = 0;
  int rec = world_closest<false>(sc, r, tmin, tmax, t, stack, 1, a, b);
  double* o = out9 + 9 * (size_t)i;
  for (int k = 0; k < 9; k++) o[k] = 0.0;
  ids[i] = -1;
  if (rec < 0) return;
  PrimRec pr = load_rec(sc.prims + rec);
  DHit h;
  if (!prim_hit<true>(sc, rec, pr, r, tmin, tmax, h)) return;
  ids[i] = pr.orig_id;
  o[0] = h.t; o[1] = h.u; o[2] = h.v; o[3] = h.p.x; o[4] = h.p.y; o[5] = h.p.z; o[6] = h.n.x; o[7] = h.n.y; o[8] = h.n.z;
}

// Diagnostic: dependent-FMA throughput of the fp32 / fp64 vector pipes (the FLOP side of the traversal
// roofline; MEASURED_PEAKS.json carries HBM and tensor-core peaks only).  16 independent chains per thread.
template <typename T>
__global__ void fma_peak_kernel(int iters, T seed, T* out) {
  T a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed + (T)(threadIdx.x + k);
  const T m = (T)1.0000001, c = (T)1e-7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = fma(a[k], m, c);
  }
  T acc = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) acc += a[k];
  if (acc == (T)12345.678) out[0] = acc;  // keeps the chains alive
}

int launch_trace(izpi_ctx* ctx, int64_t n, const double* d_org, const double* d_dir, double tmin, double tmax,
                 int mode, int32_t* d_ids, double* d_t, cudaStream_t st, bool count, unsigned long long* counters) {
  if (mode != IZPI_TRACE_EXACT && mode != IZPI_TRACE_FP32) { set_error("izpi_trace_closest: unknown mode"); return IZPI_EINVAL; }
  if (mode == IZPI_TRACE_FP32 && !(ctx->scene.world_kind == IZPI_WORLD_BVH4 && ctx->scene.g4_ok)) {
    set_error("izpi_trace_closest: IZPI_TRACE_FP32 needs a BVH4 world built by NewBVH4");
    return IZPI_EINVAL;
  }
  if (n > (1ll << 30)) { set_error("izpi_trace_closest: at most 2^30 rays per launch; split the batch"); return IZPI_EINVAL; }
  IZ_CUDA(cudaMemsetAsync(counters, 0, 3 * sizeof(unsigned long long), st));
  // Two lanes per ray (16 rays per warp) when the tree's worst-case stack fits the kG2Stack-entry slab (six resident
  // blocks x 32 rays x 464 B of shared memory, or 13 x 32 x 496 B for the 52-entry variant); with the reference's full 64 entries only five blocks would fit, and the
  // 4-lane kernel (8 rays per warp, 7 blocks) is faster than that.  NewBVH4 trees are balanced -- three entries per level
  // of inner nodes: 27 for the 1 M-triangle mesh, 33 for 11.5 M -- so they fit; a tree that does not is still traced
  // exactly, by the 4-lane kernel.
  if (ctx->scene.world_kind == IZPI_WORLD_BVH4 && ctx->scene.g4_ok && (!ctx->force_scalar || mode == IZPI_TRACE_FP32) &&
      ctx->trace_lanes == 2 && ctx->scene.g4_need <= kG2StackDeep) {
    if (ctx->scene.g4_need <= kG2Stack) return launch_g2<kG2Stack>(ctx, n, d_org, d_dir, tmin, tmax, mode, d_ids, d_t, st, count, counters);
    return launch_g2<kG2StackDeep>(ctx, n, d_org, d_dir, tmin, tmax, mode, d_ids, d_t, st, count, counters);
  }
  if (ctx->scene.world_kind == IZPI_WORLD_BVH4 && ctx->scene.g4_ok && (!ctx->force_scalar || mode == IZPI_TRACE_FP32)) {
    size_t smem4 = (size_t)(kTraceThreads / 4) * kG4Slab * sizeof(int2);
    auto k4 = mode == IZPI_TRACE_FP32 ? trace_g4_kernel<false, true> : (count ? trace_g4_kernel<true, false> : trace_g4_kernel<false, false>);
    int& bps4 = ctx->occ.trace_g4;
    if (!bps4) {
      IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps4, trace_g4_kernel<false, false>, kTraceThreads, smem4));
      if (bps4 < 1) bps4 = 1;
    }
    long long want4 = (n + (kTraceThreads / 4) - 1) / (kTraceThreads / 4);
    int use_bps = bps4;
    if (const char* e = getenv("IZPI_TRACE_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < use_bps) use_bps = v; }  // occupancy experiments
    long long grid4 = (long long)ctx->sm_count * use_bps;
    if (grid4 > want4) grid4 = want4;
    if (grid4 < 1) grid4 = 1;
    k4<<<(unsigned)grid4, kTraceThreads, smem4, st>>>(ctx->scene, (long long)n, d_org, d_dir, tmin, tmax, d_ids, d_t, counters,
                                                        ctx->node_stragglers);
    IZ_CUDA(cudaGetLastError());
    ctx->launches++;
    return IZPI_OK;
  }
  size_t smem = (size_t)kStackDepth * kTraceThreads * sizeof(int32_t);
  auto kern = count ? trace_kernel<true> : trace_kernel<false>;
  int& blocks_per_sm = ctx->occ.trace_scalar;
  if (!blocks_per_sm) {
    IZ_CUDA(cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IZ_CUDA(cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, trace_kernel<false>, kTraceThreads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  long long want = (n + kTraceThreads - 1) / kTraceThreads;
  long long grid = (long long)ctx->sm_count * blocks_per_sm;  // persistent: a whole number of waves
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kTraceThreads, smem, st>>>(ctx->scene, (long long)n, d_org, d_dir, tmin, tmax, d_ids, d_t,
                                                     counters);
  IZ_CUDA(cudaGetLastError());
  ctx->launches++;
  return IZPI_OK;
}

}  // namespace

extern "C" {

int izpi_trace_closest_device(izpi_ctx* ctx, int64_t n, const double* d_org, const double* d_dir, double tmin, double tmax,
                              int mode, int32_t* d_ids, double* d_t, void* stream) {
  if (!ctx || n < 0) { set_error("izpi_trace_closest_device: bad argument"); return IZPI_EINVAL; }
  if (!ctx->has_scene) { set_error("izpi_trace_closest_device: no scene uploaded"); return IZPI_ESTATE; }
  if (n == 0) return IZPI_OK;
  IZ_CUDA(cudaSetDevice(ctx->device));
  return launch_trace(ctx, n, d_org, d_dir, tmin, tmax, mode, d_ids, d_t, stream ? (cudaStream_t)stream : ctx->stream, false,
                      ctx->d_counters);
}

static int trace_closest_one(izpi_ctx* ctx, int64_t n, const double* org, const double* dir, double tmin, double tmax, int mode,
                             int32_t* prim_id, double* t, izpi_trace_stats* stats);

int izpi_trace_closest(izpi_ctx* ctx, int64_t n, const double* org, const double* dir, double tmin, double tmax, int mode,
                       int32_t* prim_id, double* t, izpi_trace_stats* stats) {
  if (!ctx || n < 0 || (n > 0 && (!org || !dir || !prim_id || !t))) { set_error("izpi_trace_closest: bad argument"); return IZPI_EINVAL; }
  if (!ctx->has_scene) { set_error("izpi_trace_closest: no scene uploaded"); return IZPI_ESTATE; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return IZPI_OK;
  const int members = 1 + (int)ctx->subs.size();
  if (members == 1 || n < 2 * members) return trace_closest_one(ctx, n, org, dir, tmin, tmax, mode, prim_id, t, stats);
  // Device group: rays are independent, so the batch is cut into contiguous slices, one per member; every member runs the
  // single-device pipeline on its slice and the answers land in disjoint ranges of the caller's buffers (no exchange step).
  std::vector<izpi_trace_stats> st((size_t)members);
  int rc = group_run(ctx, [&](izpi_ctx* m, int i) -> int {
    const int64_t b = n * i / members, e = n * (i + 1) / members;
    return trace_closest_one(m, e - b, org + 3 * b, dir + 3 * b, tmin, tmax, mode, prim_id + b, t + b, stats ? &st[(size_t)i] : nullptr);
  });
  if (rc == IZPI_OK && stats) {
    for (const izpi_trace_stats& x : st) {
      stats->rays += x.rays; stats->nodes_visited += x.nodes_visited; stats->prim_tests += x.prim_tests;
      if (x.kernel_ms > stats->kernel_ms) stats->kernel_ms = x.kernel_ms;  // members run concurrently
    }
  }
  return rc;
}

static int trace_closest_one(izpi_ctx* ctx, int64_t n, const double* org, const double* dir, double tmin, double tmax, int mode,
                             int32_t* prim_id, double* t, izpi_trace_stats* stats) {
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return IZPI_OK;
  IZ_CUDA(cudaSetDevice(ctx->device));
  if (n > ctx->ray_capacity) {
    cudaFree(ctx->d_org); cudaFree(ctx->d_dir); cudaFree(ctx->d_ids); cudaFree(ctx->d_t);
    ctx->d_org = ctx->d_dir = ctx->d_t = nullptr; ctx->d_ids = nullptr; ctx->ray_capacity = 0;
    IZ_CUDA(cudaMalloc(&ctx->d_org, (size_t)n * 24));
    IZ_CUDA(cudaMalloc(&ctx->d_dir, (size_t)n * 24));
    IZ_CUDA(cudaMalloc(&ctx->d_ids, (size_t)n * 4));
    IZ_CUDA(cudaMalloc(&ctx->d_t, (size_t)n * 8));
    ctx->ray_capacity = n;
  }
  if (stats) {  // counting run: one launch on one stream
    cudaStream_t st = ctx->stream;
    IZ_CUDA(cudaMemcpyAsync(ctx->d_org, org, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    IZ_CUDA(cudaMemcpyAsync(ctx->d_dir, dir, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    IZ_CUDA(cudaEventRecord(ctx->ev0, st));
    int rc = launch_trace(ctx, n, ctx->d_org, ctx->d_dir, tmin, tmax, mode, ctx->d_ids, ctx->d_t, st, true, ctx->d_counters);
    if (rc != IZPI_OK) return rc;
    IZ_CUDA(cudaEventRecord(ctx->ev1, st));
    IZ_CUDA(cudaMemcpyAsync(prim_id, ctx->d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    IZ_CUDA(cudaMemcpyAsync(t, ctx->d_t, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    IZ_CUDA(cudaStreamSynchronize(st));
    unsigned long long c[3];
    IZ_CUDA(cudaMemcpy(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    stats->rays = (uint64_t)n; stats->nodes_visited = c[1]; stats->prim_tests = c[2];
    float ms = 0;
    IZ_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    stats->kernel_ms = ms;
    return IZPI_OK;
  }
  // The batch is cut into slices that alternate between two streams: while slice k is traversed, the rays of
  // slice k+1 arrive over PCIe and the answers of slice k-1 leave (the copies are truly asynchronous when the
  // caller's buffers are pinned; pageable buffers still work, just without the overlap).
  // Slices grow from 2^18 to 2^21 rays: the first upload (nothing to overlap it with) stays short, later slices are large
  // enough to keep the persistent kernel's tail small.
  cudaStream_t ss[2] = {ctx->stream, ctx->stream2};
  int k = 0;
  int64_t slice = 1 << 18, slice_max = 1 << 21;
  if (const char* e = getenv("IZPI_TRACE_SLICE_LOG2")) { int v = atoi(e); if (v >= 18 && v <= 26) slice_max = (int64_t)1 << v; }
  for (int64_t b = 0; b < n; k ^= 1) {
    int64_t m = n - b < slice ? n - b : slice;
    cudaStream_t st = ss[k];
    IZ_CUDA(cudaMemcpyAsync(ctx->d_org + 3 * b, org + 3 * b, (size_t)m * 24, cudaMemcpyHostToDevice, st));
    IZ_CUDA(cudaMemcpyAsync(ctx->d_dir + 3 * b, dir + 3 * b, (size_t)m * 24, cudaMemcpyHostToDevice, st));
    int rc = launch_trace(ctx, m, ctx->d_org + 3 * b, ctx->d_dir + 3 * b, tmin, tmax, mode, ctx->d_ids + b, ctx->d_t + b, st, false,
                          ctx->d_counters + 4 * k);
    if (rc != IZPI_OK) return rc;
    IZ_CUDA(cudaMemcpyAsync(prim_id + b, ctx->d_ids + b, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    IZ_CUDA(cudaMemcpyAsync(t + b, ctx->d_t + b, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    b += m;
    if (slice < slice_max) slice <<= 1;
  }
  IZ_CUDA(cudaStreamSynchronize(ss[0]));
  IZ_CUDA(cudaStreamSynchronize(ss[1]));
  return IZPI_OK;
}

uint64_t izpi_launch_count(const izpi_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t v = ctx->launches;
  for (const izpi_ctx* s : ctx->subs) v += s->launches;
  return v;
}

int izpi_debug_ray_aabb4(izpi_ctx* ctx, int32_t n, const float* org, const float* inv, const float* bounds, const float* tmax,
                         uint8_t* masks) {
  if (!ctx || n <= 0 || !org || !inv || !bounds || !tmax || !masks) { set_error("izpi_debug_ray_aabb4: bad argument"); return IZPI_EINVAL; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  float *d_o, *d_i, *d_b, *d_t; uint8_t* d_m;
  IZ_CUDA(cudaMalloc(&d_o, (size_t)n * 12)); IZ_CUDA(cudaMalloc(&d_i, (size_t)n * 12));
  IZ_CUDA(cudaMalloc(&d_b, (size_t)n * 96)); IZ_CUDA(cudaMalloc(&d_t, (size_t)n * 4)); IZ_CUDA(cudaMalloc(&d_m, (size_t)n));
  IZ_CUDA(cudaMemcpy(d_o, org, (size_t)n * 12, cudaMemcpyHostToDevice));
  IZ_CUDA(cudaMemcpy(d_i, inv, (size_t)n * 12, cudaMemcpyHostToDevice));
  IZ_CUDA(cudaMemcpy(d_b, bounds, (size_t)n * 96, cudaMemcpyHostToDevice));
  IZ_CUDA(cudaMemcpy(d_t, tmax, (size_t)n * 4, cudaMemcpyHostToDevice));
  box4_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, d_o, d_i, d_b, d_t, d_m);
  IZ_CUDA(cudaGetLastError());
  ctx->launches++;
  IZ_CUDA(cudaStreamSynchronize(ctx->stream));
  IZ_CUDA(cudaMemcpy(masks, d_m, (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_o); cudaFree(d_i); cudaFree(d_b); cudaFree(d_t); cudaFree(d_m);
  return IZPI_OK;
}


int izpi_debug_hit(izpi_ctx* ctx, int32_t n, const double* org, const double* dir, double tmin, double tmax, int32_t* prim_id, double* out9) {
  if (!ctx || n <= 0 || !org || !dir || !prim_id || !out9) { set_error("izpi_debug_hit: bad argument"); return IZPI_EINVAL; }
  if (!ctx->has_scene) { set_error("izpi_debug_hit: no scene uploaded"); return IZPI_ESTATE; }
  if (!ctx->scene.attrs && ctx->scene.n_prims > 0) { set_error("izpi_debug_hit: scene was uploaded without tri_attrs"); return IZPI_ESTATE; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  double *d_o = nullptr, *d_d = nullptr, *d_out = nullptr;
  int32_t* d_ids = nullptr;
  cudaError_t e = cudaMalloc(&d_o, (size_t)n * 24);
  if (e == cudaSuccess) e = cudaMalloc(&d_d, (size_t)n * 24);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * 72);
  if (e == cudaSuccess) e = cudaMalloc(&d_ids, (size_t)n * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_o, org, (size_t)n * 24, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_d, dir, (size_t)n * 24, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    debug_hit_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(ctx->scene, n, d_o, d_d, tmin, tmax, d_ids, d_out);
    e = cudaGetLastError();
    ctx->launches++;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(prim_id, d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out9, d_out, (size_t)n * 72, cudaMemcpyDeviceToHost);
  cudaFree(d_o); cudaFree(d_d); cudaFree(d_out); cudaFree(d_ids);
  if (e != cudaSuccess) { set_error(std::string("izpi_debug_hit: ") + cudaGetErrorString(e)); return IZPI_ECUDA; }
  return IZPI_OK;
}

int izpi_debug_fma_peak(izpi_ctx* ctx, int fp64, double* tflops) {
  if (!ctx || !tflops) { set_error("izpi_debug_fma_peak: bad argument"); return IZPI_EINVAL; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  void* d_out = nullptr;
  IZ_CUDA(cudaMalloc(&d_out, 16));
  const int iters = 1 << 14, threads = 256, blocks = ctx->sm_count * 8;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    IZ_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (fp64) fma_peak_kernel<double><<<blocks, threads, 0, ctx->stream>>>(iters, 1.0, static_cast<double*>(d_out));
    else fma_peak_kernel<float><<<blocks, threads, 0, ctx->stream>>>(iters, 1.0f, static_cast<float*>(d_out));
    IZ_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    IZ_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    IZ_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->launches++;
    double tf = 2.0 * 16.0 * (double)iters * threads * (double)blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaFree(d_out);
  *tflops = best;
  return IZPI_OK;
}

}  // extern "C"
