// bvh_build.cu -- device-side construction of a BVH4 in the reference's node format.
//
// The reference builds its tree on one host thread with a random-axis median split (hitable/bvh4.go:558-855).
// This builder is the optional replacement for scenes where setup time matters: a linear BVH (63-bit Morton
// codes of the box centres, radix sort, Karras' parallel binary radix tree, bottom-up box fitting) collapsed to
// 4-wide nodes.  It emits exactly the reference's layout -- 128-byte BVH4Node, leaves as own nodes using slot 0
// only (bvh4.go:737-760), boxes rounded outward to float32 (bvh4.go:494-514), primitives reordered so that every
// leaf is a contiguous range -- so every traversal kernel and the Go code path consume it unchanged.
// The TREE differs from the reference's, so parity is "same closest hit", not "same nodes" (tests/test_bvh_device.py).
#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../host/bvh4_builder.hpp"
#include "dscene.cuh"

namespace izpi {

namespace {

struct Box32 {
  float mn[3], mx[3];
};

__device__ __forceinline__ float round_down(double v) {  // conservativeFloat32Min (bvh4.go:494-502)
  float f = (float)v;
  return ((double)f > v) ? nextafterf(f, -INFINITY) : f;
}
__device__ __forceinline__ float round_up(double v) {  // conservativeFloat32Max (bvh4.go:506-514)
  float f = (float)v;
  return ((double)f < v) ? nextafterf(f, INFINITY) : f;
}

__global__ void centroid_kernel(const double* __restrict__ boxes, int n, double* cx, double* cy, double* cz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* b = boxes + 6 * (size_t)i;
  cx[i] = 0.5 * (b[0] + b[3]); cy[i] = 0.5 * (b[1] + b[4]); cz[i] = 0.5 * (b[2] + b[5]);
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void morton_kernel(const double* cx, const double* cy, const double* cz, int n, double lox, double loy, double loz, double sx,
                              double sy, double sz, unsigned long long* keys, int* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto q = [](double c, double lo, double s) {
    double v = (c - lo) * s;
    v = v < 0 ? 0 : (v > 2097151.0 ? 2097151.0 : v);
    return (unsigned long long)v;
  };
  keys[i] = (spread21(q(cx[i], lox, sx)) << 2) | (spread21(q(cy[i], loy, sy)) << 1) | spread21(q(cz[i], loz, sz));
  vals[i] = i;
}

// Karras 2012: longest common prefix of the keys at sorted positions i and j, ties broken by position
__device__ __forceinline__ int delta(const unsigned long long* k, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = k[i], b = k[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

// binary radix tree over n leaves: internal nodes 0..n-2; child index c >= 0 internal, ~c leaf (sorted position)
__global__ void radix_tree_kernel(const unsigned long long* k, int n, int* left, int* right, int* parent_int, int* parent_leaf, int* first,
                                  int* last) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = delta(k, n, i, i + 1) - delta(k, n, i, i - 1) >= 0 ? 1 : -1;
  int dmin = delta(k, n, i, i - d);
  int lmax = 2;
  while (delta(k, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(k, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(k, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(k, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + (d < 0 ? -1 : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  first[i] = lo; last[i] = hi;
  int lc = (lo == gamma) ? ~gamma : gamma;
  int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  left[i] = lc; right[i] = rc;
  if (lc >= 0) parent_int[lc] = i; else parent_leaf[~lc] = i;
  if (rc >= 0) parent_int[rc] = i; else parent_leaf[~rc] = i;
}

__global__ void leaf_box_kernel(const double* __restrict__ boxes, const int* __restrict__ perm, int n, Box32* leaf_box) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* b = boxes + 6 * (size_t)perm[i];
  Box32 o;
  for (int a = 0; a < 3; a++) { o.mn[a] = round_down(b[a]); o.mx[a] = round_up(b[3 + a]); }
  leaf_box[i] = o;
}

__device__ __forceinline__ Box32 merge(const Box32& a, const Box32& b) {
  Box32 o;
  for (int k = 0; k < 3; k++) { o.mn[k] = fminf(a.mn[k], b.mn[k]); o.mx[k] = fmaxf(a.mx[k], b.mx[k]); }
  return o;
}

// bottom-up: the second thread to reach an internal node fits its box from the two finished children
__global__ void fit_kernel(int n, const int* left, const int* right, const int* parent_int, const int* parent_leaf, const Box32* leaf_box,
                           Box32* node_box, int* visits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cur = parent_leaf[i];
  while (cur >= 0) {
    if (atomicAdd(&visits[cur], 1) == 0) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    int l = left[cur], r = right[cur];
    // boxes written by other SMs: read through L2 (an L1 line fetched for a neighbouring 24-byte record may be stale)
    auto ld = [](const Box32* p) {
      Box32 o;
      const float* f = reinterpret_cast<const float*>(p);
      for (int k = 0; k < 3; k++) { o.mn[k] = __ldcg(f + k); o.mx[k] = __ldcg(f + 3 + k); }
      return o;
    };
    Box32 a = l >= 0 ? ld(node_box + l) : leaf_box[~l];
    Box32 b = r >= 0 ? ld(node_box + r) : leaf_box[~r];
    node_box[cur] = merge(a, b);
    __threadfence();
    cur = cur == 0 ? -1 : parent_int[cur];
  }
}

struct Item {
  int bin;   // binary node (>= 0 internal, ~leaf position)
  int node;  // BVH4 node index to write
};

__device__ __forceinline__ int range_size(int c, const int* first, const int* last) { return c >= 0 ? last[c] - first[c] + 1 : 1; }
__device__ __forceinline__ float half_area(const Box32& b) {
  float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ void clear_node(izpi_bvh4_node& n) {  // bvh4.go:725-734
  for (int i = 0; i < 4; i++) {
    n.child_index[i] = -1; n.primitive_count[i] = 0;
    n.min_x[i] = n.min_y[i] = n.min_z[i] = n.max_x[i] = n.max_y[i] = n.max_z[i] = FLT_MAX;
  }
}
__device__ __forceinline__ void set_slot(izpi_bvh4_node& n, int s, const Box32& b, int child, int count) {
  n.min_x[s] = b.mn[0]; n.min_y[s] = b.mn[1]; n.min_z[s] = b.mn[2];
  n.max_x[s] = b.mx[0]; n.max_y[s] = b.mx[1]; n.max_z[s] = b.mx[2];
  n.child_index[s] = child; n.primitive_count[s] = count;
}

// One level of the 4-ary collapse: every item turns a binary subtree root into a BVH4 node whose <= 4 children are found
// by opening, largest box first, binary children that hold more than 4 primitives.  Subtrees of <= 4 primitives become the
// reference's leaf-nodes (own node, slot 0 only).
__global__ void collapse_kernel(const Item* __restrict__ in, int n_in, const int* left, const int* right, const int* first, const int* last,
                                const Box32* node_box, const Box32* leaf_box, izpi_bvh4_node* nodes, int* n_nodes, Item* out, int* n_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  Item it = in[i];
  izpi_bvh4_node nd;
  clear_node(nd);
  auto box_of = [&](int c) { return c >= 0 ? node_box[c] : leaf_box[~c]; };
  if (range_size(it.bin, first, last) <= 4) {  // leaf-node
    int f = it.bin >= 0 ? first[it.bin] : ~it.bin;
    set_slot(nd, 0, box_of(it.bin), f, range_size(it.bin, first, last));
    nodes[it.node] = nd;
    return;
  }
  int ch[4];
  int k = 2;
  ch[0] = left[it.bin]; ch[1] = right[it.bin];
  while (k < 4) {
    int best = -1;
    float best_a = -1.0f;
    for (int c = 0; c < k; c++)
      if (range_size(ch[c], first, last) > 4) {
        float a = half_area(box_of(ch[c]));
        if (a > best_a) { best_a = a; best = c; }
      }
    if (best < 0) break;
    int b = ch[best];
    ch[best] = left[b];
    ch[k++] = right[b];
  }
  int base = atomicAdd(n_nodes, k);
  int qbase = atomicAdd(n_out, k);
  for (int c = 0; c < k; c++) {
    set_slot(nd, c, box_of(ch[c]), base + c, 0);
    out[qbase + c] = Item{ch[c], base + c};
  }
  nodes[it.node] = nd;
}

#define BV_CUDA(call)                                                                    \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); ok = false; goto done; } \
  } while (0)

}  // namespace

// boxes: BoundingBox() of every hitable (fp64).  Returns an empty build on failure (izpi_last_error() says why).
BVH4Build BuildBVH4Device(const std::vector<BoxD>& boxes) {
  BVH4Build out;
  const int n = (int)boxes.size();
  if (n == 0) return out;
  bool ok = true;
  double *d_boxes = nullptr, *d_c[3] = {nullptr, nullptr, nullptr}, *d_red = nullptr;
  unsigned long long *d_k0 = nullptr, *d_k1 = nullptr;
  int *d_v0 = nullptr, *d_v1 = nullptr, *d_left = nullptr, *d_right = nullptr, *d_pi = nullptr, *d_pl = nullptr, *d_first = nullptr, *d_last = nullptr,
      *d_visits = nullptr, *d_counts = nullptr;
  Box32 *d_lbox = nullptr, *d_nbox = nullptr;
  izpi_bvh4_node* d_nodes = nullptr;
  Item *d_q0 = nullptr, *d_q1 = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  const int max_nodes = 2 * n + 1;
  const int T = 256, G = (n + T - 1) / T;
  double red[6];
  {
    BV_CUDA(cudaMalloc(&d_boxes, (size_t)n * 48));
    BV_CUDA(cudaMemcpy(d_boxes, boxes.data(), (size_t)n * 48, cudaMemcpyHostToDevice));
    for (int a = 0; a < 3; a++) BV_CUDA(cudaMalloc(&d_c[a], (size_t)n * 8));
    BV_CUDA(cudaMalloc(&d_red, 6 * 8));
    centroid_kernel<<<G, T>>>(d_boxes, n, d_c[0], d_c[1], d_c[2]);
    for (int a = 0; a < 3; a++) {
      size_t need = 0;
      cub::DeviceReduce::Min(nullptr, need, d_c[a], d_red + a, n);
      if (need > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; BV_CUDA(cudaMalloc(&d_tmp, need)); tmp_bytes = need; }
      cub::DeviceReduce::Min(d_tmp, need, d_c[a], d_red + a, n);
      cub::DeviceReduce::Max(d_tmp, need, d_c[a], d_red + 3 + a, n);
    }
    BV_CUDA(cudaMemcpy(red, d_red, sizeof(red), cudaMemcpyDeviceToHost));
    double s[3];
    for (int a = 0; a < 3; a++) { double ext = red[3 + a] - red[a]; s[a] = ext > 0 ? 2097151.0 / ext : 0.0; }
    BV_CUDA(cudaMalloc(&d_k0, (size_t)n * 8)); BV_CUDA(cudaMalloc(&d_k1, (size_t)n * 8));
    BV_CUDA(cudaMalloc(&d_v0, (size_t)n * 4)); BV_CUDA(cudaMalloc(&d_v1, (size_t)n * 4));
    morton_kernel<<<G, T>>>(d_c[0], d_c[1], d_c[2], n, red[0], red[1], red[2], s[0], s[1], s[2], d_k0, d_v0);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, d_k0, d_k1, d_v0, d_v1, n, 0, 63);
    if (need > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; BV_CUDA(cudaMalloc(&d_tmp, need)); tmp_bytes = need; }
    cub::DeviceRadixSort::SortPairs(d_tmp, need, d_k0, d_k1, d_v0, d_v1, n, 0, 63);  // stable: equal codes keep input order
    for (int a = 0; a < 3; a++) { cudaFree(d_c[a]); d_c[a] = nullptr; }
    BV_CUDA(cudaMalloc(&d_lbox, (size_t)n * sizeof(Box32)));
    leaf_box_kernel<<<G, T>>>(d_boxes, d_v1, n, d_lbox);
    BV_CUDA(cudaMalloc(&d_nodes, (size_t)max_nodes * sizeof(izpi_bvh4_node)));
    BV_CUDA(cudaMalloc(&d_counts, 4 * sizeof(int)));
    out.perm.resize(n);
    BV_CUDA(cudaMemcpy(out.perm.data(), d_v1, (size_t)n * 4, cudaMemcpyDeviceToHost));
    int n_nodes = 1;
    if (n == 1) {  // a single hitable: one leaf-node (bvh4.go:604-610)
      BoxD b = boxes[0];
      izpi_bvh4_node nd;
      for (int i = 0; i < 4; i++) {
        nd.child_index[i] = -1; nd.primitive_count[i] = 0;
        nd.min_x[i] = nd.min_y[i] = nd.min_z[i] = nd.max_x[i] = nd.max_y[i] = nd.max_z[i] = FLT_MAX;
      }
      nd.min_x[0] = ConservativeFloat32Min(b.mn[0]); nd.min_y[0] = ConservativeFloat32Min(b.mn[1]); nd.min_z[0] = ConservativeFloat32Min(b.mn[2]);
      nd.max_x[0] = ConservativeFloat32Max(b.mx[0]); nd.max_y[0] = ConservativeFloat32Max(b.mx[1]); nd.max_z[0] = ConservativeFloat32Max(b.mx[2]);
      nd.child_index[0] = 0; nd.primitive_count[0] = 1;
      out.nodes.push_back(nd);
      goto done;
    }
    BV_CUDA(cudaMalloc(&d_left, (size_t)n * 4)); BV_CUDA(cudaMalloc(&d_right, (size_t)n * 4));
    BV_CUDA(cudaMalloc(&d_pi, (size_t)n * 4)); BV_CUDA(cudaMalloc(&d_pl, (size_t)n * 4));
    BV_CUDA(cudaMalloc(&d_first, (size_t)n * 4)); BV_CUDA(cudaMalloc(&d_last, (size_t)n * 4));
    BV_CUDA(cudaMalloc(&d_visits, (size_t)n * 4));
    BV_CUDA(cudaMemset(d_visits, 0, (size_t)n * 4));
    BV_CUDA(cudaMemset(d_pi, 0xff, (size_t)n * 4));
    radix_tree_kernel<<<G, T>>>(d_k1, n, d_left, d_right, d_pi, d_pl, d_first, d_last);
    BV_CUDA(cudaMalloc(&d_nbox, (size_t)n * sizeof(Box32)));
    fit_kernel<<<G, T>>>(n, d_left, d_right, d_pi, d_pl, d_lbox, d_nbox, d_visits);
    BV_CUDA(cudaGetLastError());
    // 4-ary collapse, level by level
    BV_CUDA(cudaMalloc(&d_q0, (size_t)max_nodes * sizeof(Item))); BV_CUDA(cudaMalloc(&d_q1, (size_t)max_nodes * sizeof(Item)));
    Item root{0, 0};
    BV_CUDA(cudaMemcpy(d_q0, &root, sizeof(root), cudaMemcpyHostToDevice));
    int n_in = 1;
    for (int level = 0; n_in > 0 && level < 128; level++) {
      int counts[2] = {n_nodes, 0};
      BV_CUDA(cudaMemcpy(d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice));
      collapse_kernel<<<(n_in + T - 1) / T, T>>>(d_q0, n_in, d_left, d_right, d_first, d_last, d_nbox, d_lbox, d_nodes, d_counts, d_q1, d_counts + 1);
      BV_CUDA(cudaMemcpy(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost));
      n_nodes = counts[0]; n_in = counts[1];
      std::swap(d_q0, d_q1);
    }
    BV_CUDA(cudaGetLastError());
    out.nodes.resize(n_nodes);
    BV_CUDA(cudaMemcpy(out.nodes.data(), d_nodes, (size_t)n_nodes * sizeof(izpi_bvh4_node), cudaMemcpyDeviceToHost));
  }
done:
  cudaFree(d_boxes); for (int a = 0; a < 3; a++) cudaFree(d_c[a]);
  cudaFree(d_red); cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_v0); cudaFree(d_v1); cudaFree(d_left); cudaFree(d_right); cudaFree(d_pi);
  cudaFree(d_pl); cudaFree(d_first); cudaFree(d_last); cudaFree(d_visits); cudaFree(d_counts); cudaFree(d_lbox); cudaFree(d_nbox);
  cudaFree(d_nodes); cudaFree(d_q0); cudaFree(d_q1); cudaFree(d_tmp);
  if (!ok) { out.nodes.clear(); out.perm.clear(); return out; }
  {
    // The reference's traversal has a fixed 64-entry stack (bvh4.go:71) and would index past it on a deeper tree;
    // bound the worst case here: visiting a node with k hit children leaves k-1 entries below the child being visited.
    // Children are numbered after their parents, so one reverse sweep suffices.
    std::vector<int> need(out.nodes.size(), 0);
    for (int i = (int)out.nodes.size() - 1; i >= 0; i--) {
      const izpi_bvh4_node& nd = out.nodes[i];
      int k = 0, deepest = 0;
      for (int c = 0; c < 4; c++)
        if (nd.child_index[c] != -1 && nd.primitive_count[c] == 0) { k++; deepest = std::max(deepest, need[nd.child_index[c]]); }
      need[i] = k > 0 ? (k - 1) + deepest : 0;
    }
    if (need[0] > 64) {
      set_error("device BVH4 build: tree needs " + std::to_string(need[0]) + " stack entries, the traversal has 64 (bvh4.go:71)");
      out.nodes.clear(); out.perm.clear();
    }
  }
  return out;
}

}  // namespace izpi

namespace {
struct BuildResult {
  izpi::BVH4Build b;
};
}  // namespace

void bvh_build_result_free(izpi_ctx* ctx) {
  delete static_cast<BuildResult*>(ctx->bvh_build);
  ctx->bvh_build = nullptr;
}

extern "C" {

int izpi_bvh4_build(izpi_ctx* ctx, int32_t n, const double* boxes6, int32_t* n_nodes) {
  if (!ctx || n < 0 || (n > 0 && !boxes6) || !n_nodes) { izpi::set_error("izpi_bvh4_build: bad argument"); return IZPI_EINVAL; }
  IZ_CUDA(cudaSetDevice(ctx->device));
  bvh_build_result_free(ctx);
  std::vector<izpi::BoxD> boxes((size_t)n);
  if (n) std::memcpy(boxes.data(), boxes6, (size_t)n * sizeof(izpi::BoxD));
  auto* r = new BuildResult();
  r->b = izpi::BuildBVH4Device(boxes);
  if (n > 0 && r->b.nodes.empty()) { delete r; return IZPI_ECUDA; }  // message set by the builder
  ctx->bvh_build = r;
  ctx->launches += 8;  // centroid, morton, sort, leaf boxes, radix tree, fit, collapse levels (lower bound)
  *n_nodes = (int32_t)r->b.nodes.size();
  return IZPI_OK;
}

int izpi_bvh4_build_fetch(izpi_ctx* ctx, izpi_bvh4_node* nodes, int32_t* perm) {
  if (!ctx || !ctx->bvh_build) { izpi::set_error("izpi_bvh4_build_fetch: no build result"); return IZPI_ESTATE; }
  auto* r = static_cast<BuildResult*>(ctx->bvh_build);
  if (nodes) std::memcpy(nodes, r->b.nodes.data(), r->b.nodes.size() * sizeof(izpi_bvh4_node));
  if (perm) std::memcpy(perm, r->b.perm.data(), r->b.perm.size() * sizeof(int32_t));
  return IZPI_OK;
}

}  // extern "C"
