// oracle_bvh.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of hitable/bvh4.go (build + traversal) and of the two RayAABB4_SIMD
// flavours (bvh4_simd_amd64.go, bvh4_simd_generic.go).
//
// Third-party algorithm restated here: Go's sort.Slice (stdlib `sort`, pdqsort_func in
// zsortfunc.go, Go >= 1.19; the reference pins go 1.26 in go.mod:1).  The reference calls it at
// bvh4.go:669-703.  It is restated from the published algorithm; for n <= 12 it is insertion
// sort (pinned by bvh_node_test.go:13-137); for n > 12 with duplicate keys the tie order is
// PARITY UNPINNED (no Go toolchain here to confirm).  Harmless for the closest-hit parity
// tests because device and oracle consume the same node array.
#include "oracle_geom.hpp"

#include <algorithm>
#include <cstring>

namespace orc {

thread_local Stats* g_stats = nullptr;

// ------------------------------------------------------------------ float32 helpers
float conservativeFloat32Min(double v) {  // bvh4.go:494-502
  float f = (float)v;
  if ((double)f > v) return std::nextafterf(f, -std::numeric_limits<float>::infinity());
  return f;
}
float conservativeFloat32Max(double v) {  // bvh4.go:506-514
  float f = (float)v;
  if ((double)f < v) return std::nextafterf(f, std::numeric_limits<float>::infinity());
  return f;
}

// MINPS/MAXPS semantics of `x.Min(y)` / `x.Max(y)`: second operand wins on NaN / equal.
static inline float sse_min(float x, float y) { return x < y ? x : y; }
static inline float sse_max(float x, float y) { return x > y ? x : y; }
// bvh4.go:478-490
static inline float min32(float a, float b) { return a < b ? a : b; }
static inline float max32(float a, float b) { return a > b ? a : b; }

uint8_t RayAABB4(int flavour, const float org[3], const float inv[3], const izpi_bvh4_node& n, float tMaxParam) {
  uint8_t mask = 0;
  for (int i = 0; i < 4; i++) {
    if (flavour == BOX_SSE) {  // bvh4_simd_amd64.go:27-110, one lane
      float t0x = (n.min_x[i] - org[0]) * inv[0];
      float t1x = (n.max_x[i] - org[0]) * inv[0];
      float tMin = sse_min(t0x, t1x);
      float tMax = sse_max(t0x, t1x);
      float t0y = (n.min_y[i] - org[1]) * inv[1];
      float t1y = (n.max_y[i] - org[1]) * inv[1];
      float tNearY = sse_min(t0y, t1y), tFarY = sse_max(t0y, t1y);
      tMin = sse_max(tMin, tNearY);
      tMax = sse_min(tMax, tFarY);
      float t0z = (n.min_z[i] - org[2]) * inv[2];
      float t1z = (n.max_z[i] - org[2]) * inv[2];
      float tNearZ = sse_min(t0z, t1z), tFarZ = sse_max(t0z, t1z);
      tMin = sse_max(tMin, tNearZ);
      tMax = sse_min(tMax, tFarZ);
      bool c1 = tMax >= tMin, c2 = tMax >= 0.0f, c3 = tMaxParam >= tMin;
      if (c1 && c2 && c3) mask |= (uint8_t)(1u << i);
    } else {  // bvh4_simd_generic.go:10-52 (== arm64.go:27-69 == test's rayAABB4_Reference)
      float t0x = (n.min_x[i] - org[0]) * inv[0];
      float t1x = (n.max_x[i] - org[0]) * inv[0];
      if (t0x > t1x) std::swap(t0x, t1x);
      float t0y = (n.min_y[i] - org[1]) * inv[1];
      float t1y = (n.max_y[i] - org[1]) * inv[1];
      if (t0y > t1y) std::swap(t0y, t1y);
      float t0z = (n.min_z[i] - org[2]) * inv[2];
      float t1z = (n.max_z[i] - org[2]) * inv[2];
      if (t0z > t1z) std::swap(t0z, t1z);
      float tNear = max32(max32(t0x, t0y), t0z);
      float tFar = min32(min32(t1x, t1y), t1z);
      if (tNear <= tFar && tFar >= 0 && tNear <= tMaxParam) mask |= (uint8_t)(1u << i);
    }
  }
  return mask;
}

// ------------------------------------------------------------------ traversal (bvh4.go:49-164)
bool BVH4::Hit(const Ray& r, double tMin, double tMax, HitRecord& best, const Material*& bestMat) const {
  if (Nodes.empty()) return false;
  bool hitFound = false;
  float inv[3] = {(float)(1.0 / r.direction.X), (float)(1.0 / r.direction.Y), (float)(1.0 / r.direction.Z)};
  float org[3] = {(float)r.origin.X, (float)r.origin.Y, (float)r.origin.Z};
  int32_t stack[64];
  int32_t stackPtr = 0;
  int32_t cur = 0;
  HitRecord rec; const Material* m = nullptr;
  for (;;) {
    if (cur == -1) break;
    if (cur >= (int32_t)Nodes.size()) break;
    const izpi_bvh4_node& node = Nodes[cur];
    if (g_stats) g_stats->nodes++;
    uint8_t mask = RayAABB4(flavour, org, inv, node, (float)tMax);
    int32_t next = -1;
    for (int i = 0; i < 4; i++) {
      if (((mask >> i) & 1) == 0) continue;
      int32_t childIndex = node.child_index[i];
      int32_t primitiveCount = node.primitive_count[i];
      if (childIndex == -1) continue;
      if (primitiveCount > 0) {
        for (int32_t p = 0; p < primitiveCount; p++) {
          if (Primitives[childIndex + p]->Hit(r, tMin, tMax, rec, m)) {
            tMax = rec.t; best = rec; bestMat = m; hitFound = true;
          }
        }
      } else {
        if (next == -1) next = childIndex;
        else stack[stackPtr++] = childIndex;
      }
    }
    if (next != -1) cur = next;
    else if (stackPtr > 0) cur = stack[--stackPtr];
    else cur = -1;
  }
  return hitFound;
}

bool BVH4::BoundingBox(AABB& box) const {  // bvh4.go:265-296
  if (Nodes.empty()) return false;
  const izpi_bvh4_node& root = Nodes[0];
  double mnx = root.min_x[0], mny = root.min_y[0], mnz = root.min_z[0];
  double mxx = root.max_x[0], mxy = root.max_y[0], mxz = root.max_z[0];
  for (int i = 1; i < 4; i++) {
    if (root.child_index[i] != -1) {
      mnx = gomin(mnx, root.min_x[i]); mny = gomin(mny, root.min_y[i]); mnz = gomin(mnz, root.min_z[i]);
      mxx = gomax(mxx, root.max_x[i]); mxy = gomax(mxy, root.max_y[i]); mxz = gomax(mxz, root.max_z[i]);
    }
  }
  box.min = V(mnx, mny, mnz); box.max = V(mxx, mxy, mxz);
  return true;
}

// ------------------------------------------------------------------ Go sort.Slice restated
namespace gosort {
// data.Less(i,j) / data.Swap(i,j) over the `pairs` slice of bvh4.go:662-665; the key of a
// pair is box.min[axis] of its hitable (aabb.BoxLessX/Y/Z, aabb.go:42-54).
struct LessSwap {
  int* w;              // working indices (the `pairs`)
  const double* keys;  // keys[prim] = box.min[axis]
  bool Less(int i, int j) const { return keys[w[i]] < keys[w[j]]; }
  void Swap(int i, int j) { int t = w[i]; w[i] = w[j]; w[j] = t; }
};
static void insertionSort(LessSwap& d, int a, int b) {
  for (int i = a + 1; i < b; i++)
    for (int j = i; j > a && d.Less(j, j - 1); j--) d.Swap(j, j - 1);
}
static void siftDown(LessSwap& d, int lo, int hi, int first) {
  int root = lo;
  for (;;) {
    int child = 2 * root + 1;
    if (child >= hi) return;
    if (child + 1 < hi && d.Less(first + child, first + child + 1)) child++;
    if (!d.Less(first + root, first + child)) return;
    d.Swap(first + root, first + child);
    root = child;
  }
}
static void heapSort(LessSwap& d, int a, int b) {
  int first = a, lo = 0, hi = b - a;
  for (int i = (hi - 1) / 2; i >= 0; i--) siftDown(d, i, hi, first);
  for (int i = hi - 1; i >= 0; i--) { d.Swap(first, first + i); siftDown(d, lo, i, first); }
}
static int bitsLen(unsigned long long x) { int n = 0; while (x) { n++; x >>= 1; } return n; }
static void breakPatterns(LessSwap& d, int a, int b) {
  int length = b - a;
  if (length >= 8) {
    uint64_t random = (uint64_t)length;
    unsigned long long modulus = 1ull << bitsLen((unsigned long long)length);
    int idx = a + (length / 4) * 2 - 1;
    for (int i = 0; i < 3; i++) {
      random ^= random << 13; random ^= random >> 7; random ^= random << 17;
      int other = (int)((unsigned long long)random & (modulus - 1));
      if (other >= length) other -= length;
      d.Swap(idx - 1 + i, a + other);
    }
  }
}
enum Hint { unknownHint = 0, increasingHint, decreasingHint };
static void order2(LessSwap& d, int& a, int& b, int* swaps) {
  if (d.Less(b, a)) { (*swaps)++; std::swap(a, b); }
}
static int median(LessSwap& d, int a, int b, int c, int* swaps) {
  order2(d, a, b, swaps); order2(d, b, c, swaps); order2(d, a, b, swaps);
  return b;
}
static int medianAdjacent(LessSwap& d, int a, int* swaps) { return median(d, a - 1, a, a + 1, swaps); }
static int choosePivot(LessSwap& d, int a, int b, Hint* hint) {
  const int shortestNinther = 50, maxSwaps = 4 * 3;
  int l = b - a, swaps = 0;
  int i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
  if (l >= 8) {
    if (l >= shortestNinther) {
      i = medianAdjacent(d, i, &swaps); j = medianAdjacent(d, j, &swaps); k = medianAdjacent(d, k, &swaps);
    }
    j = median(d, i, j, k, &swaps);
  }
  if (swaps == 0) *hint = increasingHint;
  else if (swaps == maxSwaps) *hint = decreasingHint;
  else *hint = unknownHint;
  return j;
}
static void reverseRange(LessSwap& d, int a, int b) {
  int i = a, j = b - 1;
  while (i < j) { d.Swap(i, j); i++; j--; }
}
static bool partialInsertionSort(LessSwap& d, int a, int b) {
  const int maxSteps = 5, shortestShifting = 50;
  int i = a + 1;
  for (int j = 0; j < maxSteps; j++) {
    while (i < b && !d.Less(i, i - 1)) i++;
    if (i == b) return true;
    if (b - a < shortestShifting) return false;
    d.Swap(i, i - 1);
    if (i - a >= 2) {
      for (int jj = i - 1; jj >= 1; jj--) { if (!d.Less(jj, jj - 1)) break; d.Swap(jj, jj - 1); }
    }
    if (b - i >= 2) {
      for (int jj = i + 1; jj < b; jj++) { if (!d.Less(jj, jj - 1)) break; d.Swap(jj, jj - 1); }
    }
  }
  return false;
}
static int partitionEqual(LessSwap& d, int a, int b, int pivot) {
  d.Swap(a, pivot);
  int i = a + 1, j = b - 1;
  for (;;) {
    while (i <= j && !d.Less(a, i)) i++;
    while (i <= j && d.Less(a, j)) j--;
    if (i > j) break;
    d.Swap(i, j); i++; j--;
  }
  return i;
}
static int partition(LessSwap& d, int a, int b, int pivot, bool* already) {
  d.Swap(a, pivot);
  int i = a + 1, j = b - 1;
  while (i <= j && d.Less(i, a)) i++;
  while (i <= j && !d.Less(j, a)) j--;
  if (i > j) { d.Swap(j, a); *already = true; return j; }
  d.Swap(i, j); i++; j--;
  for (;;) {
    while (i <= j && d.Less(i, a)) i++;
    while (i <= j && !d.Less(j, a)) j--;
    if (i > j) break;
    d.Swap(i, j); i++; j--;
  }
  d.Swap(j, a);
  *already = false;
  return j;
}
static void pdqsort(LessSwap& d, int a, int b, int limit) {
  const int maxInsertion = 12;
  bool wasBalanced = true, wasPartitioned = true;
  for (;;) {
    int length = b - a;
    if (length <= maxInsertion) { insertionSort(d, a, b); return; }
    if (limit == 0) { heapSort(d, a, b); return; }
    if (!wasBalanced) { breakPatterns(d, a, b); limit--; }
    Hint hint;
    int pivot = choosePivot(d, a, b, &hint);
    if (hint == decreasingHint) {
      reverseRange(d, a, b);
      pivot = (b - 1) - (pivot - a);
      hint = increasingHint;
    }
    if (wasBalanced && wasPartitioned && hint == increasingHint) {
      if (partialInsertionSort(d, a, b)) return;
    }
    if (a > 0 && !d.Less(a - 1, pivot)) { a = partitionEqual(d, a, b, pivot); continue; }
    bool already;
    int mid = partition(d, a, b, pivot, &already);
    wasPartitioned = already;
    int leftLen = mid - a, rightLen = b - mid;
    int balanceThreshold = length / 8;
    if (leftLen < rightLen) {
      wasBalanced = leftLen >= balanceThreshold;
      pdqsort(d, a, mid, limit);
      a = mid + 1;
    } else {
      wasBalanced = rightLen >= balanceThreshold;
      pdqsort(d, mid + 1, b, limit);
      b = mid;
    }
  }
}
static void Slice(int length, LessSwap& d) { pdqsort(d, 0, length, bitsLen((unsigned long long)length)); }
}  // namespace gosort

// ------------------------------------------------------------------ build (bvh4.go:551-855)
struct buildNode {
  AABB box; bool hasBox = false;
  std::vector<buildNode*> children;
  std::vector<int> primitiveIndices;
};
struct BuildCtx {
  const std::vector<Hitable*>* all;
  std::vector<AABB> boxes;  // BoundingBox(time0,time1) of hitable i (pure function of the hitable)
  std::vector<double> keys[3];  // box.min.{X,Y,Z}
  Rng* lcg; bool rand_zero;
  std::vector<std::unique_ptr<buildNode>> pool;
  double randomFunc() { return rand_zero ? 0.0 : lcg->Float64(); }
};

static buildNode* buildBinaryBVH(BuildCtx& c, const std::vector<int>& indices) {  // bvh4.go:596-652
  if (indices.empty()) return nullptr;
  c.pool.push_back(std::make_unique<buildNode>());
  buildNode* node = c.pool.back().get();
  if (indices.size() == 1) {
    node->primitiveIndices = indices;
    node->box = c.boxes[indices[0]]; node->hasBox = true;
    return node;
  }
  AABB overall = c.boxes[indices[0]];
  for (size_t i = 1; i < indices.size(); i++) overall = SurroundingBox(overall, c.boxes[indices[i]]);
  node->box = overall; node->hasBox = true;
  int axis = (int)(3 * c.randomFunc());
  std::vector<int> working(indices);
  {  // sortHitablesWithIndices (bvh4.go:655-711): sort.Slice on box.min[axis] with `<`
    gosort::LessSwap d{working.data(), c.keys[axis].data()};
    gosort::Slice((int)working.size(), d);
  }
  if (working.size() <= 4) { node->primitiveIndices = working; return node; }
  size_t mid = working.size() / 2;
  std::vector<int> left(working.begin(), working.begin() + mid), right(working.begin() + mid, working.end());
  buildNode* l = buildBinaryBVH(c, left);
  buildNode* r = buildBinaryBVH(c, right);
  node->children = {l, r};
  return node;
}

static std::vector<buildNode*> collectChildren(buildNode* node, int maxChildren) {  // bvh4.go:796-855
  if (node == nullptr || !node->primitiveIndices.empty()) return {node};
  if (node->children.empty()) return {node};
  std::vector<buildNode*> result;
  for (buildNode* ch : node->children) if (ch) result.push_back(ch);
  bool expanded = true;
  while (expanded && (int)result.size() < maxChildren) {
    expanded = false;
    for (size_t i = 0; i < result.size(); i++) {
      buildNode* cur = result[i];
      if (!cur->primitiveIndices.empty()) continue;
      if (cur->children.empty()) continue;
      int numAfter = (int)result.size() - 1 + (int)cur->children.size();
      if (numAfter <= maxChildren) {
        result.erase(result.begin() + i);
        for (buildNode* ch : cur->children) result.push_back(ch);
        expanded = true;
        break;
      }
    }
  }
  if ((int)result.size() > maxChildren) result.resize(maxChildren);
  return result;
}

static int32_t flattenBVH4(buildNode* node, BVH4& bvh, std::vector<int>& primitiveIndices) {  // bvh4.go:714-792
  if (!node) return -1;
  int32_t nodeIndex = (int32_t)bvh.Nodes.size();
  izpi_bvh4_node n;
  for (int i = 0; i < 4; i++) {
    n.child_index[i] = -1; n.primitive_count[i] = 0;
    n.min_x[i] = n.min_y[i] = n.min_z[i] = n.max_x[i] = n.max_y[i] = n.max_z[i] = FLT_MAX;
  }
  if (!node->primitiveIndices.empty()) {
    int32_t primStart = (int32_t)primitiveIndices.size();
    primitiveIndices.insert(primitiveIndices.end(), node->primitiveIndices.begin(), node->primitiveIndices.end());
    n.child_index[0] = primStart;
    n.primitive_count[0] = (int32_t)node->primitiveIndices.size();
    if (node->hasBox) {
      n.min_x[0] = conservativeFloat32Min(node->box.min.X); n.min_y[0] = conservativeFloat32Min(node->box.min.Y);
      n.min_z[0] = conservativeFloat32Min(node->box.min.Z); n.max_x[0] = conservativeFloat32Max(node->box.max.X);
      n.max_y[0] = conservativeFloat32Max(node->box.max.Y); n.max_z[0] = conservativeFloat32Max(node->box.max.Z);
    }
    bvh.Nodes.push_back(n);
    return nodeIndex;
  }
  std::vector<buildNode*> children = collectChildren(node, 4);
  bvh.Nodes.push_back(n);
  for (size_t i = 0; i < children.size() && i < 4; i++) {
    buildNode* child = children[i];
    int32_t childIndex = flattenBVH4(child, bvh, primitiveIndices);
    bvh.Nodes[nodeIndex].child_index[i] = childIndex;
    if (child->hasBox) {
      izpi_bvh4_node& nn = bvh.Nodes[nodeIndex];
      nn.min_x[i] = conservativeFloat32Min(child->box.min.X); nn.min_y[i] = conservativeFloat32Min(child->box.min.Y);
      nn.min_z[i] = conservativeFloat32Min(child->box.min.Z); nn.max_x[i] = conservativeFloat32Max(child->box.max.X);
      nn.max_y[i] = conservativeFloat32Max(child->box.max.Y); nn.max_z[i] = conservativeFloat32Max(child->box.max.Z);
    }
  }
  return nodeIndex;
}

std::unique_ptr<BVH4> newBVH4(const std::vector<Hitable*>& hitables, Rng* lcg, bool rand_zero) {  // bvh4.go:558-593
  if (hitables.empty()) return nullptr;
  auto bvh = std::make_unique<BVH4>();
  BuildCtx c;
  c.all = &hitables; c.lcg = lcg; c.rand_zero = rand_zero;
  c.boxes.resize(hitables.size());
  for (int a = 0; a < 3; a++) c.keys[a].resize(hitables.size());
  for (size_t i = 0; i < hitables.size(); i++) {
    hitables[i]->BoundingBox(c.boxes[i]);
    c.keys[0][i] = c.boxes[i].min.X; c.keys[1][i] = c.boxes[i].min.Y; c.keys[2][i] = c.boxes[i].min.Z;
  }
  std::vector<int> indices(hitables.size());
  for (size_t i = 0; i < indices.size(); i++) indices[i] = (int)i;
  buildNode* root = buildBinaryBVH(c, indices);
  std::vector<int> primitiveIndices;
  flattenBVH4(root, *bvh, primitiveIndices);
  bvh->Primitives.resize(primitiveIndices.size());
  bvh->PrimitiveIndices.resize(primitiveIndices.size());
  for (size_t i = 0; i < primitiveIndices.size(); i++) {
    bvh->Primitives[i] = hitables[primitiveIndices[i]];
    bvh->PrimitiveIndices[i] = primitiveIndices[i];
  }
  return bvh;
}

}  // namespace orc
