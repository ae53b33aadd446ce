// bvh4_builder.cpp -- see bvh4_builder.hpp.  Mirrors hitable.newBVH4 (bvh4.go:558-855).
#include "bvh4_builder.hpp"

#include <cfloat>
#include <cmath>
#include <limits>
#include <thread>
#include <unordered_map>

namespace izpi {

float ConservativeFloat32Min(double v) {
  float f = (float)v;
  return ((double)f > v) ? nextafterf(f, -INFINITY) : f;
}
float ConservativeFloat32Max(double v) {
  float f = (float)v;
  return ((double)f < v) ? nextafterf(f, INFINITY) : f;
}

namespace {

// ---------------------------------------------------------------------------------------
// Ordering of Go's sort.Slice (pattern-defeating quicksort, stdlib sort/zsortfunc.go) over a
// window of primitive indices keyed by box.min[axis].  The reference relies on it at
// bvh4.go:669-703; ties between equal keys resolve exactly as this sequence of swaps does.
// Positions are relative to the window start, as they are for the slice Go sorts.
class GoSliceSorter {
 public:
  GoSliceSorter(int32_t* window, const double* key) : w_(window), k_(key) {}
  void sort(int n) {
    int limit = 0;
    for (unsigned v = (unsigned)n; v; v >>= 1) limit++;  // bits.Len(uint(n))
    pdq(0, n, limit);
  }

 private:
  int32_t* w_;
  const double* k_;
  bool less(int i, int j) const { return k_[w_[i]] < k_[w_[j]]; }
  void swap(int i, int j) { int32_t t = w_[i]; w_[i] = w_[j]; w_[j] = t; }

  void insertion(int a, int b) {
    for (int i = a + 1; i < b; i++)
      for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
  }
  void sift(int lo, int hi, int first) {
    for (int root = lo;;) {
      int child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  void heap(int a, int b) {
    int hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) sift(i, hi, a);
    for (int i = hi - 1; i >= 0; i--) { swap(a, a + i); sift(0, i, a); }
  }
  void scramble(int a, int b) {  // breakPatterns
    int n = b - a;
    if (n < 8) return;
    uint64_t r = (uint64_t)n;
    unsigned shift = 0;
    for (unsigned v = (unsigned)n; v; v >>= 1) shift++;
    uint64_t mask = (1ull << shift) - 1;
    int idx = a + (n / 4) * 2 - 1;
    for (int i = 0; i < 3; i++) {
      r ^= r << 13; r ^= r >> 7; r ^= r << 17;
      int other = (int)(r & mask);
      if (other >= n) other -= n;
      swap(idx - 1 + i, a + other);
    }
  }
  int med3(int a, int b, int c, int& swaps) {
    if (less(b, a)) { swaps++; int t = a; a = b; b = t; }
    if (less(c, b)) { swaps++; int t = b; b = c; c = t; }
    if (less(b, a)) { swaps++; int t = a; a = b; b = t; }
    return b;
  }
  // returns pivot; hint: 0 unknown, 1 increasing, 2 decreasing
  int pivot(int a, int b, int& hint) {
    int n = b - a, swaps = 0;
    int i = a + n / 4 * 1, j = a + n / 4 * 2, k = a + n / 4 * 3;
    if (n >= 8) {
      if (n >= 50) {
        i = med3(i - 1, i, i + 1, swaps);
        j = med3(j - 1, j, j + 1, swaps);
        k = med3(k - 1, k, k + 1, swaps);
      }
      j = med3(i, j, k, swaps);
    }
    hint = swaps == 0 ? 1 : (swaps == 12 ? 2 : 0);
    return j;
  }
  bool partialInsertion(int a, int b) {
    int i = a + 1;
    for (int step = 0; step < 5; step++) {
      while (i < b && !less(i, i - 1)) i++;
      if (i == b) return true;
      if (b - a < 50) return false;
      swap(i, i - 1);
      if (i - a >= 2)
        for (int j = i - 1; j >= 1; j--) { if (!less(j, j - 1)) break; swap(j, j - 1); }
      if (b - i >= 2)
        for (int j = i + 1; j < b; j++) { if (!less(j, j - 1)) break; swap(j, j - 1); }
    }
    return false;
  }
  int splitEqual(int a, int b, int p) {
    swap(a, p);
    int i = a + 1, j = b - 1;
    for (;;) {
      while (i <= j && !less(a, i)) i++;
      while (i <= j && less(a, j)) j--;
      if (i > j) break;
      swap(i, j); i++; j--;
    }
    return i;
  }
  int split(int a, int b, int p, bool& already) {
    swap(a, p);
    int i = a + 1, j = b - 1;
    while (i <= j && less(i, a)) i++;
    while (i <= j && !less(j, a)) j--;
    if (i > j) { swap(j, a); already = true; return j; }
    swap(i, j); i++; j--;
    for (;;) {
      while (i <= j && less(i, a)) i++;
      while (i <= j && !less(j, a)) j--;
      if (i > j) break;
      swap(i, j); i++; j--;
    }
    swap(j, a);
    already = false;
    return j;
  }
  void pdq(int a, int b, int limit) {
    bool balanced = true, partitioned = true;
    for (;;) {
      int n = b - a;
      if (n <= 12) { insertion(a, b); return; }
      if (limit == 0) { heap(a, b); return; }
      if (!balanced) { scramble(a, b); limit--; }
      int hint;
      int p = pivot(a, b, hint);
      if (hint == 2) {
        for (int i = a, j = b - 1; i < j; i++, j--) swap(i, j);
        p = (b - 1) - (p - a);
        hint = 1;
      }
      if (balanced && partitioned && hint == 1 && partialInsertion(a, b)) return;
      if (a > 0 && !less(a - 1, p)) { a = splitEqual(a, b, p); continue; }
      bool already;
      int mid = split(a, b, p, already);
      partitioned = already;
      int left = mid - a, right = b - mid;
      if (left < right) {
        balanced = left >= n / 8;
        pdq(a, mid, limit);
        a = mid + 1;
      } else {
        balanced = right >= n / 8;
        pdq(mid + 1, b, limit);
        b = mid;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
struct Affine {  // x -> a*x + c (mod 2^32): one LCG step or a power of it
  uint32_t a, c;
};
Affine compose(Affine f, Affine g) { return Affine{f.a * g.a, f.a * g.c + f.c}; }  // f(g(x))
uint32_t lcg_jump(uint32_t s, uint64_t k) {
  Affine step{1664525u, 1013904223u}, acc{1u, 0u};
  for (; k; k >>= 1) {
    if (k & 1) acc = compose(step, acc);
    step = compose(step, step);
  }
  return acc.a * s + acc.c;
}

struct BinNode {
  int32_t lo, hi;       // window of `order`
  int32_t left, right;  // -1 for a leaf
  BoxD box;
};

struct Builder {
  const std::vector<BoxD>& boxes;
  std::vector<double> key[3];
  std::vector<int32_t> order;
  std::vector<BinNode> bin;
  uint32_t seed32;
  bool rand_zero;
  std::unordered_map<int, uint64_t> nodes_memo, draws_memo;

  explicit Builder(const std::vector<BoxD>& b) : boxes(b) {}

  // sizes of the binary tree for a window of m primitives (bvh4.go:596-652)
  uint64_t count_nodes(int m) {
    if (m <= 4) return 1;
    auto it = nodes_memo.find(m);
    if (it != nodes_memo.end()) return it->second;
    uint64_t v = 1 + count_nodes(m / 2) + count_nodes(m - m / 2);
    nodes_memo[m] = v;
    return v;
  }
  uint64_t count_draws(int m) {
    if (m <= 1) return 0;
    if (m <= 4) return 1;
    auto it = draws_memo.find(m);
    if (it != draws_memo.end()) return it->second;
    uint64_t v = 1 + count_draws(m / 2) + count_draws(m - m / 2);
    draws_memo[m] = v;
    return v;
  }
  void warm(int m) {  // fill the memo tables single-threaded so the parallel phase only reads them
    if (m <= 4 || nodes_memo.count(m)) return;
    count_nodes(m); count_draws(m);
    warm(m / 2); warm(m - m / 2);
  }
  uint64_t nodes_of(int m) const { return m <= 4 ? 1 : nodes_memo.at(m); }
  uint64_t draws_of(int m) const { return m <= 1 ? 0 : (m <= 4 ? 1 : draws_memo.at(m)); }

  void leaf_box(BinNode& n) {
    n.box = boxes[order[n.lo]];
    for (int i = n.lo + 1; i < n.hi; i++) {
      const BoxD& b = boxes[order[i]];
      for (int a = 0; a < 3; a++) {
        if (b.mn[a] < n.box.mn[a]) n.box.mn[a] = b.mn[a];  // aabb.SurroundingBox (aabb.go:26-39)
        if (b.mx[a] > n.box.mx[a]) n.box.mx[a] = b.mx[a];
      }
    }
  }

  // node `self` covers order[lo,hi); `draw` = index of the next randomFunc draw (pre-order)
  void build(int32_t self, int lo, int hi, uint64_t draw, int par_depth) {
    BinNode& n = bin[self];
    n.lo = lo; n.hi = hi; n.left = n.right = -1;
    int m = hi - lo;
    if (m == 1) { leaf_box(n); return; }
    int axis = 0;
    if (!rand_zero) {
      uint32_t st = lcg_jump(seed32, draw + 1);
      axis = (int)(3 * ((double)st / 4294967296.0));  // int(3 * randomFunc()) bvh4.go:626
    }
    GoSliceSorter(order.data() + lo, key[axis].data()).sort(m);
    if (m <= 4) { leaf_box(n); return; }
    int mid = m / 2;
    int32_t l = self + 1, r = self + 1 + (int32_t)nodes_of(mid);
    n.left = l; n.right = r;
    uint64_t dl = draw + 1, dr = draw + 1 + draws_of(mid);
    if (par_depth > 0 && m > 32768) {
      std::thread t([=] { build(l, lo, lo + mid, dl, par_depth - 1); });
      build(r, lo + mid, hi, dr, par_depth - 1);
      t.join();
    } else {
      build(l, lo, lo + mid, dl, 0);
      build(r, lo + mid, hi, dr, 0);
    }
    const BoxD &a = bin[l].box, &b = bin[r].box;
    for (int k = 0; k < 3; k++) {
      n.box.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k];
      n.box.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
    }
  }

  // collectChildren (bvh4.go:796-855): expand the first expandable member, append its children
  int collapse(int32_t node, int32_t out[4]) {
    int n = 0;
    out[n++] = bin[node].left;
    out[n++] = bin[node].right;
    for (bool grew = true; grew && n < 4;) {
      grew = false;
      for (int i = 0; i < n; i++) {
        int32_t c = out[i];
        if (bin[c].left < 0) continue;
        if (n - 1 + 2 > 4) continue;
        for (int j = i; j + 1 < n; j++) out[j] = out[j + 1];
        out[n - 1] = bin[c].left;
        out[n] = bin[c].right;
        n++;
        grew = true;
        break;
      }
    }
    return n;
  }

  static void set_slot(izpi_bvh4_node& n, int s, const BoxD& b) {
    n.min_x[s] = ConservativeFloat32Min(b.mn[0]); n.min_y[s] = ConservativeFloat32Min(b.mn[1]);
    n.min_z[s] = ConservativeFloat32Min(b.mn[2]); n.max_x[s] = ConservativeFloat32Max(b.mx[0]);
    n.max_y[s] = ConservativeFloat32Max(b.mx[1]); n.max_z[s] = ConservativeFloat32Max(b.mx[2]);
  }

  int32_t flatten(int32_t node, BVH4Build& out) {  // flattenBVH4 (bvh4.go:714-792)
    int32_t self = (int32_t)out.nodes.size();
    izpi_bvh4_node n;
    for (int i = 0; i < 4; i++) {
      n.child_index[i] = -1; n.primitive_count[i] = 0;
      n.min_x[i] = n.min_y[i] = n.min_z[i] = n.max_x[i] = n.max_y[i] = n.max_z[i] = FLT_MAX;
    }
    const BinNode& b = bin[node];
    if (b.left < 0) {
      n.child_index[0] = (int32_t)out.perm.size();
      n.primitive_count[0] = b.hi - b.lo;
      for (int i = b.lo; i < b.hi; i++) out.perm.push_back(order[i]);
      set_slot(n, 0, b.box);
      out.nodes.push_back(n);
      return self;
    }
    out.nodes.push_back(n);
    int32_t kids[4];
    int nk = collapse(node, kids);
    for (int i = 0; i < nk; i++) {
      int32_t ci = flatten(kids[i], out);
      out.nodes[self].child_index[i] = ci;
      set_slot(out.nodes[self], i, bin[kids[i]].box);
    }
    return self;
  }
};

}  // namespace

BVH4Build NewBVH4(const std::vector<BoxD>& boxes, uint64_t seed, bool rand_zero, int threads) {
  BVH4Build out;
  int n = (int)boxes.size();
  if (n == 0) return out;  // bvh4.go:559-562 logs an error and returns nil
  Builder b(boxes);
  for (int a = 0; a < 3; a++) {
    b.key[a].resize(n);
    for (int i = 0; i < n; i++) b.key[a][i] = boxes[i].mn[a];
  }
  b.order.resize(n);
  for (int i = 0; i < n; i++) b.order[i] = i;
  b.seed32 = (uint32_t)(seed & 0xffffffffull);
  b.rand_zero = rand_zero;
  b.warm(n);
  b.bin.resize(b.count_nodes(n));
  int par_depth = 0;
  for (int t = threads < 1 ? 1 : threads; t > 1; t >>= 1) par_depth++;
  b.build(0, 0, n, 0, par_depth);
  out.nodes.reserve(b.bin.size());
  out.perm.reserve(n);
  b.flatten(0, out);
  return out;
}

}  // namespace izpi
