// bvh4_builder.hpp -- host-side stand-in for hitable.NewBVH4 (internal/hitable/bvh4.go:517-855).
//
// In a Go deployment the BVH4 is built by the reference's own Go code and its exported
// `Nodes` / `Primitives` fields (bvh4.go:42-47) are handed to izpi_scene_upload.  This image has
// no Go toolchain, so the host runtime carries a C++ builder that produces the SAME node array:
// random-axis median split on box.min[axis] (aabb.go:42-54) ordered like Go's sort.Slice, leaves of
// <= 4 primitives, binary -> 4-ary collapse, pre-order numbering, outward fp32 rounding.
//
// Unlike the reference it sorts index ranges in place and builds subtrees in parallel: the number
// of randomFunc draws in a subtree depends only on its size, so every node's LCG state is reached
// by an O(log k) jump-ahead and the result is independent of the thread count.
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/izpi_scene.h"

namespace izpi {

struct BoxD {
  double mn[3], mx[3];
};

struct BVH4Build {
  std::vector<izpi_bvh4_node> nodes;  // BVH4.Nodes
  std::vector<int32_t> perm;          // Primitives[i] = hitables[perm[i]]
};

// boxes[i] = BoundingBox(time0, time1) of hitable i.  seed/rand_zero: the injected randomFunc
// (fastrandom LCG constants, fastrandom.go:7-11; or the tests' constant 0, bvh4_test.go:57).
BVH4Build NewBVH4(const std::vector<BoxD>& boxes, uint64_t seed, bool rand_zero, int threads);

// Optional device build (csrc/device/bvh_build.cu): Morton-ordered LBVH collapsed into the same node format.  The tree
// differs from the reference's; closest hits do not.  Empty result + izpi_last_error() on failure.
BVH4Build BuildBVH4Device(const std::vector<BoxD>& boxes);

float ConservativeFloat32Min(double v);  // bvh4.go:494
float ConservativeFloat32Max(double v);  // bvh4.go:506

}  // namespace izpi
