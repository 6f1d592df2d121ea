// Launch-time form of the compiled tree for the TMA-staged kernels (us_transform_tiled.cu,
// us_transform_ring.cu): one 32-byte descriptor per node in program order, holding everything the
// node's step needs already resolved — where its operand rows sit, where its results go, and where
// a deferred branch is parked — so that the kernels decode nothing.
//
// Deferred branches.  In the pre-order program a `c` node defers its sin branch and the `a` node
// that closes its cos subtree resumes it (include/ultrasphere_b200.h).  The staged kernels keep
// the deferred value in shared memory instead of a per-thread stack (which, indexed at run time,
// ptxas can only place in local memory).  Every thread touches only its own 16 bytes of a parking
// row, so no synchronisation is involved.
//   tile kernels  the value is parked in a row of the tile that is dead by then:
//       to_cartesian    `c` parks r*sin(...) in the row of its own angle (consumed by its sincos);
//                       the matching `a` reloads it from there
//       from_cartesian  (walked backwards) the matching `a` parks the finished sin-subtree norm in
//                       the row of its own cos leaf right after loading it; the `c` node reloads it
//   ring kernels  rows are recycled as soon as they are read, so the values go to a small parking
//                 area indexed by the depth of the deferral (plan.max_stack rows)
#pragma once

#include "us_common.cuh"
#include "us_params.cuh"

namespace us {

constexpr uint32_t kNodeKindMask = 3u;
constexpr uint32_t kNodeLinked = 4u;  // to: this `a` resumes a parked product; from: this `a` parks the running norm

struct alignas(16) NodeDesc {
  uint32_t row_a;  // to: the node's angle                 from: cos operand if it is a leaf
  uint32_t row_b;  // to: unused                           from: sin operand if it is a leaf
  uint32_t kind;   // US_NODE_* | kNodeLinked
  uint32_t park;   // `c` nodes and linked `a` nodes: where the deferred value is parked
  void* out_a;     // to: cos-leaf output or nullptr       from: angle output or nullptr
  void* out_b;     // to: sin-leaf output or nullptr       from: unused
};
static_assert(sizeof(NodeDesc) == 32, "descriptor layout");

// Parameter block of the staged kernels, by value in the launch's constant bank.  CAP = 16 for the
// tile kernels (trees of up to 16 rows), US_MAX_NODES for the row-ring kernels.
template <int CAP>
struct WalkParams {
  NodeDesc d[CAP];
  const void* in[CAP + 1];  // staged rows in row order: to: angles (s_nodes order) then r; from: leaves (c_nodes order)
  void* r_out;              // from only; nullptr => not wanted
  int32_t n_nodes;
  int32_t n_rows;           // rows staged per tile
  int32_t has_r;            // to: the last row is r
  int32_t max_park;         // parking rows needed (plan.max_stack); 0: no `c` node
  int32_t zero;             // always 0, but only the host knows: lets a kernel tie an instruction to a loaded register (ring kernels)
  int32_t reserved[3];
};

// Fill the descriptors from the plan and the pointer tables of an XformParams.  `row_scale` turns a
// row index into whatever the kernel addresses rows by (bytes per row for the tile kernels, 1 for
// the ring kernels, whose producer only needs indices); `park_scale` likewise for parking rows of
// the ring kernels (0: park in place, tile kernels).  Returns false if the tree does not fit CAP.
template <int CAP>
inline bool build_walk(const XformParams& p, bool to, uint32_t row_scale, uint32_t park_scale, WalkParams<CAP>& w) {
  const int nn = p.plan.n_nodes;
  const int rows = to ? nn + (p.r_in ? 1 : 0) : p.plan.n_leaves;
  if (nn > CAP || rows > CAP + 1) return false;
  w.n_nodes = nn;
  w.n_rows = rows;
  w.has_r = (to && p.r_in) ? 1 : 0;
  w.r_out = to ? nullptr : p.r_out;
  w.max_park = p.plan.max_stack;
  w.zero = 0;
  w.reserved[0] = w.reserved[1] = w.reserved[2] = 0;
  const int n_in = to ? nn : p.plan.n_leaves;
  for (int i = 0; i < n_in; ++i) w.in[i] = p.in[i];
  if (w.has_r) w.in[nn] = p.r_in;
  int open[US_MAX_STACK];  // `c` nodes whose sin branch is still deferred
  int sp = 0;
  for (int k = 0; k < nn; ++k) {
    const uint64_t word = p.plan.node[k];
    const int kind = node_kind(word);
    NodeDesc& d = w.d[k];
    d.kind = (uint32_t)kind;
    d.park = 0;
    d.row_b = 0;
    d.out_b = nullptr;
    const uint32_t lc = (uint32_t)node_cos_leaf(word), ls = (uint32_t)node_sin_leaf(word);
    if (to) {
      d.row_a = (uint32_t)node_slot(word) * row_scale;
      d.out_a = (kind == US_NODE_A || kind == US_NODE_B) ? p.out[lc] : nullptr;
      d.out_b = (kind == US_NODE_A || kind == US_NODE_BP) ? p.out[ls] : nullptr;
    } else {
      d.row_a = (kind == US_NODE_A || kind == US_NODE_B) ? lc * row_scale : 0;
      d.row_b = (kind == US_NODE_A || kind == US_NODE_BP) ? ls * row_scale : 0;
      d.out_a = p.out[node_slot(word)];
    }
    if (kind == US_NODE_C) {
      if (sp >= US_MAX_STACK) return false;
      if (park_scale) d.park = (uint32_t)sp * park_scale;
      else if (to) d.park = d.row_a;
      open[sp++] = k;
    } else if (kind == US_NODE_A && k != nn - 1) {
      if (sp < 1) return false;  // us_plan_compile never produces this
      NodeDesc& c = w.d[open[--sp]];
      d.kind |= kNodeLinked;
      if (park_scale || to)
        d.park = c.park;
      else
        d.park = c.park = d.row_a;  // from, in place: this node's cos-leaf row, free once it has been read
    }
  }
  return sp == 0;
}

}  // namespace us
