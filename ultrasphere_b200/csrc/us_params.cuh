// Kernel-parameter blocks.  They are passed BY VALUE as __grid_constant__ parameters, so the
// compiled tree, the pointer tables and the broadcast masks all live in the per-launch constant
// bank: no __constant__ symbol to race on across streams, no H2D copy per call.
#pragma once

#include "us_common.cuh"

namespace us {

struct XformParams {
  us_plan_t plan;                   // pre-order program
  const void* in[US_MAX_LEAVES];    // to: angles (s_nodes order)   from: leaves (c_nodes order)
  void* out[US_MAX_LEAVES];         // to: leaves (c_nodes order)   from: angles (s_nodes order)
  const void* r_in;                 // to only; nullptr => r = 1
  void* r_out;                      // from only; nullptr => not wanted
  uint32_t in_mask[US_MAX_LEAVES];  // bit d set: input spans launch dim d
  const void* tab[US_MAX_LEAVES];   // to, broadcast path: (sin, cos) table of a broadcast angle, or nullptr
  uint32_t r_mask;
  int32_t ndim;
  int64_t n;
  int64_t shape[US_MAX_DIMS];
  // to, linear-functional variant (us_to_cartesian_dot): out[g] = sum_l dot_coef[l] * x_l[g]
  const void* dot_coef;  // n_leaves device values, c_nodes order; nullptr => ordinary to_cartesian
  void* dot_out;         // n elements
};
static_assert(sizeof(XformParams) < 32000, "kernel parameter block must fit the 32 KB limit");

// counts, capacities and the seal of us_plan_compile (us_api.cu)
int check_plan(const us_plan_t* plan, const char* who);

int launch_to_cartesian_strided(const XformParams& p, int dtype, cudaStream_t st);
int launch_from_cartesian_strided(const XformParams& p, int dtype, cudaStream_t st);
int launch_jacobian_strided(const XformParams& p, const int32_t* cos_pow, const int32_t* sin_pow, int dtype, cudaStream_t st);

// broadcast path of to_cartesian (us_transform_bcast.cu): sincos tables for broadcast angles
int64_t bcast_workspace_bytes(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape, const uint32_t* angle_mask);
// with p.dot_coef set: `functionals` launches over the same tables, coefficient rows n_leaves apart, output rows n apart
int try_launch_to_cartesian_bcast(XformParams& p, int dtype, void* workspace, int64_t workspace_bytes, cudaStream_t st, int functionals = 1);

// contiguous fast path (us_transform_tiled.cu); returns 0 if it took the launch, kNotTaken if the
// caller should use the strided kernels, any other value is an error status
constexpr int kNotTaken = -1000;
int try_launch_to_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st);
int try_launch_from_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st);

// wide trees: row-ring kernels (us_transform_ring.cu); *points_per_tile = points each tile covered
template <bool TO>
int launch_ring(const XformParams& p, int dtype, cudaStream_t st, int* points_per_tile);
int ring_points_per_tile(int dtype);  // with the current tuning
// cached cudaFuncSetAttribute(max dynamic smem) + occupancy query
int resident_ctas(const void* fn, size_t smem, int block_threads, int* per_sm);
int sm_count();  // of the current device, cached
int set_tile_tuning(const char* name, int value);

}  // namespace us
