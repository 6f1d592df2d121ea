/*
 * ultrasphere_b200 — C ABI of the B200 (sm_100a) hot path.
 *
 * The reference (ultrasphere 2.0.4) is pure Python and has no FFI of its own; its only seams
 * for this path are the Python methods below.  Each entry point here is what a binding for
 * that method calls once the host side has flattened the dict-of-arrays into pointer tables
 * (INTEGRATION.md shows the ctypes stub a maintainer would add):
 *
 *   us_to_cartesian      replaces the array work of SphericalCoordinates.to_cartesian
 *                        (src/ultrasphere/_coordinates.py:636-656)
 *   us_from_cartesian    replaces the array work of SphericalCoordinates.from_cartesian
 *                        (src/ultrasphere/_coordinates.py:555-579)
 *   us_weighted_reduce   replaces the axis-by-axis vecdot contraction of integrate()'s dense
 *                        branch (src/ultrasphere/_integral.py:261-276)
 *   us_separable_reduce  replaces the per-axis vecdot of integrate()'s separable branch
 *                        (src/ultrasphere/_integral.py:245-260)
 *   us_plan_compile      replaces the per-call graph walk (nx.topological_sort + get_child,
 *                        _coordinates.py:566-574, 638-644) by a one-off flattening of the
 *                        branching-type expression (_coordinates.py:206-242) into a program
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer owned by the
 *     caller; the library never allocates, frees or retains caller-visible memory
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream); no entry point synchronises
 *   - return value: 0 on success, >0 a cudaError_t, <0 one of US_E*; the message is available
 *     from us_last_error_string() (thread-local).  Nothing throws across the boundary.
 *   - there is no CPU fallback: without a usable CUDA device the calls return an error.
 */
#ifndef ULTRASPHERE_B200_H
#define ULTRASPHERE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define US_ABI_VERSION 1

/* exported from libultrasphere_b200.so (everything else in the library is hidden) */
#if defined(__GNUC__)
#define US_API __attribute__((visibility("default")))
#else
#define US_API
#endif

/* capacity of one compiled tree (kernel-parameter resident, "constant memory") */
#define US_MAX_NODES 256 /* internal nodes = angles (s_ndim) */
#define US_MAX_LEAVES (US_MAX_NODES + 1)
#define US_MAX_DIMS 12   /* broadcast dims of the strided path after collapsing */
#define US_MAX_STACK 64  /* pending `c` branches of the pre-order program */
#define US_MAX_AXES 64   /* axes per us_separable_reduce / us_quadrature_rules call */
#define US_MAX_RULE_POINTS 2048 /* Gauss points per axis of us_quadrature_rules (48 KB of coefficients) */

/* element types (the path computes in the type it is given) */
#define US_F32 0
#define US_F64 1

/* library error codes (negative; positive values are cudaError_t) */
#define US_EINVAL (-1)   /* bad argument (null pointer, bad dtype, bad sizes) */
#define US_ECAPACITY (-2) /* tree / dims exceed the US_MAX_* capacities */
#define US_EPLAN (-3)    /* malformed branching expression or plan */

/* node kinds of the pre-order program == reference BranchingType (_coordinates.py:16-26) */
#define US_NODE_A 0  /* cos child leaf, sin child leaf */
#define US_NODE_B 1  /* cos child leaf, sin child internal */
#define US_NODE_BP 2 /* cos child internal, sin child leaf */
#define US_NODE_C 3  /* both internal */

/*
 * A compiled coordinate tree.  node[i] describes the i-th internal node in PRE-ORDER with the
 * cos subtree before the sin subtree — the order of `branching_types_expression`
 * (_coordinates.py:206-242).  In that order a `c` node defers its sin branch on a stack and an
 * `a` node resumes the most recently deferred branch, so a running product (to_cartesian) or a
 * running norm (from_cartesian, walked backwards) needs a stack only at `c`/`a` nodes.
 *
 *   bits  0..1   kind (US_NODE_*)
 *   bits  8..23  angle slot  = position of this node in `s_nodes`  (pointer-table index)
 *   bits 24..39  cos leaf    = position of the cos child in `c_nodes`, 0xFFFF if internal
 *   bits 40..55  sin leaf    = position of the sin child in `c_nodes`, 0xFFFF if internal
 */
typedef struct us_plan {
  int32_t n_nodes;   /* s_ndim */
  int32_t n_leaves;  /* c_ndim = n_nodes + 1 */
  int32_t max_stack; /* deepest deferred-branch stack the program needs */
  int32_t reserved;  /* seal written by us_plan_compile; every entry point refuses a plan whose seal does not match */
  uint64_t node[US_MAX_NODES];
} us_plan_t;

#define US_NO_LEAF 0xFFFFu

US_API int us_abi_version(void);
US_API const char* us_version(void);
US_API const char* us_last_error_string(void);

/*
 * Tuning / measurement hook, never needed in normal use: overrides how the tile kernels cut their
 * shared-memory pool ("tile_groups", "tile_bufs"; 0 = automatic) and how many 16-byte chunks of a
 * row a thread owns ("tile_u": 1 or 2; automatic = 2 for float64 to_cartesian), from which width on trees take
 * the row-ring kernels ("tile_max_rows"), the small-batch cut-over ("tiled_min_tiles", per SM), and
 * the ring kernels' geometry: 16-byte chunks of a row per thread ("ring_u": 1 or 2), CTAs per SM
 * their slots are sized for ("ring_ctas") and a cap on the row slots ("ring_slots"); 0 = automatic.
 * The same knobs are read from US_TILE_GROUPS, US_TILE_BUFS, US_TILE_U, US_TILE_MAX_ROWS, US_TILED_MIN_TILES,
 * US_RING_U, US_RING_CTAS, US_RING_SLOTS at first use.  Process-wide; results never depend on it,
 * only speed.
 */
US_API int us_set_tuning(const char* name, int value);

/* Number of CUDA devices visible (>= 0), or minus the cudaError_t that the query raised. */
US_API int us_device_count(void);

/*
 * Compile a branching-type expression.  kinds[i] in pre-order; angle_slot[i] = index of node i
 * in s_nodes; leaf_slot[j] = index in c_nodes of the j-th leaf OPERAND in program order (node by
 * node, a node's cos operand before its sin operand: `a` consumes two entries, `b` and `b'` one,
 * `c` none).  Validates that the expression closes exactly (what _creation.py:66-70 raises
 * ValueError for) and that slots are permutations.
 */
US_API int us_plan_compile(int32_t n_nodes, const uint8_t* kinds, const int32_t* angle_slot,
                    const int32_t* leaf_slot, us_plan_t* out_plan);

/*
 * x_leaf = r * prod over the root→leaf path of cos(theta) or sin(theta)   (multiply order
 * root→leaf, one rounding per product — _coordinates.py:645-648).
 *
 * The launch covers `n` points laid out as a C-contiguous index space shape[0..ndim).
 * Input k contributes the element at offset sum_d idx_d * stride_k,d where stride is the
 * dense stride of the sub-shape selected by mask bit d (bit d set = the input spans dim d,
 * clear = broadcast along d).  ndim==1 with all masks 1 is the contiguous fast path.
 *
 * angle_dev[s]  : n_nodes pointers, s_nodes order         r_dev : may be NULL (r = 1)
 * out_dev[l]    : n_leaves pointers, c_nodes order, each n elements; NULL entries are skipped
 */
US_API int us_to_cartesian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                    const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                    uint32_t r_mask, void* const* out_dev, void* stream);

/*
 * Same as us_to_cartesian, with scratch for the broadcast case (an integrand inside integrate()
 * passes s_ndim one-dimensional axes and asks for the full tensor grid): every angle operand
 * that is at least 4x smaller than the launch gets its sin/cos evaluated ONCE into
 * workspace_dev by a pre-pass, and the grid kernel then only looks them up, multiplies and
 * writes.  Values are identical to us_to_cartesian (same sincos on the same angle values).
 * us_to_cartesian_workspace_bytes returns the scratch needed (0: nothing to tabulate);
 * workspace_dev must be 16-byte aligned; NULL / too small falls back to us_to_cartesian.
 */
US_API int64_t us_to_cartesian_workspace_bytes(const us_plan_t* plan, int dtype, int64_t n, int ndim,
                                               const int64_t* shape, const uint32_t* angle_mask);
US_API int us_to_cartesian_ws(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                              const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                              uint32_t r_mask, void* const* out_dev, void* workspace_dev,
                              int64_t workspace_bytes, void* stream);

/*
 * r = sqrt(sum over leaves in c_nodes order of x^2)  (sequential, each square and each add
 * rounded once — bit-identical to numpy.linalg.vector_norm(axis=0), _coordinates.py:556-563);
 * per internal node, children first: theta = atan2(v_sin, v_cos), v = sqrt(v_sin^2 + v_cos^2)
 * (_coordinates.py:575-578).
 *
 * x_dev[l]         : n_leaves pointers, c_nodes order (+ masks as above)
 * r_out_dev        : n elements or NULL
 * angle_out_dev[s] : n_nodes pointers, s_nodes order, each n elements; NULL entries skipped
 */
US_API int us_from_cartesian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                      const void* const* x_dev, const uint32_t* x_mask, void* r_out_dev,
                      void* const* angle_out_dev, void* stream);

/*
 * Row-matrix forms of the two transforms for the common dense case: the leaves are the rows of one
 * C-contiguous (c_ndim, n) matrix x_dev (row pitch x_row_bytes), the results the rows of out_dev —
 * from: row 0 = r, row 1 + s = angle s (s_nodes order); to: row l = leaf l (c_nodes order).  Same
 * kernels and results as us_from_cartesian / us_to_cartesian; a binding calls these to avoid
 * building pointer and mask tables per call (what SphericalCoordinates.from_cartesian(array) and
 * to_cartesian(..., as_array=True) need, _coordinates.py:555-563, 650-654).
 */
US_API int us_from_cartesian_rows(const us_plan_t* plan, int dtype, int64_t n, const void* x_dev, int64_t x_row_bytes,
                                  void* out_dev, int64_t out_row_bytes, void* stream);
US_API int us_to_cartesian_rows(const us_plan_t* plan, int dtype, int64_t n, const void* const* angle_dev,
                                const void* r_dev, void* out_dev, int64_t out_row_bytes, void* stream);

/*
 * out[m] (+)= sum over the grid of  prod_i w_i[idx_i] * val[idx_0, ..., idx_{ndim-1}, m]
 * val is C-contiguous with shape (extent[0..ndim), m); w_dev[i] has extent[i] entries (an axis
 * the integrand did not depend on arrives with extent 1 and w = sum of that axis' weights,
 * _integral.py:272-273).  Complex values are passed as m*2 reals (weights are real).
 * Accumulation is in float64 whatever `dtype` is; the cross-block order is fixed (deterministic).
 * scratch_dev: at least us_weighted_reduce_scratch_bytes(m) bytes, 16-byte aligned.
 */
US_API int64_t us_weighted_reduce_scratch_bytes(int64_t m);
US_API int us_weighted_reduce(int dtype, int ndim, const int64_t* extent, const void* const* w_dev,
                       const void* val_dev, int64_t m, void* out_dev, int accumulate,
                       void* scratch_dev, int64_t scratch_bytes, void* stream);

/*
 * For every axis a < n_axes:  out_a[m] = sum_j w_a[j] * val_a[j, m]   (val_a: (len[a], m[a])
 * C-contiguous).  One launch for all axes.  _integral.py:245-260.
 */
US_API int us_separable_reduce(int dtype, int n_axes, const int64_t* len, const int64_t* m,
                        const void* const* w_dev, const void* const* val_dev,
                        void* const* out_dev, void* stream);

/*
 * ---- the integrand's side of integrate(): a linear functional of the Cartesian coordinates ----
 * out[j][g] = sum over leaves l of coef[j][l] * x_l[g],  x = to_cartesian(spherical)   (j < m)
 * — what the reference's own quadrature test computes with
 * `einsum("v,v...->...", k, c.to_cartesian(s, as_array=True))` (tests/test_integral.py:88-92) on the
 * tensor grid of quadrature angles.  The leaf products are folded into the sum as they are formed
 * (program order, one FMA per leaf) and never written: sizeof(T) instead of c_ndim*sizeof(T) bytes
 * per grid point and no second pass.  Operands are addressed exactly as in us_to_cartesian_ws
 * (launch shape + masks, optional sincos-table workspace of us_to_cartesian_workspace_bytes()).
 * coef_dev : m * n_leaves device values of the compute dtype, row-major, c_nodes order
 * out_dev  : m * n elements, row j at out + j*n
 */
US_API int us_to_cartesian_dot(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                               const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                               uint32_t r_mask, const void* coef_dev, int m, void* out_dev,
                               void* workspace_dev, int64_t workspace_bytes, void* stream);

/*
 * ---- backward passes, for torch.autograd ----
 * The reference is differentiable because it is made of array-API calls
 * (src/ultrasphere/_coordinates.py:555-579, 636-656); these two entry points give the CUDA path the
 * same property.  Dense operands only: every array has n elements (the host broadcasts before the
 * call and autograd sums afterwards).  NULL gradient inputs mean zero.
 *
 * us_to_cartesian_backward:  given dL/dx_l (c_nodes order) returns dL/dtheta_s (s_nodes order, all
 *   n_nodes buffers required) and, if grad_r_dev is not NULL, dL/dr.  r_dev NULL = radius 1.
 * us_from_cartesian_backward: given dL/dr and dL/dtheta_s returns dL/dx_l (c_nodes order).
 *   scratch_dev: 2 * n_nodes * n elements of the compute dtype.  Points whose subtree norm is zero
 *   (no direction) get a zero gradient.
 */
US_API int us_to_cartesian_backward(const us_plan_t* plan, int dtype, int64_t n, const void* const* angle_dev,
                                    const void* r_dev, const void* const* grad_leaf_dev,
                                    void* const* grad_angle_dev, void* grad_r_dev, void* stream);
US_API int us_from_cartesian_backward(const us_plan_t* plan, int dtype, int64_t n, const void* const* x_dev,
                                      const void* grad_r_dev, const void* const* grad_angle_dev,
                                      void* const* grad_x_dev, void* scratch_dev, void* stream);

/*
 * ---- "next" row: SphericalCoordinates.jacobian (_coordinates.py:668-702) ----
 * out = r * prod over internal nodes of cos(theta)^cos_pow * sin(theta)^sin_pow, one fused pass that
 * shares the sincos with nothing else to write: (s_ndim + 2) * sizeof(T) bytes per point.  The
 * exponents are S[child] + 1 for an internal child and 0 for a leaf (the reference reads NaN for
 * leaves and therefore returns NaN for every tree; this is the formula it intends).  Operands are
 * addressed exactly as in us_to_cartesian (launch shape + per-operand masks); r_dev may be NULL.
 * cos_pow / sin_pow: n_nodes entries in s_nodes order.
 */
US_API int us_jacobian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                       const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                       uint32_t r_mask, const int32_t* cos_pow, const int32_t* sin_pow, void* out_dev,
                       void* stream);

/*
 * ---- "next" row of the path: the producer of transform inputs ----
 * Uniform points in the unit ball (surface = 0) or on the unit sphere (surface = 1) of dimension
 * `dim`, written as `dim` arrays of n elements (reference: src/ultrasphere/_random.py:16-57, which
 * samples with NumPy on the host and copies to the device).  Philox4x32-10 keyed by `seed`; the
 * value of point i depends only on (seed, i).  Distribution parity, not bit parity: the reference's
 * stream is NumPy's PCG64.
 */
US_API int us_random_ball(int dtype, int dim, int64_t n, uint64_t seed, int surface, void* const* out_dev,
                          void* stream);

/*
 * out_k[i] ~ U[lo_k, hi_k) for k < count (per-angle uniform sampling, _random.py:170-189).
 */
US_API int us_random_uniform(int dtype, int count, int64_t n, uint64_t seed, const double* lo,
                             const double* hi, void* const* out_dev, void* stream);

/*
 * ---- "next" row: the 1-D rules of roots() built on the device ----
 * One launch fills the angle and weight tables of n_axes angles (reference: the per-node loop of
 * roots(), src/ultrasphere/_integral.py:83-109, whose Gauss rules come from the host call
 * scipy.special.roots_jacobi followed by one H2D copy per table).  Per axis i:
 *   kind US_NODE_A : count[i] = 2n points  x_j = j*pi/n, w_j = pi/n   (_integral.py:85-87), rounded
 *                    in the output type exactly as torch's CUDA kernels round `arange(2n)*pi/n`
 *   kind US_NODE_B : count[i] Gauss-Jacobi(alpha, beta) points t_j, x_j = acos(t_j)     (:88-92)
 *   kind US_NODE_BP: x_j = asin(t_j)                                                    (:93-97)
 *   kind US_NODE_C : x_j = acos(t_j)/2, w_j divided by 2^(alpha+beta+2)                 (:98-105)
 * (alpha, beta) are passed exactly as the reference passes them to roots_jacobi.  The Gauss nodes
 * are the eigenvalues of the same Jacobi matrix scipy diagonalises, found per root by Sturm
 * multisection + Newton; weights are Christoffel sums.  Agreement with roots_jacobi: nodes
 * <= 2.3e-16 absolute, weights within scipy's own error (both checked against 50-digit arithmetic
 * in tests/test_gpu_rules.py).  Tables ascend in t like scipy's.
 * dtype[i] is US_F32 / US_F64 per axis (the reference's dtype=None gives float32 trapezoid tables
 * next to float64 Gauss tables); x_dev[i], w_dev[i]: count[i] elements each.
 */
US_API int us_quadrature_rules(int n_axes, const uint8_t* kind, const int64_t* count, const double* alpha,
                               const double* beta, const int32_t* dtype, void* const* x_dev,
                               void* const* w_dev, void* stream);

/*
 * Measurement helper: runs a dependent-free DFMA (dtype=US_F64) or FFMA (US_F32) loop on every
 * SM and returns the achieved TFLOP/s in *tflops (this one call synchronises).  Used by bench.py
 * as the denominator for the FP64-pipe-bound kernels (MEASURED_PEAKS.json has no FP64 entry).
 */
US_API int us_probe_fma_peak(int dtype, int iters, double* tflops);

/*
 * Test hook: evaluates the library's own float32 device math on n elements so that
 * tests/test_gpu_math.py can bound its error against float64 and check the IEEE claims.
 *   fn 0: out0 = sin(a), out1 = cos(a)          (the sincos used by us_to_cartesian)
 *   fn 1: out0 = atan2(a, b)                    (the atan2 used by us_from_cartesian)
 *   fn 2: out0 = library sqrt(a), out1 = __fsqrt_rn(a)   (must be bit-identical)
 * All pointers are device pointers to n floats; out1 may be NULL for fn 1.
 */
US_API int us_math_probe_f32(int fn, int64_t n, const float* a_dev, const float* b_dev, float* out0_dev,
                             float* out1_dev, void* stream);

/* float64 twin: fn 0: out0 = sin(a), out1 = cos(a); fn 1: out0 = atan2(a, b), out1 = sqrt(a*a + b*b);
 * fn 2: out0 = library sqrt(a), out1 = __dsqrt_rn(a). */
US_API int us_math_probe_f64(int fn, int64_t n, const double* a_dev, const double* b_dev, double* out0_dev,
                             double* out1_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ULTRASPHERE_B200_H */
