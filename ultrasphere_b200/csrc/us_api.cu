// C ABI entry points (include/ultrasphere_b200.h): argument checking, plan compilation,
// dispatch between the tiled fast path and the strided path, error reporting.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "us_common.cuh"
#include "us_params.cuh"

namespace us {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

// us_plan_t is a plain public struct and the kernels index their pointer tables with the fields of
// its node words without range checks, so a plan is only accepted with the seal us_plan_compile put
// on it (FNV-1a over the counts and the words; never 0): a hand-built or corrupted plan is refused
// with US_EPLAN instead of reading or writing out of bounds on the device.
static int32_t plan_seal(const us_plan_t* plan) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](uint64_t v) {
    for (int b = 0; b < 8; ++b) {
      h ^= (v >> (8 * b)) & 0xFFu;
      h *= 1099511628211ull;
    }
  };
  mix((uint64_t)(uint32_t)plan->n_nodes | ((uint64_t)(uint32_t)plan->n_leaves << 32));
  mix((uint64_t)(uint32_t)plan->max_stack);
  for (int i = 0; i < plan->n_nodes; ++i) mix(plan->node[i]);
  const int32_t s = (int32_t)(h ^ (h >> 32));
  return s ? s : 1;
}

int check_plan(const us_plan_t* plan, const char* who) {
  if (!plan) return fail(US_EINVAL, "%s: null plan", who);
  if (plan->n_nodes < 1 || plan->n_nodes > US_MAX_NODES || plan->n_leaves != plan->n_nodes + 1)
    return fail(US_EPLAN, "%s: plan has %d nodes / %d leaves", who, plan->n_nodes, plan->n_leaves);
  if (plan->max_stack < 0 || plan->max_stack > US_MAX_STACK)
    return fail(US_ECAPACITY, "%s: plan needs a stack of %d > %d", who, plan->max_stack, US_MAX_STACK);
  if (plan->reserved != plan_seal(plan))
    return fail(US_EPLAN, "%s: plan was not produced by us_plan_compile (or was modified afterwards)", who);
  return 0;
}

static int fill_common(XformParams& p, const us_plan_t* plan, int dtype, int64_t n, int ndim,
                       const int64_t* shape, const char* who) {
  if (int rc = check_plan(plan, who)) return rc;
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "%s: bad dtype %d", who, dtype);
  if (n < 0) return fail(US_EINVAL, "%s: n < 0", who);
  if (ndim < 1 || ndim > US_MAX_DIMS) return fail(US_ECAPACITY, "%s: ndim %d not in 1..%d", who, ndim, US_MAX_DIMS);
  if (!shape) return fail(US_EINVAL, "%s: null shape", who);
  int64_t prod = 1;
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] < 0) return fail(US_EINVAL, "%s: negative extent", who);
    p.shape[d] = shape[d];
    prod *= shape[d];
  }
  if (prod != n) return fail(US_EINVAL, "%s: shape product %lld != n %lld", who, (long long)prod, (long long)n);
  memcpy(&p.plan, plan, sizeof(us_plan_t));
  p.ndim = ndim;
  p.n = n;
  return 0;
}

}  // namespace us

using namespace us;

extern "C" int us_abi_version(void) { return US_ABI_VERSION; }
extern "C" const char* us_version(void) { return "ultrasphere_b200 0.1.0 (sm_100a)"; }
extern "C" const char* us_last_error_string(void) { return g_err; }

extern "C" int us_set_tuning(const char* name, int value) {
  if (!name) return fail(US_EINVAL, "us_set_tuning: null name");
  return set_tile_tuning(name, value);
}

extern "C" int us_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return -cuda_fail(e, "cudaGetDeviceCount");
  return n;
}

extern "C" int us_plan_compile(int32_t n_nodes, const uint8_t* kinds, const int32_t* angle_slot,
                               const int32_t* leaf_slot, us_plan_t* out) {
  if (!out) return fail(US_EINVAL, "us_plan_compile: null output");
  if (n_nodes < 1) return fail(US_EPLAN, "us_plan_compile: a tree needs at least one internal node");
  if (n_nodes > US_MAX_NODES)
    return fail(US_ECAPACITY, "us_plan_compile: %d internal nodes exceed the capacity of %d", n_nodes, US_MAX_NODES);
  if (!kinds || !angle_slot || !leaf_slot) return fail(US_EINVAL, "us_plan_compile: null table");
  memset(out, 0, sizeof *out);
  bool seen_slot[US_MAX_NODES] = {false};
  bool seen_leaf[US_MAX_LEAVES] = {false};
  int open_branches = 1;  // subtrees still to be written, the root first
  int sp = 0, max_sp = 0, li = 0;
  const int n_leaves = n_nodes + 1;
  for (int i = 0; i < n_nodes; ++i) {
    const int k = kinds[i];
    if (k < 0 || k > 3) return fail(US_EPLAN, "us_plan_compile: node %d has invalid kind %d", i, k);
    if (open_branches < 1) return fail(US_EPLAN, "us_plan_compile: expression closes before node %d", i);
    const int s = angle_slot[i];
    if (s < 0 || s >= n_nodes || seen_slot[s]) return fail(US_EPLAN, "us_plan_compile: bad or repeated angle slot %d", s);
    seen_slot[s] = true;
    uint32_t lc = US_NO_LEAF, ls = US_NO_LEAF;
    if (k == US_NODE_A || k == US_NODE_B) {
      if (li >= n_leaves) return fail(US_EPLAN, "us_plan_compile: too many leaf operands");
      lc = (uint32_t)leaf_slot[li++];
    }
    if (k == US_NODE_A || k == US_NODE_BP) {
      if (li >= n_leaves) return fail(US_EPLAN, "us_plan_compile: too many leaf operands");
      ls = (uint32_t)leaf_slot[li++];
    }
    for (uint32_t l : {lc, ls}) {
      if (l == US_NO_LEAF) continue;
      if (l >= (uint32_t)n_leaves || seen_leaf[l]) return fail(US_EPLAN, "us_plan_compile: bad or repeated leaf slot %u", l);
      seen_leaf[l] = true;
    }
    open_branches += (k == US_NODE_C) ? 1 : (k == US_NODE_A ? -1 : 0);
    if (k == US_NODE_C) {
      if (++sp > max_sp) max_sp = sp;
    } else if (k == US_NODE_A && i != n_nodes - 1) {
      if (sp < 1) return fail(US_EPLAN, "us_plan_compile: `a` at node %d has no deferred branch to resume", i);
      --sp;
    }
    out->node[i] = (uint64_t)k | ((uint64_t)(uint32_t)s << 8) | ((uint64_t)lc << 24) | ((uint64_t)ls << 40);
  }
  if (open_branches != 0 || sp != 0 || li != n_leaves)
    return fail(US_EPLAN, "us_plan_compile: expression does not close (open=%d, stack=%d, leaves=%d/%d)",
                open_branches, sp, li, n_leaves);
  if (max_sp > US_MAX_STACK)
    return fail(US_ECAPACITY, "us_plan_compile: deferred-branch depth %d exceeds %d", max_sp, US_MAX_STACK);
  out->n_nodes = n_nodes;
  out->n_leaves = n_leaves;
  out->max_stack = max_sp;
  out->reserved = plan_seal(out);
  return 0;
}

extern "C" int us_to_cartesian_ws(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                                  const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                                  uint32_t r_mask, void* const* out_dev, void* workspace_dev,
                                  int64_t workspace_bytes, void* stream) {
  static const char* who = "us_to_cartesian";
  XformParams p;
  memset(&p, 0, sizeof p);
  if (int rc = fill_common(p, plan, dtype, n, ndim, shape, who)) return rc;
  if (!angle_dev || !angle_mask || !out_dev) return fail(US_EINVAL, "%s: null pointer table", who);
  for (int s = 0; s < plan->n_nodes; ++s) {
    if (!angle_dev[s] && n > 0) return fail(US_EINVAL, "%s: angle %d is null", who, s);
    p.in[s] = angle_dev[s];
    p.in_mask[s] = angle_mask[s];
  }
  for (int l = 0; l < plan->n_leaves; ++l) p.out[l] = out_dev[l];
  p.r_in = r_dev;
  p.r_mask = r_mask;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = try_launch_to_cartesian_tiled(p, dtype, st);
  if (rc != kNotTaken) return rc;
  rc = try_launch_to_cartesian_bcast(p, dtype, workspace_dev, workspace_bytes, st);
  if (rc != kNotTaken) return rc;
  return launch_to_cartesian_strided(p, dtype, st);
}

extern "C" int us_to_cartesian_dot(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                                   const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                                   uint32_t r_mask, const void* coef_dev, int m, void* out_dev, void* workspace_dev,
                                   int64_t workspace_bytes, void* stream) {
  static const char* who = "us_to_cartesian_dot";
  XformParams p;
  memset(&p, 0, sizeof p);
  if (int rc = fill_common(p, plan, dtype, n, ndim, shape, who)) return rc;
  if (!angle_dev || !angle_mask) return fail(US_EINVAL, "%s: null pointer table", who);
  if (m < 0) return fail(US_EINVAL, "%s: m < 0", who);
  if (m == 0 || n == 0) return 0;
  if (!coef_dev || !out_dev) return fail(US_EINVAL, "%s: null coefficient or output pointer", who);
  for (int s = 0; s < plan->n_nodes; ++s) {
    if (!angle_dev[s]) return fail(US_EINVAL, "%s: angle %d is null", who, s);
    p.in[s] = angle_dev[s];
    p.in_mask[s] = angle_mask[s];
  }
  p.r_in = r_dev;
  p.r_mask = r_mask;
  p.dot_coef = coef_dev;
  p.dot_out = out_dev;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = try_launch_to_cartesian_bcast(p, dtype, workspace_dev, workspace_bytes, st, m);
  if (rc != kNotTaken) return rc;
  const size_t esz = dtype == US_F32 ? 4 : 8;
  for (int j = 0; j < m; ++j) {
    p.dot_coef = (const char*)coef_dev + (size_t)j * plan->n_leaves * esz;
    p.dot_out = (char*)out_dev + (size_t)j * n * esz;
    if ((rc = launch_to_cartesian_strided(p, dtype, st)) != 0) return rc;
  }
  return 0;
}

extern "C" int us_to_cartesian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                               const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                               uint32_t r_mask, void* const* out_dev, void* stream) {
  return us_to_cartesian_ws(plan, dtype, n, ndim, shape, angle_dev, angle_mask, r_dev, r_mask, out_dev, nullptr, 0, stream);
}

extern "C" int64_t us_to_cartesian_workspace_bytes(const us_plan_t* plan, int dtype, int64_t n, int ndim,
                                                   const int64_t* shape, const uint32_t* angle_mask) {
  if (!plan || !shape || !angle_mask || ndim < 1 || ndim > US_MAX_DIMS || (dtype != US_F32 && dtype != US_F64)) return 0;
  return bcast_workspace_bytes(plan, dtype, n, ndim, shape, angle_mask);
}

extern "C" int us_jacobian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                           const void* const* angle_dev, const uint32_t* angle_mask, const void* r_dev,
                           uint32_t r_mask, const int32_t* cos_pow, const int32_t* sin_pow, void* out_dev,
                           void* stream) {
  static const char* who = "us_jacobian";
  XformParams p;
  memset(&p, 0, sizeof p);
  if (int rc = fill_common(p, plan, dtype, n, ndim, shape, who)) return rc;
  if (!angle_dev || !angle_mask || !cos_pow || !sin_pow || (!out_dev && n > 0)) return fail(US_EINVAL, "%s: null pointer", who);
  for (int s = 0; s < plan->n_nodes; ++s) {
    if (!angle_dev[s] && n > 0) return fail(US_EINVAL, "%s: angle %d is null", who, s);
    p.in[s] = angle_dev[s];
    p.in_mask[s] = angle_mask[s];
  }
  p.out[0] = out_dev;
  p.r_in = r_dev;
  p.r_mask = r_mask;
  if (n == 0) return 0;
  return launch_jacobian_strided(p, cos_pow, sin_pow, dtype, (cudaStream_t)stream);
}

extern "C" int us_from_cartesian(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                                 const void* const* x_dev, const uint32_t* x_mask, void* r_out_dev,
                                 void* const* angle_out_dev, void* stream) {
  static const char* who = "us_from_cartesian";
  XformParams p;
  memset(&p, 0, sizeof p);
  if (int rc = fill_common(p, plan, dtype, n, ndim, shape, who)) return rc;
  if (!x_dev || !x_mask || !angle_out_dev) return fail(US_EINVAL, "%s: null pointer table", who);
  for (int l = 0; l < plan->n_leaves; ++l) {
    if (!x_dev[l] && n > 0) return fail(US_EINVAL, "%s: leaf %d is null", who, l);
    p.in[l] = x_dev[l];
    p.in_mask[l] = x_mask[l];
  }
  for (int s = 0; s < plan->n_nodes; ++s) p.out[s] = angle_out_dev[s];
  p.r_out = r_out_dev;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = try_launch_from_cartesian_tiled(p, dtype, st);
  if (rc != kNotTaken) return rc;
  return launch_from_cartesian_strided(p, dtype, st);
}

// ---- row-matrix forms: the common dense case with the least work on the caller's side -----------
// Every operand is a row of one C-contiguous (rows, n) matrix, so the caller passes a base pointer
// and a row pitch instead of building pointer and mask tables (a binding's per-call cost is what
// bounds small batches: a 10^6-point float64 create_spherical() launch takes ~7 us).
extern "C" int us_from_cartesian_rows(const us_plan_t* plan, int dtype, int64_t n, const void* x_dev, int64_t x_row_bytes,
                                      void* out_dev, int64_t out_row_bytes, void* stream) {
  static const char* who = "us_from_cartesian_rows";
  if (int rc = check_plan(plan, who)) return rc;
  if (n > 0 && (!x_dev || !out_dev)) return fail(US_EINVAL, "%s: null matrix", who);
  const void* x[US_MAX_LEAVES];
  uint32_t mask[US_MAX_LEAVES];
  void* out[US_MAX_NODES];
  for (int l = 0; l < plan->n_leaves; ++l) {
    x[l] = (const char*)x_dev + (int64_t)l * x_row_bytes;
    mask[l] = 1u;
  }
  for (int s = 0; s < plan->n_nodes; ++s) out[s] = (char*)out_dev + (int64_t)(1 + s) * out_row_bytes;
  return us_from_cartesian(plan, dtype, n, 1, &n, x, mask, out_dev, out, stream);
}

extern "C" int us_to_cartesian_rows(const us_plan_t* plan, int dtype, int64_t n, const void* const* angle_dev,
                                    const void* r_dev, void* out_dev, int64_t out_row_bytes, void* stream) {
  static const char* who = "us_to_cartesian_rows";
  if (int rc = check_plan(plan, who)) return rc;
  if (!angle_dev || (n > 0 && !out_dev)) return fail(US_EINVAL, "%s: null pointer", who);
  uint32_t mask[US_MAX_NODES];
  void* out[US_MAX_LEAVES];
  for (int s = 0; s < plan->n_nodes; ++s) mask[s] = 1u;
  for (int l = 0; l < plan->n_leaves; ++l) out[l] = (char*)out_dev + (int64_t)l * out_row_bytes;
  return us_to_cartesian_ws(plan, dtype, n, 1, &n, angle_dev, mask, r_dev, 1u, out, nullptr, 0, stream);
}

// ---- FMA peak probe ----------------------------------------------------------------------
namespace us {
template <typename T>
__global__ void __launch_bounds__(256) fma_probe(T* out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == (T)123456789) out[0] = s;  // never true; keeps the loop alive
}
}  // namespace us

extern "C" int us_probe_fma_peak(int dtype, int iters, double* tflops) {
  if (!tflops || iters < 1) return fail(US_EINVAL, "us_probe_fma_peak: bad arguments");
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "us_probe_fma_peak: bad dtype");
  int dev = 0, sms = 0;
  US_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  US_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
  void* buf = nullptr;
  US_CUDA_TRY(cudaMalloc(&buf, 64), "cudaMalloc");  // measurement helper only: owns its 64 bytes
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    if (dtype == US_F64)
      fma_probe<double><<<blocks, threads>>>((double*)buf, iters, 0.999999, 1e-9);
    else
      fma_probe<float><<<blocks, threads>>>((float*)buf, iters, 0.999999f, 1e-9f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  US_CUDA_TRY(cudaGetLastError(), "fma_probe");
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  return 0;
}

// ---- math test hook ------------------------------------------------------------------------
namespace us {
__global__ void __launch_bounds__(256) math_probe_f32(int fn, int64_t n, const float* a, const float* b, float* o0, float* o1) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float x[4], y[4], r0[4], r1[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int64_t j = (i + v < n) ? i + v : n - 1;
    x[v] = a[j];
    y[v] = b ? b[j] : 0.0f;
  }
  if (fn == 0) {
    sincos_vec<4>(x, r0, r1);
  } else if (fn == 1) {
    atan2_vec<4>(x, y, r0);
#pragma unroll
    for (int v = 0; v < 4; ++v) r1[v] = 0.0f;
  } else {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      r0[v] = x[v];
      r1[v] = __fsqrt_rn(x[v]);
    }
    sqrt_vec<4>(r0);
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    if (i + v < n) {
      o0[i + v] = r0[v];
      if (o1) o1[i + v] = r1[v];
    }
  }
}
}  // namespace us

extern "C" int us_math_probe_f32(int fn, int64_t n, const float* a_dev, const float* b_dev, float* out0_dev,
                                 float* out1_dev, void* stream) {
  if (fn < 0 || fn > 2 || n < 0 || !a_dev || !out0_dev || (fn == 1 && !b_dev))
    return fail(US_EINVAL, "us_math_probe_f32: bad arguments");
  if (n == 0) return 0;
  const int64_t blocks = (n + 1023) / 1024;
  math_probe_f32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(fn, n, a_dev, b_dev, out0_dev, out1_dev);
  US_CUDA_TRY(cudaGetLastError(), "math_probe_f32 launch");
  return 0;
}

namespace us {
__global__ void __launch_bounds__(256) math_probe_f64(int fn, int64_t n, const double* a, const double* b, double* o0, double* o1) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  double x[2], y[2], r0[2], r1[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int64_t j = (i + v < n) ? i + v : n - 1;
    x[v] = a[j];
    y[v] = b ? b[j] : 0.0;
  }
  if (fn == 0) {
    sincos_vec<2>(x, r0, r1);
  } else if (fn == 1) {
    // the kernels' node op: atan2 of (sin operand, cos operand) and their norm
    node_from_vec<2>(x, y, r0, r1);
  } else {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const bool ok = x[v] > 1e-300 && x[v] < 1e300;
      r0[v] = ok ? sqrt_fast(x[v]) : __dsqrt_rn(x[v]);
      r1[v] = __dsqrt_rn(x[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    if (i + v < n) {
      o0[i + v] = r0[v];
      if (o1) o1[i + v] = r1[v];
    }
  }
}
}  // namespace us

extern "C" int us_math_probe_f64(int fn, int64_t n, const double* a_dev, const double* b_dev, double* out0_dev,
                                 double* out1_dev, void* stream) {
  if (fn < 0 || fn > 2 || n < 0 || !a_dev || !out0_dev || (fn == 1 && !b_dev))
    return fail(US_EINVAL, "us_math_probe_f64: bad arguments");
  if (n == 0) return 0;
  const int64_t blocks = (n + 511) / 512;
  math_probe_f64<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(fn, n, a_dev, b_dev, out0_dev, out1_dev);
  US_CUDA_TRY(cudaGetLastError(), "math_probe_f64 launch");
  return 0;
}
