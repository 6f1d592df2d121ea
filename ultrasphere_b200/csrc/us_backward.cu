// Backward passes of the two transforms, for torch.autograd (the reference is differentiable for
// free because it is made of array-API calls, src/ultrasphere/_coordinates.py:555-579, 636-656).
// Gradients are not the hot path: dense equal-shape operands only (the host broadcasts before and
// lets autograd sum afterwards), one thread per point, two walks over the pre-order program.
//
// to_cartesian, x_l = r * prod(ancestor factors).  With P_k the running product entering node k
// (top-down walk) and, bottom-up, G_T = sum over the leaves of subtree T of g_l * (x_l / P_T):
//   dL/dtheta_k = P_k * (cos_k * G_sin - sin_k * G_cos),   G_k = cos_k * G_cos + sin_k * G_sin,
//   dL/dr = G_root.   P_k is parked in the grad_theta_k output between the two walks.
//
// from_cartesian, theta_k = atan2(v_sin, v_cos), v_k = sqrt(v_sin^2 + v_cos^2), r = v_root.  With
// D_k = dL/dv_k flowing top-down (D_root = g_r):
//   dL/dv_sin = g_k * v_cos / v_k^2 + D_k * v_sin / v_k,   dL/dv_cos = -g_k * v_sin / v_k^2 + D_k * v_cos / v_k
// (a leaf child receives its gradient, an internal child its D).  The operand pairs (v_cos, v_sin)
// of the bottom-up recomputation are parked in caller scratch (2 * n_nodes * n elements).
#include <mutex>
#include <string.h>

#include "us_common.cuh"
#include "us_params.cuh"

namespace us {

struct BackwardParams {
  us_plan_t plan;
  const void* in[US_MAX_LEAVES];    // to: angles (s_nodes order)        from: leaves (c_nodes order)
  const void* grad[US_MAX_LEAVES];  // to: grad of leaves (c_nodes)      from: grad of angles (s_nodes); nullptr = 0
  void* out[US_MAX_LEAVES];         // to: grad of angles (s_nodes)      from: grad of leaves (c_nodes)
  const void* r_in;                 // to: r or nullptr (= 1)            from: grad of r or nullptr (= 0)
  void* r_out;                      // to: grad of r or nullptr
  void* scratch;                    // from: 2 * n_nodes * n elements
  int64_t n;
};
static_assert(sizeof(BackwardParams) < 32000, "kernel parameter block must fit the 32 KB limit");

template <typename T>
__device__ __forceinline__ void sincos1(T th, T& s, T& c) {
  const T tv[1] = {th};
  T sv[1], cv[1];
  sincos_vec<1>(tv, sv, cv);
  s = sv[0];
  c = cv[0];
}

template <typename T>
__global__ void __launch_bounds__(256) to_cartesian_backward(const __grid_constant__ BackwardParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int nn = p.plan.n_nodes;
  T stack[US_MAX_STACK];
  int sp = 0;
  // top-down: P_k into the grad_theta_k slot
  T cur = p.r_in ? __ldg((const T*)p.r_in + i) : T(1);
  for (int k = 0; k < nn; ++k) {
    const uint64_t w = p.plan.node[k];
    const int slot = node_slot(w);
    T sn, cs;
    sincos1(__ldg((const T*)p.in[slot] + i), sn, cs);
    ((T*)p.out[slot])[i] = cur;
    const T pc = cur * cs, ps = cur * sn;
    switch (node_kind(w)) {
      case US_NODE_A:
        if (sp > 0) cur = stack[--sp];
        break;
      case US_NODE_B:
        cur = ps;
        break;
      case US_NODE_BP:
        cur = pc;
        break;
      default:
        stack[sp++] = ps;
        cur = pc;
        break;
    }
  }
  // bottom-up: G of the finished subtree in `cur`
  auto leaf_grad = [&](int l) -> T { return p.grad[l] ? __ldg((const T*)p.grad[l] + i) : T(0); };
  sp = 0;
  cur = T(0);
  for (int k = nn - 1; k >= 0; --k) {
    const uint64_t w = p.plan.node[k];
    const int slot = node_slot(w);
    T sn, cs;
    sincos1(__ldg((const T*)p.in[slot] + i), sn, cs);
    T gc, gs;
    switch (node_kind(w)) {
      case US_NODE_A:
        if (k != nn - 1) stack[sp++] = cur;
        gc = leaf_grad(node_cos_leaf(w));
        gs = leaf_grad(node_sin_leaf(w));
        break;
      case US_NODE_B:
        gc = leaf_grad(node_cos_leaf(w));
        gs = cur;
        break;
      case US_NODE_BP:
        gs = leaf_grad(node_sin_leaf(w));
        gc = cur;
        break;
      default:
        gc = cur;
        gs = stack[--sp];
        break;
    }
    T* const o = (T*)p.out[slot];
    o[i] = o[i] * (cs * gs - sn * gc);
    cur = cs * gc + sn * gs;
  }
  if (p.r_out) ((T*)p.r_out)[i] = cur;
}

template <typename T>
__global__ void __launch_bounds__(256) from_cartesian_backward(const __grid_constant__ BackwardParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int nn = p.plan.n_nodes;
  T* const park = (T*)p.scratch;
  T stack[US_MAX_STACK];
  int sp = 0;
  // bottom-up recomputation of the operand pairs (same walk as from_cartesian_strided)
  T cur = T(0);
  for (int k = nn - 1; k >= 0; --k) {
    const uint64_t w = p.plan.node[k];
    T vc, vs;
    switch (node_kind(w)) {
      case US_NODE_A:
        if (k != nn - 1) stack[sp++] = cur;
        vc = __ldg((const T*)p.in[node_cos_leaf(w)] + i);
        vs = __ldg((const T*)p.in[node_sin_leaf(w)] + i);
        break;
      case US_NODE_B:
        vc = __ldg((const T*)p.in[node_cos_leaf(w)] + i);
        vs = cur;
        break;
      case US_NODE_BP:
        vs = __ldg((const T*)p.in[node_sin_leaf(w)] + i);
        vc = cur;
        break;
      default:
        vc = cur;
        vs = stack[--sp];
        break;
    }
    park[(int64_t)(2 * k) * p.n + i] = vc;
    park[(int64_t)(2 * k + 1) * p.n + i] = vs;
    cur = sqrt(vc * vc + vs * vs);
  }
  // top-down: D of the subtree being entered in `cur`
  sp = 0;
  cur = p.r_in ? __ldg((const T*)p.r_in + i) : T(0);
  for (int k = 0; k < nn; ++k) {
    const uint64_t w = p.plan.node[k];
    const int slot = node_slot(w);
    const T vc = park[(int64_t)(2 * k) * p.n + i], vs = park[(int64_t)(2 * k + 1) * p.n + i];
    const T v2 = vc * vc + vs * vs;
    const T inv2 = v2 > T(0) ? T(1) / v2 : T(0);  // an all-zero subtree has no direction: gradient 0
    const T inv = v2 > T(0) ? T(1) / sqrt(v2) : T(0);
    const T gt = p.grad[slot] ? __ldg((const T*)p.grad[slot] + i) : T(0);
    const T ds = gt * vc * inv2 + cur * vs * inv;
    const T dc = -gt * vs * inv2 + cur * vc * inv;
    switch (node_kind(w)) {
      case US_NODE_A:
        ((T*)p.out[node_cos_leaf(w)])[i] = dc;
        ((T*)p.out[node_sin_leaf(w)])[i] = ds;
        if (sp > 0) cur = stack[--sp];
        break;
      case US_NODE_B:
        ((T*)p.out[node_cos_leaf(w)])[i] = dc;
        cur = ds;
        break;
      case US_NODE_BP:
        ((T*)p.out[node_sin_leaf(w)])[i] = ds;
        cur = dc;
        break;
      default:
        stack[sp++] = ds;
        cur = dc;
        break;
    }
  }
}

static int check_backward(const us_plan_t* plan, int dtype, int64_t n, const char* who) {
  if (int rc = check_plan(plan, who)) return rc;
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "%s: bad dtype %d", who, dtype);
  if (n < 0) return fail(US_EINVAL, "%s: n < 0", who);
  if ((n + 255) / 256 > 0x7FFFFFFFLL) return fail(US_EINVAL, "%s: n too large for one launch", who);
  return 0;
}

}  // namespace us

using namespace us;

extern "C" int us_to_cartesian_backward(const us_plan_t* plan, int dtype, int64_t n, const void* const* angle_dev,
                                        const void* r_dev, const void* const* grad_leaf_dev, void* const* grad_angle_dev,
                                        void* grad_r_dev, void* stream) {
  static const char* who = "us_to_cartesian_backward";
  if (int rc = check_backward(plan, dtype, n, who)) return rc;
  if (n == 0) return 0;
  if (!angle_dev || !grad_leaf_dev || !grad_angle_dev) return fail(US_EINVAL, "%s: null pointer table", who);
  static BackwardParams p;  // 8 KB parameter block; copied at launch, guarded for concurrent callers
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  memset(&p, 0, sizeof p);
  memcpy(&p.plan, plan, sizeof(us_plan_t));
  for (int s = 0; s < plan->n_nodes; ++s) {
    if (!angle_dev[s] || !grad_angle_dev[s]) return fail(US_EINVAL, "%s: angle %d or its gradient buffer is null", who, s);
    p.in[s] = angle_dev[s];
    p.out[s] = grad_angle_dev[s];
  }
  for (int l = 0; l < plan->n_leaves; ++l) p.grad[l] = grad_leaf_dev[l];
  p.r_in = r_dev;
  p.r_out = grad_r_dev;
  p.n = n;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == US_F32)
    to_cartesian_backward<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else
    to_cartesian_backward<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "to_cartesian_backward launch");
  return 0;
}

extern "C" int us_from_cartesian_backward(const us_plan_t* plan, int dtype, int64_t n, const void* const* x_dev,
                                          const void* grad_r_dev, const void* const* grad_angle_dev, void* const* grad_x_dev,
                                          void* scratch_dev, void* stream) {
  static const char* who = "us_from_cartesian_backward";
  if (int rc = check_backward(plan, dtype, n, who)) return rc;
  if (n == 0) return 0;
  if (!x_dev || !grad_angle_dev || !grad_x_dev || !scratch_dev) return fail(US_EINVAL, "%s: null pointer table or scratch", who);
  static BackwardParams p;
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  memset(&p, 0, sizeof p);
  memcpy(&p.plan, plan, sizeof(us_plan_t));
  for (int l = 0; l < plan->n_leaves; ++l) {
    if (!x_dev[l] || !grad_x_dev[l]) return fail(US_EINVAL, "%s: leaf %d or its gradient buffer is null", who, l);
    p.in[l] = x_dev[l];
    p.out[l] = grad_x_dev[l];
  }
  for (int s = 0; s < plan->n_nodes; ++s) p.grad[s] = grad_angle_dev[s];
  p.r_in = grad_r_dev;
  p.scratch = scratch_dev;
  p.n = n;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == US_F32)
    from_cartesian_backward<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else
    from_cartesian_backward<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "from_cartesian_backward launch");
  return 0;
}
