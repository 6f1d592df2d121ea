// Cartesian <-> polyspherical transforms: strided/broadcast kernels ("generic path").
//
// One thread owns one point and interprets the compiled pre-order program (us_plan) that sits
// in the kernel-parameter constant bank.  All control flow depends only on the program, so it
// is warp-uniform; the only per-thread state is the running product / norm and the small
// deferred-branch stack.  Inputs may be broadcast along any launch dimension (mask bits), which
// is what integrands inside integrate() pass: s_ndim one-dimensional axes of a tensor grid.
//
// Reference semantics: src/ultrasphere/_coordinates.py:555-579 (from) and :636-656 (to).
#include <mutex>

#include "us_common.cuh"
#include "us_params.cuh"

namespace us {

// Decompose a linear index over shape[0..ndim) once per thread; offsets of broadcast inputs are
// then a short dot product with the dense strides of the dims their mask selects.
struct Indexer {
  int64_t lin;
  int ndim;
  uint32_t full;
  int64_t idx[US_MAX_DIMS];
  const int64_t* shape;

  __device__ __forceinline__ Indexer(const XformParams& p, int64_t i)
      : lin(i), ndim(p.ndim), full(p.ndim >= 32 ? 0xFFFFFFFFu : ((1u << p.ndim) - 1u)), shape(p.shape) {
    if (ndim > 1) {
      int64_t rem = i;
      for (int d = ndim - 1; d >= 0; --d) {
        int64_t e = shape[d];
        int64_t qd = rem / e;
        idx[d] = rem - qd * e;
        rem = qd;
      }
    }
  }
  __device__ __forceinline__ int64_t off(uint32_t mask) const {
    if (mask == full) return lin;
    if (mask == 0u) return 0;
    int64_t o = 0, mul = 1;
    for (int d = ndim - 1; d >= 0; --d) {
      if ((mask >> d) & 1u) {
        o += idx[d] * mul;
        mul *= shape[d];
      }
    }
    return o;
  }
};

// DOT: the leaf products are not stored but folded into one linear functional sum_l coef[l] * x_l
// (us_to_cartesian_dot), in program order, one FMA per leaf.
template <typename T, bool DOT>
__global__ void __launch_bounds__(256) to_cartesian_strided(const __grid_constant__ XformParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  Indexer ix(p, i);
  const T* const coef = (const T*)p.dot_coef;
  T acc = T(0);
  T cur = p.r_in ? ld_stream((const T*)p.r_in + ix.off(p.r_mask)) : T(1);
  T stack[US_MAX_STACK];
  int sp = 0;
  const int nn = p.plan.n_nodes;
  for (int k = 0; k < nn; ++k) {
    const uint64_t w = p.plan.node[k];
    const int slot = node_slot(w);
    const T th = ld_stream((const T*)p.in[slot] + ix.off(p.in_mask[slot]));
    const T thv[1] = {th};
    T sv[1], cv[1];
    sincos_vec<1>(thv, sv, cv);
    const T pc = mul_rn(cur, cv[0]);
    const T ps = mul_rn(cur, sv[0]);
    T* oc = nullptr;
    T* os = nullptr;
    const int kind = node_kind(w);
    if constexpr (DOT) {
      if (kind == US_NODE_A || kind == US_NODE_B) acc = fma(__ldg(coef + node_cos_leaf(w)), pc, acc);
      if (kind == US_NODE_A || kind == US_NODE_BP) acc = fma(__ldg(coef + node_sin_leaf(w)), ps, acc);
    }
    switch (kind) {
      case US_NODE_A:
        if constexpr (!DOT) {
          oc = (T*)p.out[node_cos_leaf(w)];
          os = (T*)p.out[node_sin_leaf(w)];
        }
        if (sp > 0) cur = stack[--sp];
        break;
      case US_NODE_B:
        if constexpr (!DOT) oc = (T*)p.out[node_cos_leaf(w)];
        cur = ps;
        break;
      case US_NODE_BP:
        if constexpr (!DOT) os = (T*)p.out[node_sin_leaf(w)];
        cur = pc;
        break;
      default:  // US_NODE_C: defer the sin branch, descend into the cos branch
        stack[sp++] = ps;
        cur = pc;
        break;
    }
    if (oc) st_stream(oc + i, pc);
    if (os) st_stream(os + i, ps);
  }
  if constexpr (DOT) st_stream((T*)p.dot_out + i, acc);
}

template <typename T>
__global__ void __launch_bounds__(256) from_cartesian_strided(const __grid_constant__ XformParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  Indexer ix(p, i);
  const int nn = p.plan.n_nodes;
  // r: leaves in c_nodes order, sequential, every square and add rounded once.
  if (p.r_out) {
    T acc = T(0);
    for (int l = 0; l < p.plan.n_leaves; ++l) {
      const T x = __ldg((const T*)p.in[l] + ix.off(p.in_mask[l]));
      const T sq = mul_rn(x, x);
      acc = (l == 0) ? sq : add_rn(acc, sq);
    }
    T rv[1] = {acc};
    sqrt_vec<1>(rv);
    st_stream((T*)p.r_out + i, rv[0]);
  }
  // children before parents = the pre-order program walked backwards
  T stack[US_MAX_STACK];
  int sp = 0;
  T cur = T(0);
  for (int k = nn - 1; k >= 0; --k) {
    const uint64_t w = p.plan.node[k];
    T vc, vs;
    switch (node_kind(w)) {
      case US_NODE_A: {
        if (k != nn - 1) stack[sp++] = cur;  // a finished sibling subtree waits for its `c` parent
        const int lc = node_cos_leaf(w), ls = node_sin_leaf(w);
        vc = __ldg((const T*)p.in[lc] + ix.off(p.in_mask[lc]));
        vs = __ldg((const T*)p.in[ls] + ix.off(p.in_mask[ls]));
        break;
      }
      case US_NODE_B: {
        const int lc = node_cos_leaf(w);
        vc = __ldg((const T*)p.in[lc] + ix.off(p.in_mask[lc]));
        vs = cur;
        break;
      }
      case US_NODE_BP: {
        const int ls = node_sin_leaf(w);
        vs = __ldg((const T*)p.in[ls] + ix.off(p.in_mask[ls]));
        vc = cur;
        break;
      }
      default:
        vc = cur;
        vs = stack[--sp];
        break;
    }
    const T vsv[1] = {vs}, vcv[1] = {vc};
    T av[1], nv[1];
    // theta = atan2(sin, cos); numpy's vector_norm(stack((sin, cos))) = sqrt(sin*sin + cos*cos)
    node_from_vec<1>(vsv, vcv, av, nv);
    T* o = (T*)p.out[node_slot(w)];
    if (o) st_stream(o + i, av[0]);
    cur = nv[0];
  }
}

// x^e for a small non-negative integer e (exponentiation by squaring, warp-uniform control flow)
template <typename T>
__device__ __forceinline__ T powi(T x, int e) {
  T acc = T(1);
  while (e > 0) {
    if (e & 1) acc *= x;
    x *= x;
    e >>= 1;
  }
  return acc;
}

// jacobian: out = r * prod_k cos(theta_k)^a_k * sin(theta_k)^b_k; exponents ride in the low / high
// 16 bits of in_mask's neighbour table `pow` (one word per node)
struct JacobianParams {
  XformParams x;
  uint32_t pow[US_MAX_NODES];  // cos exponent | sin exponent << 16, s_nodes order
};

template <typename T>
__global__ void __launch_bounds__(256) jacobian_strided(const __grid_constant__ JacobianParams jp) {
  const XformParams& p = jp.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  Indexer ix(p, i);
  T acc = p.r_in ? ld_stream((const T*)p.r_in + ix.off(p.r_mask)) : T(1);
  const int nn = p.plan.n_nodes;
  for (int s = 0; s < nn; ++s) {
    const uint32_t e = jp.pow[s];
    if (e == 0) continue;  // both children are leaves: the node contributes nothing
    const T thv[1] = {ld_stream((const T*)p.in[s] + ix.off(p.in_mask[s]))};
    T sv[1], cv[1];
    sincos_vec<1>(thv, sv, cv);
    acc *= powi(cv[0], (int)(e & 0xFFFFu)) * powi(sv[0], (int)(e >> 16));
  }
  st_stream((T*)p.out[0] + i, acc);
}

int launch_jacobian_strided(const XformParams& p, const int32_t* cos_pow, const int32_t* sin_pow, int dtype, cudaStream_t st) {
  static JacobianParams jp;  // 10.5 KB parameter block; copied at launch, guarded for concurrent callers
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  jp.x = p;
  for (int s = 0; s < p.plan.n_nodes; ++s) {
    if (cos_pow[s] < 0 || cos_pow[s] > 0xFFFF || sin_pow[s] < 0 || sin_pow[s] > 0xFFFF)
      return fail(US_EINVAL, "us_jacobian: exponent of node %d out of range", s);
    jp.pow[s] = (uint32_t)cos_pow[s] | ((uint32_t)sin_pow[s] << 16);
  }
  const int threads = 256;
  const int64_t blocks = (p.n + threads - 1) / threads;
  if (blocks > 0x7FFFFFFFLL) return fail(US_EINVAL, "us_jacobian: n too large for one launch");
  if (dtype == US_F32)
    jacobian_strided<float><<<(unsigned)blocks, threads, 0, st>>>(jp);
  else
    jacobian_strided<double><<<(unsigned)blocks, threads, 0, st>>>(jp);
  US_CUDA_TRY(cudaGetLastError(), "jacobian_strided launch");
  return 0;
}

int launch_to_cartesian_strided(const XformParams& p, int dtype, cudaStream_t st) {
  const int threads = 256;
  const int64_t blocks = (p.n + threads - 1) / threads;
  if (blocks > 0x7FFFFFFFLL) return fail(US_EINVAL, "us_to_cartesian: n too large for one launch");
  if (p.dot_coef) {
    if (dtype == US_F32)
      to_cartesian_strided<float, true><<<(unsigned)blocks, threads, 0, st>>>(p);
    else
      to_cartesian_strided<double, true><<<(unsigned)blocks, threads, 0, st>>>(p);
  } else if (dtype == US_F32)
    to_cartesian_strided<float, false><<<(unsigned)blocks, threads, 0, st>>>(p);
  else
    to_cartesian_strided<double, false><<<(unsigned)blocks, threads, 0, st>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "to_cartesian_strided launch");
  return 0;
}

int launch_from_cartesian_strided(const XformParams& p, int dtype, cudaStream_t st) {
  const int threads = 256;
  const int64_t blocks = (p.n + threads - 1) / threads;
  if (blocks > 0x7FFFFFFFLL) return fail(US_EINVAL, "us_from_cartesian: n too large for one launch");
  if (dtype == US_F32)
    from_cartesian_strided<float><<<(unsigned)blocks, threads, 0, st>>>(p);
  else
    from_cartesian_strided<double><<<(unsigned)blocks, threads, 0, st>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "from_cartesian_strided launch");
  return 0;
}

}  // namespace us
