// to_cartesian on a broadcast grid: the integrand's view of the transform.
//
// Inside integrate() the user's f receives s_ndim one-dimensional quadrature axes shaped
// (1,..,n_i,..,1) and typically calls to_cartesian(..., as_array=True) to materialise the Cartesian
// coordinates of the whole tensor grid (reference: tests/test_integral.py:88-92,
// src/ultrasphere/_coordinates.py:650-654).  The grid has prod(n_i) points but only sum(n_i)
// distinct angles, so evaluating sin/cos per grid point wastes the FP pipes (for float64, ~80
// DFMA-class instructions each).  Here a pre-pass tabulates (sin, cos) of every broadcast angle
// once, and the grid kernel is pure table look-up (L1-resident), multiply and 128-bit streaming
// stores: it is bound by the HBM write of c_ndim * sizeof(T) bytes per grid point.
#include "us_common.cuh"
#include "us_params.cuh"

namespace us {

template <typename T>
struct Pair;
template <>
struct Pair<float> { using type = float2; };
template <>
struct Pair<double> { using type = double2; };

struct TableJob {
  const void* src[US_MAX_LEAVES];
  void* dst[US_MAX_LEAVES];
  int64_t count[US_MAX_LEAVES];
  int32_t n_jobs;
};

template <typename T>
__global__ void __launch_bounds__(256) sincos_tables(const __grid_constant__ TableJob job) {
  const int j = blockIdx.y;
  const int64_t cnt = job.count[j];
  const T* src = (const T*)job.src[j];
  typename Pair<T>::type* dst = (typename Pair<T>::type*)job.dst[j];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
    const T th[1] = {src[i]};
    T s[1], c[1];
    sincos_vec<1>(th, s, c);
    typename Pair<T>::type sc;
    sc.x = s[0];
    sc.y = c[0];
    dst[i] = sc;
  }
}

template <typename T, int V>
struct alignas(sizeof(T) * V) BVec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* dst, const BVec<T, V>& x) {
  if constexpr (sizeof(T) * V == 16) {
    __stcs(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(&x));
  } else if constexpr (sizeof(T) * V == 8) {
    __stcs(reinterpret_cast<float2*>(dst), *reinterpret_cast<const float2*>(&x));
  } else {
    __stcs(reinterpret_cast<float*>(dst), *reinterpret_cast<const float*>(&x));
  }
}

// One thread owns U groups of V consecutive grid points (U*V divides the innermost extent, so they
// share all outer indices: one index decomposition and, for operands that do not span the
// innermost dim, one table look-up serve all of them).  I is the index type: uint32_t whenever the
// grid has < 2^32 points.
template <typename T, int V, int U, bool STACK, int MAXD, typename I>
__global__ void __launch_bounds__(256) to_cartesian_bcast(const __grid_constant__ XformParams p) {
  constexpr int E = U * V;  // elements per thread
  const I g0 = ((I)blockIdx.x * blockDim.x + threadIdx.x) * E;
  if (g0 >= (I)p.n) return;
  const int nd = p.ndim;
  I idx[MAXD];
  {
    I rem = g0;
#pragma unroll
    for (int d = MAXD - 1; d >= 0; --d) {
      if (d < nd) {
        const I e = (I)p.shape[d];
        const I q = rem / e;
        idx[d] = rem - q * e;
        rem = q;
      } else {
        idx[d] = 0;
      }
    }
  }
  // offset of an operand = dot(idx, dense strides of the dims its mask selects)
  auto offset = [&](uint32_t mask) -> I {
    I o = 0, mul = 1;
#pragma unroll
    for (int d = MAXD - 1; d >= 0; --d) {
      if (d < nd && ((mask >> d) & 1u)) {
        o += idx[d] * mul;
        mul *= (I)p.shape[d];
      }
    }
    return o;
  };
  const uint32_t last_bit = 1u << (nd - 1);
  T cur[E];
  if (p.r_in) {
    const I o = offset(p.r_mask);
    const I step = (p.r_mask & last_bit) ? 1 : 0;
#pragma unroll
    for (int v = 0; v < E; ++v) cur[v] = __ldg((const T*)p.r_in + o + (I)v * step);
  } else {
#pragma unroll
    for (int v = 0; v < E; ++v) cur[v] = T(1);
  }
  T stack[STACK ? US_MAX_STACK : 1][E];  // chains (create_standard...) need none
  int sp = 0;
  const int nn = p.plan.n_nodes;
  for (int k = 0; k < nn; ++k) {
    const uint64_t w = p.plan.node[k];
    const int slot = node_slot(w);
    const uint32_t mask = p.in_mask[slot];
    const I o = offset(mask);
    const bool varies = (mask & last_bit) != 0;  // warp-uniform
    T sn[E], cs[E];
    if (p.tab[slot]) {
      const typename Pair<T>::type* tb = (const typename Pair<T>::type*)p.tab[slot];
      if (varies) {
#pragma unroll
        for (int v = 0; v < E; ++v) {
          const typename Pair<T>::type sc = __ldg(tb + o + (I)v);
          sn[v] = sc.x;
          cs[v] = sc.y;
        }
      } else {
        const typename Pair<T>::type sc = __ldg(tb + o);
#pragma unroll
        for (int v = 0; v < E; ++v) {
          sn[v] = sc.x;
          cs[v] = sc.y;
        }
      }
    } else {
      T th[E];
#pragma unroll
      for (int v = 0; v < E; ++v) th[v] = __ldg((const T*)p.in[slot] + o + (varies ? (I)v : (I)0));
      sincos_vec<E>(th, sn, cs);
    }
    T pc[E], ps[E];
    mul_vec<E>(cur, cs, pc);
    mul_vec<E>(cur, sn, ps);
    const int kind = node_kind(w);
    T* const oc = (kind == US_NODE_A || kind == US_NODE_B) ? (T*)p.out[node_cos_leaf(w)] : nullptr;
    T* const os = (kind == US_NODE_A || kind == US_NODE_BP) ? (T*)p.out[node_sin_leaf(w)] : nullptr;
    if constexpr (STACK) {
      if (kind == US_NODE_C) {
#pragma unroll
        for (int v = 0; v < E; ++v) stack[sp][v] = ps[v];
        ++sp;
      }
    }
#pragma unroll
    for (int v = 0; v < E; ++v) cur[v] = (kind == US_NODE_B) ? ps[v] : pc[v];
    if constexpr (STACK) {
      if (kind == US_NODE_A && sp > 0) {
        --sp;
#pragma unroll
        for (int v = 0; v < E; ++v) cur[v] = stack[sp][v];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      BVec<T, V> tc, ts;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        tc.v[v] = pc[u * V + v];
        ts.v[v] = ps[u * V + v];
      }
      if (oc) store_vec<T, V>(oc + g0 + u * V, tc);
      if (os) store_vec<T, V>(os + g0 + u * V, ts);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------
static int64_t operand_count(uint32_t mask, int ndim, const int64_t* shape) {
  int64_t c = 1;
  for (int d = 0; d < ndim; ++d)
    if ((mask >> d) & 1u) c *= shape[d];
  return c;
}

static bool wants_table(uint32_t mask, int ndim, const int64_t* shape, int64_t n) {
  const uint32_t full = ndim >= 32 ? 0xFFFFFFFFu : ((1u << ndim) - 1u);
  if (mask == full) return false;
  return operand_count(mask, ndim, shape) * 4 <= n;  // at least 4 grid points per distinct angle
}

static int64_t align16(int64_t b) { return (b + 15) & ~(int64_t)15; }

int64_t bcast_workspace_bytes(const us_plan_t* plan, int dtype, int64_t n, int ndim, const int64_t* shape,
                              const uint32_t* angle_mask) {
  const int esz = dtype == US_F32 ? 4 : 8;
  int64_t bytes = 0;
  for (int s = 0; s < plan->n_nodes; ++s)
    if (wants_table(angle_mask[s], ndim, shape, n)) bytes += align16(2 * esz * operand_count(angle_mask[s], ndim, shape));
  return bytes;
}

template <typename T, int V, int U, bool STACK, int MAXD>
static void launch_bcast(const XformParams& p, cudaStream_t st) {
  const int64_t threads_total = (p.n + V * U - 1) / (V * U);
  const unsigned blocks = (unsigned)((threads_total + 255) / 256);
  if (p.n < (int64_t)0xFFFFFFFFLL - 4096)
    to_cartesian_bcast<T, V, U, STACK, MAXD, uint32_t><<<blocks, 256, 0, st>>>(p);
  else
    to_cartesian_bcast<T, V, U, STACK, MAXD, uint64_t><<<blocks, 256, 0, st>>>(p);
}

template <typename T, int MAXD>
static void launch_bcast_shape(const XformParams& p, bool vec_ok, cudaStream_t st) {
  constexpr int VMAX = 16 / (int)sizeof(T);
  if (p.plan.max_stack > 0) {  // trees with `c` nodes: one vector per thread, stack in local memory
    vec_ok ? launch_bcast<T, VMAX, 1, true, MAXD>(p, st) : launch_bcast<T, 1, 1, true, MAXD>(p, st);
  } else if (vec_ok) {
    // (U = 4 groups per thread was measured SLOWER: 2.39 vs 1.44 ms on C5 — each store instruction
    // of a warp then covers 32 x 16 B at a 64 B stride instead of one contiguous 512 B segment)
    launch_bcast<T, VMAX, 1, false, MAXD>(p, st);
  } else {
    launch_bcast<T, 1, 1, false, MAXD>(p, st);
  }
}

int try_launch_to_cartesian_bcast(XformParams& p, int dtype, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (!workspace || p.ndim > 8) return kNotTaken;
  const int esz = dtype == US_F32 ? 4 : 8;
  const int nn = p.plan.n_nodes;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return kNotTaken;
  TableJob job;
  job.n_jobs = 0;
  int64_t used = 0, longest = 0;
  for (int s = 0; s < nn; ++s) {
    p.tab[s] = nullptr;
    if (!wants_table(p.in_mask[s], p.ndim, p.shape, p.n)) continue;
    const int64_t cnt = operand_count(p.in_mask[s], p.ndim, p.shape);
    const int64_t need = align16(2 * esz * cnt);
    if (used + need > workspace_bytes) return kNotTaken;
    job.src[job.n_jobs] = p.in[s];
    job.dst[job.n_jobs] = (char*)workspace + used;
    job.count[job.n_jobs] = cnt;
    p.tab[s] = (char*)workspace + used;
    used += need;
    if (cnt > longest) longest = cnt;
    ++job.n_jobs;
  }
  if (job.n_jobs == 0) return kNotTaken;
  if ((p.n + 255) / 256 > 0x7FFFFFFFLL) return kNotTaken;
  {
    int64_t bx = (longest + 255) / 256;
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)job.n_jobs);
    if (dtype == US_F32)
      sincos_tables<float><<<grid, 256, 0, st>>>(job);
    else
      sincos_tables<double><<<grid, 256, 0, st>>>(job);
    US_CUDA_TRY(cudaGetLastError(), "sincos_tables launch");
  }
  // vector width: 16 bytes per thread and output row when the innermost extent and the row
  // alignment allow it
  const int vmax = 16 / esz;
  bool vec_ok = (p.shape[p.ndim - 1] % vmax) == 0;
  for (int l = 0; l < p.plan.n_leaves && vec_ok; ++l)
    if (p.out[l] && (reinterpret_cast<uintptr_t>(p.out[l]) & 15u)) vec_ok = false;
  const bool small = p.ndim <= 4;
  if (dtype == US_F32) {
    small ? launch_bcast_shape<float, 4>(p, vec_ok, st) : launch_bcast_shape<float, 8>(p, vec_ok, st);
  } else {
    small ? launch_bcast_shape<double, 4>(p, vec_ok, st) : launch_bcast_shape<double, 8>(p, vec_ok, st);
  }
  US_CUDA_TRY(cudaGetLastError(), "to_cartesian_bcast launch");
  return 0;
}

}  // namespace us
