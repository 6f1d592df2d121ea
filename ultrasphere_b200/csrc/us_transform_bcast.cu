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
#include <mutex>

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
  *reinterpret_cast<BVec<T, V>*>(dst) = x;  // plain (write-back) store: see the measurement note at the launcher
}

// Launch geometry of the grid kernel, precomputed on the host so that the kernel needs neither
// run-time divisions nor per-operand mask decoding: division by each extent via a multiply-high
// magic number (exact for dividends < 2^31), and per operand the element stride along every launch
// dim (0 where it is broadcast).
constexpr int kBcastDims = 8;
struct BcastParams {
  XformParams x;
  uint32_t stride[US_MAX_LEAVES][kBcastDims];  // per angle operand (s_nodes order)
  uint32_t r_stride[kBcastDims];
  uint32_t extent[kBcastDims];
  uint32_t magic[kBcastDims];
  uint32_t shift[kBcastDims];  // 0xFFFFFFFF: extent 1 (quotient = dividend)
};
static_assert(sizeof(BcastParams) < 32000, "kernel parameter block must fit the 32 KB limit");

__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t magic, uint32_t shift) {
  return shift == 0xFFFFFFFFu ? n : (__umulhi(n, magic) >> shift);
}

// One thread owns V consecutive grid points (V divides the innermost extent, so they share all
// outer indices).  Persistent grid-stride loop; grids of < 2^31 points (wider ones take the strided
// kernel).  MAXD = 4 or 8 keeps the index vector in registers.
template <typename T, int V, bool STACK, int MAXD>
__global__ void __launch_bounds__(256) to_cartesian_bcast(const __grid_constant__ BcastParams bp) {
  const XformParams& p = bp.x;
  const int nd = p.ndim;
  const uint32_t n = (uint32_t)p.n;
  const uint32_t step = gridDim.x * blockDim.x * V;
  const int nn = p.plan.n_nodes;
  for (uint32_t g0 = (blockIdx.x * blockDim.x + threadIdx.x) * V; g0 < n; g0 += step) {
    uint32_t idx[MAXD];
    {
      uint32_t rem = g0;
#pragma unroll
      for (int d = MAXD - 1; d >= 0; --d) {
        if (d < nd) {
          const uint32_t q = fast_div(rem, bp.magic[d], bp.shift[d]);
          idx[d] = rem - q * bp.extent[d];
          rem = q;
        } else {
          idx[d] = 0;
        }
      }
    }
    auto offset = [&](const uint32_t (&st)[kBcastDims]) -> uint32_t {
      uint32_t o = 0;
#pragma unroll
      for (int d = 0; d < MAXD; ++d) o += idx[d] * st[d];
      return o;
    };
    T cur[V];
    if (p.r_in) {
      const uint32_t o = offset(bp.r_stride);
      const uint32_t inc = bp.r_stride[nd - 1];
#pragma unroll
      for (int v = 0; v < V; ++v) cur[v] = __ldg((const T*)p.r_in + o + v * inc);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) cur[v] = T(1);
    }
    T stack[STACK ? US_MAX_STACK : 1][V];  // chains (create_standard...) need none
    int sp = 0;
    for (int k = 0; k < nn; ++k) {
      const uint64_t w = p.plan.node[k];
      const int slot = node_slot(w);
      const uint32_t o = offset(bp.stride[slot]);
      const uint32_t inc = bp.stride[slot][nd - 1];  // 1 if the operand varies along the innermost dim
      T sn[V], cs[V];
      if (p.tab[slot]) {
        const typename Pair<T>::type* tb = (const typename Pair<T>::type*)p.tab[slot];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const typename Pair<T>::type sc = __ldg(tb + o + v * inc);
          sn[v] = sc.x;
          cs[v] = sc.y;
        }
      } else {
        T th[V];
#pragma unroll
        for (int v = 0; v < V; ++v) th[v] = __ldg((const T*)p.in[slot] + o + v * inc);
        sincos_vec<V>(th, sn, cs);
      }
      T pc[V], ps[V];
      mul_vec<V>(cur, cs, pc);
      mul_vec<V>(cur, sn, ps);
      const int kind = node_kind(w);
      T* const oc = (kind == US_NODE_A || kind == US_NODE_B) ? (T*)p.out[node_cos_leaf(w)] : nullptr;
      T* const os = (kind == US_NODE_A || kind == US_NODE_BP) ? (T*)p.out[node_sin_leaf(w)] : nullptr;
      if constexpr (STACK) {
        if (kind == US_NODE_C) {
#pragma unroll
          for (int v = 0; v < V; ++v) stack[sp][v] = ps[v];
          ++sp;
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) cur[v] = (kind == US_NODE_B) ? ps[v] : pc[v];
      if constexpr (STACK) {
        if (kind == US_NODE_A && sp > 0) {
          --sp;
#pragma unroll
          for (int v = 0; v < V; ++v) cur[v] = stack[sp][v];
        }
      }
      BVec<T, V> tc, ts;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        tc.v[v] = pc[v];
        ts.v[v] = ps[v];
      }
      if (oc) store_vec<T, V>(oc + g0, tc);
      if (os) store_vec<T, V>(os + g0, ts);
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

template <typename T, int V, bool STACK, int MAXD>
static void launch_bcast(const BcastParams& bp, cudaStream_t st) {
  const int64_t threads_total = (bp.x.n + V - 1) / V;
  const int64_t want = (threads_total + 255) / 256;
  const int64_t cap = 148 * 16;  // a few waves of resident CTAs; the grid-stride loop covers the rest
  to_cartesian_bcast<T, V, STACK, MAXD><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(bp);
}

template <typename T, int MAXD>
static void launch_bcast_shape(const BcastParams& bp, bool vec_ok, cudaStream_t st) {
  constexpr int VMAX = 16 / (int)sizeof(T);
  // (more than one vector per thread was measured SLOWER: each store instruction of a warp then
  // covers 32 x 16 B at a 64 B stride instead of one contiguous 512 B segment)
  if (bp.x.plan.max_stack > 0) {  // trees with `c` nodes: deferred-branch stack in local memory
    vec_ok ? launch_bcast<T, VMAX, true, MAXD>(bp, st) : launch_bcast<T, 1, true, MAXD>(bp, st);
  } else {
    vec_ok ? launch_bcast<T, VMAX, false, MAXD>(bp, st) : launch_bcast<T, 1, false, MAXD>(bp, st);
  }
}

// dense element stride of an operand along every launch dim (0 where broadcast)
static void operand_strides(uint32_t mask, int ndim, const int64_t* shape, uint32_t (&out)[kBcastDims]) {
  uint32_t mul = 1;
  for (int d = 0; d < kBcastDims; ++d) out[d] = 0;
  for (int d = ndim - 1; d >= 0; --d) {
    if ((mask >> d) & 1u) {
      out[d] = mul;
      mul *= (uint32_t)shape[d];
    }
  }
}

int try_launch_to_cartesian_bcast(XformParams& p, int dtype, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (!workspace || p.ndim > kBcastDims || p.n >= (int64_t)0x7FFFFFFF) return kNotTaken;
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
  static BcastParams bp;  // 18 KB: too large for the stack of a deeply nested caller; guarded below
  static std::mutex bp_lock;
  std::lock_guard<std::mutex> guard(bp_lock);
  bp.x = p;
  for (int s = 0; s < nn; ++s) operand_strides(p.in_mask[s], p.ndim, p.shape, bp.stride[s]);
  operand_strides(p.r_in ? p.r_mask : 0u, p.ndim, p.shape, bp.r_stride);
  for (int d = 0; d < kBcastDims; ++d) {
    const uint32_t e = d < p.ndim ? (uint32_t)p.shape[d] : 1u;
    bp.extent[d] = e;
    if (e <= 1) {
      bp.magic[d] = 0;
      bp.shift[d] = 0xFFFFFFFFu;
    } else {
      int l = 0;
      while ((1ull << l) < e) ++l;  // ceil(log2 e)
      bp.magic[d] = (uint32_t)(((1ull << (31 + l)) / e) + 1);
      bp.shift[d] = (uint32_t)(l - 1);
    }
  }
  const bool small = p.ndim <= 4;
  if (dtype == US_F32) {
    small ? launch_bcast_shape<float, 4>(bp, vec_ok, st) : launch_bcast_shape<float, 8>(bp, vec_ok, st);
  } else {
    small ? launch_bcast_shape<double, 4>(bp, vec_ok, st) : launch_bcast_shape<double, 8>(bp, vec_ok, st);
  }
  US_CUDA_TRY(cudaGetLastError(), "to_cartesian_bcast launch");
  return 0;
}

}  // namespace us
