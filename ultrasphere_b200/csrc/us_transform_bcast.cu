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
//
// The same file holds the linear-functional variants (us_to_cartesian_dot): sum_l k_l * x_l folded
// into the transform instead of storing the leaves (DOT template flag of the grid kernel, and the
// warp-per-row kernel to_cartesian_dot_rows for chain trees).
#include <mutex>
#include <string.h>

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
// DOT (us_to_cartesian_dot): the leaf products are folded into sum_l coef[l] * x_l instead of being
// stored, so the kernel writes sizeof(T) instead of c_ndim * sizeof(T) bytes per grid point.
template <typename T, int V, bool STACK, int MAXD, bool DOT>
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
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = T(0);
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
      T* const oc = (!DOT && (kind == US_NODE_A || kind == US_NODE_B)) ? (T*)p.out[node_cos_leaf(w)] : nullptr;
      T* const os = (!DOT && (kind == US_NODE_A || kind == US_NODE_BP)) ? (T*)p.out[node_sin_leaf(w)] : nullptr;
      if constexpr (DOT) {
        const T* const coef = (const T*)p.dot_coef;
        if (kind == US_NODE_A || kind == US_NODE_B) {
          const T kc = __ldg(coef + node_cos_leaf(w));
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = fma(kc, pc[v], acc[v]);
        }
        if (kind == US_NODE_A || kind == US_NODE_BP) {
          const T ks = __ldg(coef + node_sin_leaf(w));
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = fma(ks, ps[v], acc[v]);
        }
      }
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
    if constexpr (DOT) {
      BVec<T, V> ta;
#pragma unroll
      for (int v = 0; v < V; ++v) ta.v[v] = acc[v];
      store_vec<T, V>((T*)p.dot_out + g0, ta);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Row kernel for chain trees (no `c` node: create_standard / standard_prime / spherical / polar) on
// tensor grids of up to 4 launch dims.  One warp owns one row of the innermost launch dim at a
// time; the 32 lanes cover the row with 16-byte vectors, U vectors per lane and 512 contiguous
// bytes per warp instruction.  The outer indices are decomposed once per row, and every node whose
// angle does not vary along the row — all but one on the grids integrate() builds — is evaluated
// ONCE per lane as a scalar (one table look-up, two multiplies); only from the first row-varying
// node on does the lane carry its U*V points.  Everything a node needs sits in one 64-byte
// descriptor in program order.  Used for the linear functional only (us_to_cartesian_dot): there the
// generic grid kernel above stays issue-bound on its ~175 instructions of index and parameter
// arithmetic per point although it writes c_ndim times less, while this one needs ~80 (measured,
// C5: 0.96 ms -> 0.57 ms).  For the leaf-storing transform both kernels are bound by the same HBM
// writes and the generic one, with its higher occupancy, is the faster (1.02 vs 1.09 ms under ncu).
struct RowNode {
  const void* tab;   // (sin, cos) table of the angle, or nullptr: evaluate from `in`
  const void* in;
  uint32_t stride[4];  // element stride of the operand along the launch dims (right-aligned; [3] = along the row)
  uint32_t kind, leaf_c, leaf_s, pad;  // leaf slots index the coefficients; US_NO_LEAF: none
  uint64_t pad2[2];                    // 64 bytes
};
constexpr int kRowNodes = 64;
struct RowParams {
  RowNode node[kRowNodes];
  const void* r_in;
  const void* dot_coef;
  void* dot_out;
  uint32_t r_stride[4];
  uint32_t extent[4], magic[4], shift[4];  // launch dims right-aligned: [3] is the row
  uint32_t n_rows, nn;
};

// NN > 0: the node count is a compile-time constant and the node loop is fully unrolled, so every
// descriptor field becomes a direct constant-bank operand, the per-node tests are hoisted out of
// the row loop and the coefficients are loaded once per thread (short chains: the trees
// integrate() is used with); NN = 0: run-time loop.
template <typename T, int V, int U, int NN>
__global__ void __launch_bounds__(256, 3) to_cartesian_dot_rows(const __grid_constant__ RowParams rp) {
  constexpr int W = U * V;
  constexpr int KC = NN ? NN : 1;
  using P2 = typename Pair<T>::type;
  const uint32_t E = rp.extent[3], n_rows = rp.n_rows;
  const int nn = NN ? NN : (int)rp.nn;
  const uint32_t lane_col = (threadIdx.x & 31u) * V, warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t first_row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const T* const coef = (const T*)rp.dot_coef;
  T kc_all[KC], ks_all[KC];  // coefficients of the leaves each node produces (0: no such leaf)
  if constexpr (NN > 0) {
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      kc_all[k] = rp.node[k].leaf_c != US_NO_LEAF ? __ldg(coef + rp.node[k].leaf_c) : T(0);
      ks_all[k] = rp.node[k].leaf_s != US_NO_LEAF ? __ldg(coef + rp.node[k].leaf_s) : T(0);
    }
  }
  for (uint32_t c0 = 0; c0 < E; c0 += 32u * W) {  // one pass for rows of up to 32*U*V points
    const uint32_t col0 = c0 + lane_col;          // this lane's vector j starts at col0 + j*32*V
    bool ok[U];
#pragma unroll
    for (int j = 0; j < U; ++j) ok[j] = col0 + (uint32_t)j * 32u * V < E;
    for (uint32_t row = first_row; row < n_rows; row += warps) {
      const uint32_t q2 = fast_div(row, rp.magic[2], rp.shift[2]), i2 = row - q2 * rp.extent[2];
      const uint32_t q1 = fast_div(q2, rp.magic[1], rp.shift[1]), i1 = q2 - q1 * rp.extent[1];
      const uint32_t i0 = q1;
      auto offset = [&](const uint32_t (&st)[4]) -> uint32_t { return i0 * st[0] + i1 * st[1] + i2 * st[2]; };
      bool uni = true;  // the running product (and the functional) is still the same for the whole row
      T cur_s = T(1), acc_s = T(0);
      T cur_v[W], acc_v[W];
      if (rp.r_in) {
        const uint32_t o = offset(rp.r_stride);
        if (rp.r_stride[3] == 0) {
          cur_s = __ldg((const T*)rp.r_in + o);
        } else {
          uni = false;
          const T* const base = (const T*)rp.r_in + o + col0;
#pragma unroll
          for (int j = 0; j < U; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) {
              cur_v[j * V + v] = ok[j] ? __ldg(base + j * 32 * V + v) : T(0);
              acc_v[j * V + v] = T(0);
            }
        }
      }
#pragma unroll(NN ? NN : 1)
      for (int k = 0; k < nn; ++k) {
        const RowNode& nk = rp.node[k];
        const uint32_t o = offset(nk.stride);
        const bool row_varying = nk.stride[3] != 0;
        const bool to_sin = nk.kind == US_NODE_B;  // the chain continues through the sin child
        const P2* const tb = (const P2*)nk.tab;
        T kc, ks;
        if constexpr (NN > 0) {
          kc = kc_all[k];
          ks = ks_all[k];
        } else {
          kc = nk.leaf_c != US_NO_LEAF ? __ldg(coef + nk.leaf_c) : T(0);
          ks = nk.leaf_s != US_NO_LEAF ? __ldg(coef + nk.leaf_s) : T(0);
        }
        // one step for one point: a zero coefficient adds an exact zero, so no leaf test is needed
        auto step = [&](T& cur, T& acc, T sn, T cs) {
          const T pc = mul_rn(cur, cs), ps = mul_rn(cur, sn);
          acc = fma(ks, ps, fma(kc, pc, acc));
          cur = to_sin ? ps : pc;
        };
        if (!row_varying) {
          T s0, c0v;
          if (tb) {
            const P2 sc = __ldg(tb + o);
            s0 = sc.x;
            c0v = sc.y;
          } else {
            const T th[1] = {__ldg((const T*)nk.in + o)};
            T s1[1], c1[1];
            sincos_vec<1>(th, s1, c1);
            s0 = s1[0];
            c0v = c1[0];
          }
          if (uni) {
            step(cur_s, acc_s, s0, c0v);
          } else {
#pragma unroll
            for (int i = 0; i < W; ++i) step(cur_v[i], acc_v[i], s0, c0v);
          }
          continue;
        }
        if (uni) {  // first node that differs along the row: the lane starts carrying its points
          uni = false;
#pragma unroll
          for (int i = 0; i < W; ++i) {
            cur_v[i] = cur_s;
            acc_v[i] = acc_s;
          }
        }
        if (tb) {
          const P2* const base = tb + o + col0;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            if (!ok[j]) continue;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const P2 sc = __ldg(base + j * 32 * V + v);
              step(cur_v[j * V + v], acc_v[j * V + v], sc.x, sc.y);
            }
          }
        } else {
          const T* const base = (const T*)nk.in + o + col0;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            if (!ok[j]) continue;
            T th[V], sn[V], cs[V];
#pragma unroll
            for (int v = 0; v < V; ++v) th[v] = __ldg(base + j * 32 * V + v);
            sincos_vec<V>(th, sn, cs);
#pragma unroll
            for (int v = 0; v < V; ++v) step(cur_v[j * V + v], acc_v[j * V + v], sn[v], cs[v]);
          }
        }
      }
      T* const dst = (T*)rp.dot_out + row * E + col0;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        BVec<T, V> ta;
#pragma unroll
        for (int v = 0; v < V; ++v) ta.v[v] = uni ? acc_s : acc_v[j * V + v];
        if (ok[j]) store_vec<T, V>(dst + j * 32 * V, ta);
      }
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
  if (bp.x.dot_coef)
    to_cartesian_bcast<T, V, STACK, MAXD, true><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(bp);
  else
    to_cartesian_bcast<T, V, STACK, MAXD, false><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(bp);
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

template <typename T, int U>
static void launch_rows_u(const RowParams& rp, cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T);
  const int64_t want = ((int64_t)rp.n_rows + 7) / 8, cap = 148 * 8;  // 8 warps per CTA, grid-stride over rows
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  switch (rp.nn) {
    case 2: to_cartesian_dot_rows<T, V, U, 2><<<grid, 256, 0, st>>>(rp); break;
    case 3: to_cartesian_dot_rows<T, V, U, 3><<<grid, 256, 0, st>>>(rp); break;
    case 4: to_cartesian_dot_rows<T, V, U, 4><<<grid, 256, 0, st>>>(rp); break;
    case 5: to_cartesian_dot_rows<T, V, U, 5><<<grid, 256, 0, st>>>(rp); break;
    case 6: to_cartesian_dot_rows<T, V, U, 6><<<grid, 256, 0, st>>>(rp); break;
    default: to_cartesian_dot_rows<T, V, U, 0><<<grid, 256, 0, st>>>(rp); break;
  }
}
template <typename T>
static void launch_rows(const RowParams& rp, cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint32_t vectors = rp.extent[3] / V;  // per row; 32 * U of them per pass of the warp
  if (vectors <= 32) launch_rows_u<T, 1>(rp, st);
  else if (vectors <= 64) launch_rows_u<T, 2>(rp, st);
  else if (vectors <= 96) launch_rows_u<T, 3>(rp, st);
  else launch_rows_u<T, 4>(rp, st);
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

int try_launch_to_cartesian_bcast(XformParams& p, int dtype, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                                  int functionals) {
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
  if (p.dot_coef && (reinterpret_cast<uintptr_t>(p.dot_out) & 15u)) vec_ok = false;  // rows are n*esz apart: n % vmax == 0
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
  // a functional of a chain tree on a grid with long enough rows: the row kernel (see to_cartesian_dot_rows)
  const int64_t inner = p.shape[p.ndim - 1];
  const bool rows_ok = p.dot_coef && vec_ok && p.plan.max_stack == 0 && p.ndim >= 2 && p.ndim <= 4 && nn <= kRowNodes && inner * esz >= 256 &&
                       p.n / inner >= 2 * 148 * 8 && !getenv("US_NO_ROWS");
  static RowParams rp;  // guarded by bp_lock like bp
  if (rows_ok) {
    memset(&rp, 0, sizeof rp);
    const int shift_d = 4 - p.ndim;  // right-align the launch dims
    for (int d = 0; d < 4; ++d) {
      const int src = d - shift_d;
      rp.extent[d] = src >= 0 ? bp.extent[src] : 1u;
      rp.magic[d] = src >= 0 ? bp.magic[src] : 0u;
      rp.shift[d] = src >= 0 ? bp.shift[src] : 0xFFFFFFFFu;
      rp.r_stride[d] = src >= 0 ? bp.r_stride[src] : 0u;
    }
    for (int k = 0; k < nn; ++k) {
      const uint64_t w = p.plan.node[k];
      const int slot = node_slot(w), kind = node_kind(w);
      RowNode& nk = rp.node[k];
      nk.tab = p.tab[slot];
      nk.in = p.in[slot];
      nk.kind = (uint32_t)kind;
      nk.leaf_c = (kind == US_NODE_A || kind == US_NODE_B) ? (uint32_t)node_cos_leaf(w) : (uint32_t)US_NO_LEAF;
      nk.leaf_s = (kind == US_NODE_A || kind == US_NODE_BP) ? (uint32_t)node_sin_leaf(w) : (uint32_t)US_NO_LEAF;
      for (int d = 0; d < 4; ++d) nk.stride[d] = d - shift_d >= 0 ? bp.stride[slot][d - shift_d] : 0u;
    }
    rp.r_in = p.r_in;
    rp.n_rows = (uint32_t)(p.n / inner);
    rp.nn = (uint32_t)nn;
  }
  for (int j = 0; j < (p.dot_coef ? functionals : 1); ++j) {
    if (p.dot_coef) {
      bp.x.dot_coef = (const char*)p.dot_coef + (size_t)j * p.plan.n_leaves * esz;
      bp.x.dot_out = (char*)p.dot_out + (size_t)j * p.n * esz;
    }
    if (rows_ok) {
      rp.dot_coef = bp.x.dot_coef;
      rp.dot_out = bp.x.dot_out;
      dtype == US_F32 ? launch_rows<float>(rp, st) : launch_rows<double>(rp, st);
    } else if (dtype == US_F32) {
      small ? launch_bcast_shape<float, 4>(bp, vec_ok, st) : launch_bcast_shape<float, 8>(bp, vec_ok, st);
    } else {
      small ? launch_bcast_shape<double, 4>(bp, vec_ok, st) : launch_bcast_shape<double, 8>(bp, vec_ok, st);
    }
    US_CUDA_TRY(cudaGetLastError(), "to_cartesian_bcast launch");
  }
  return 0;
}

}  // namespace us
