// Cartesian <-> polyspherical transforms: contiguous fast path.
//
// Persistent, warp-specialised kernels: grid = #SMs x resident CTAs, each CTA = 8 consumer warps
// + 1 producer warp, looping over tiles of 256*V points.  Each input array contributes one row of
// a tile; the producer brings rows into shared memory with TMA bulk copies (cp.async.bulk, one
// per row) that complete on a "full" mbarrier, two stages deep, so a whole tile (tens of KB) is in
// flight per CTA while the previous one is computed — memory-level parallelism comes from the
// copy engine, not from registers.  Consumer warps never wait for each other: a warp that has
// finished reading a stage arrives on that stage's "empty" mbarrier, and only the producer waits
// for all eight arrivals before refilling it.  Every consumer thread owns V consecutive points:
// one 128-bit LDS per input row, the compiled pre-order program (kernel-parameter constant bank,
// warp-uniform control flow), one 128-bit streaming STG per output row.
//
// HBM traffic = algorithmic traffic: every input element is read once, every output written once
// (2 * c_ndim * sizeof(T) bytes per point and direction).
//
// Reference semantics: src/ultrasphere/_coordinates.py:555-579 (from) and :636-656 (to).
#include <stdlib.h>

#include <mutex>

#include "us_common.cuh"
#include "us_params.cuh"
#include "us_tma.cuh"

namespace us {

constexpr int kTileThreads = 256;                // consumer threads = points per tile / V
constexpr int kBlockThreads = kTileThreads + 32;  // + one producer warp
constexpr int kStages = 2;
// 228 KB of shared memory per SM, 1 KB reserved per CTA: two CTAs of two stages each
constexpr int kMaxStageBytes = 55 * 1024;
constexpr int kSmemCeiling = 2 * kMaxStageBytes + 1024;

// Issue the copies of one tile (one elected thread: UBLKCP takes uniform operands, so spreading
// the rows over lanes would only serialise them again): post the expected byte count, then one
// bulk copy per input row.
template <typename T>
__device__ __forceinline__ void issue_tile(const XformParams& p, int n_rows, bool r_row, T* stage, int P,
                                           int64_t first_point, uint64_t* bar, uint64_t policy) {
  const uint32_t row_bytes = (uint32_t)(P * sizeof(T));
  mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)n_rows);
  const int n_plain = r_row ? n_rows - 1 : n_rows;
#pragma unroll 1
  for (int row = 0; row < n_plain; ++row)
    tma_load_row(stage + (size_t)row * P, (const T*)p.in[row] + first_point, row_bytes, bar, policy);
  if (r_row) tma_load_row(stage + (size_t)n_plain * P, (const T*)p.r_in + first_point, row_bytes, bar, policy);
}

struct TileShared {
  uint64_t full[kStages];   // producer -> consumers: the stage's bytes have landed (tx count)
  uint64_t empty[kStages];  // consumers -> producer: all 8 consumer warps are done with the stage
};

__device__ __forceinline__ void init_barriers(TileShared* sh) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], kTileThreads / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
}

// The producer warp: one elected lane walks this CTA's tiles, waits until the stage it is about to
// overwrite has been released, and issues the tile's copies.
template <typename T>
__device__ __forceinline__ void producer_loop(const XformParams& p, int n_rows, bool r_row, T* stage0, size_t stage_elems,
                                              int P, int64_t n_tiles, TileShared* sh) {
  if ((threadIdx.x & 31) != 0) return;
  const uint64_t policy = policy_evict_first();
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    if (it >= kStages) mbar_wait(&sh->empty[s], (uint32_t)(((it / kStages) - 1) & 1));
    issue_tile<T>(p, n_rows, r_row, stage0 + s * stage_elems, P, tile * P, &sh->full[s], policy);
  }
}

// A consumer warp is done with a stage: order its shared-memory reads, then one arrival.
__device__ __forceinline__ void release_stage(TileShared* sh, int s) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&sh->empty[s]);
}


// =================================================================================================
// to_cartesian
// =================================================================================================
template <typename T, int V, bool STACK>
__global__ void __launch_bounds__(kBlockThreads, 2) to_cartesian_tiled(const __grid_constant__ XformParams p, int64_t n_tiles) {
  constexpr int P = kTileThreads * V;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nn = p.plan.n_nodes;
  const bool has_r = p.r_in != nullptr;
  const int n_rows = nn + (has_r ? 1 : 0);
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  const size_t stage_elems = (size_t)n_rows * P;
  TileShared* sh = reinterpret_cast<TileShared*>(smem_raw + kStages * stage_elems * sizeof(T));
  init_barriers(sh);
  const int tid = threadIdx.x;
  if (tid >= kTileThreads) {
    producer_loop<T>(p, n_rows, has_r, stage0, stage_elems, P, n_tiles, sh);
    return;
  }
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const T* stage = stage0 + s * stage_elems;
    mbar_wait(&sh->full[s], (uint32_t)((it / kStages) & 1));
    const int64_t g = tile * P + (int64_t)tid * V;
    Vec<T, V> cur;
    if (has_r) {
      cur = lds_vec<T, V>(stage + (size_t)nn * P, tid);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(1);
    }
    Vec<T, V> stack[STACK ? US_MAX_STACK : 1];
    int sp = 0;
    uint64_t w = p.plan.node[0];
    Vec<T, V> th = lds_vec<T, V>(stage + (size_t)node_slot(w) * P, tid);
#pragma unroll(sizeof(T) == 4 ? 1 : 2)  // A/B on one box: float32 1 % faster without unrolling, float64 2-4 % faster with 2
    for (int k = 0; k < nn; ++k) {
      // fetch the next node's word and angles before the long sincos chains of this one
      const uint64_t w_next = p.plan.node[(k + 1 < nn) ? k + 1 : k];
      const Vec<T, V> th_next = lds_vec<T, V>(stage + (size_t)node_slot(w_next) * P, tid);
      Vec<T, V> pc, ps, sn, cs;
      sincos_vec<V>(th.v, sn.v, cs.v);
      mul_vec<V>(cur.v, cs.v, pc.v);
      mul_vec<V>(cur.v, sn.v, ps.v);
      // data-flow form of the four node kinds (all predicates are warp-uniform):
      //   a : both products are leaves, resume the most recently deferred branch
      //   b : cos product is a leaf, continue along sin      b': sin product is a leaf, continue along cos
      //   c : defer the sin product, continue along cos
      const int kind = node_kind(w);
      T* const oc = (kind == US_NODE_A || kind == US_NODE_B) ? (T*)p.out[node_cos_leaf(w)] : nullptr;
      T* const os = (kind == US_NODE_A || kind == US_NODE_BP) ? (T*)p.out[node_sin_leaf(w)] : nullptr;
      if constexpr (STACK) {
        if (kind == US_NODE_C) stack[sp++] = ps;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = (kind == US_NODE_B) ? ps.v[v] : pc.v[v];
      if constexpr (STACK) {
        if (kind == US_NODE_A && sp > 0) cur = stack[--sp];
      }
      if (oc) stg_vec<T, V>(oc + g, pc);
      if (os) stg_vec<T, V>(os + g, ps);
      w = w_next;
      th = th_next;
    }
    release_stage(sh, s);
  }
}

// =================================================================================================
// from_cartesian
// =================================================================================================
// CHAIN: 0 = any tree (node kinds decoded per node); 1 = `b...ba` (create_standard), 2 = `b'...b'a`
// (create_standard_prime): every node but the last has one leaf operand and the running norm as the
// other, so the walk needs no kind switch, no operand selection, and the atan2 can skip the sign
// logic of the operand that is known to be a norm.
template <typename T, int V, bool STACK, int CHAIN>
__global__ void __launch_bounds__(kBlockThreads, sizeof(T) == 8 ? 3 : 2) from_cartesian_tiled(const __grid_constant__ XformParams p, int64_t n_tiles) {
  constexpr int P = kTileThreads * V;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nn = p.plan.n_nodes;
  const int n_rows = p.plan.n_leaves;
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  const size_t stage_elems = (size_t)n_rows * P;
  TileShared* sh = reinterpret_cast<TileShared*>(smem_raw + kStages * stage_elems * sizeof(T));
  init_barriers(sh);
  const int tid = threadIdx.x;
  if (tid >= kTileThreads) {
    producer_loop<T>(p, n_rows, false, stage0, stage_elems, P, n_tiles, sh);
    return;
  }
  T* const r_out = (T*)p.r_out;
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const T* stage = stage0 + s * stage_elems;
    mbar_wait(&sh->full[s], (uint32_t)((it / kStages) & 1));
    const int64_t g = tile * P + (int64_t)tid * V;
    if (r_out) {
      // leaves in c_nodes order, sequential, each square and each add rounded once
      Vec<T, V> acc;
      for (int l = 0; l < n_rows; ++l) {
        const Vec<T, V> x = lds_vec<T, V>(stage + (size_t)l * P, tid);
        sumsq_step<V>(acc.v, x.v, l == 0);
      }
      sqrt_vec<V>(acc.v);
      stg_vec<T, V>(r_out + g, acc);
    }
    Vec<T, V> cur;
    if constexpr (CHAIN != 0) {
      {
        const uint64_t w = p.plan.node[nn - 1];  // the closing `a` node: two leaves
        const Vec<T, V> vc = lds_vec<T, V>(stage + (size_t)node_cos_leaf(w) * P, tid);
        const Vec<T, V> vs = lds_vec<T, V>(stage + (size_t)node_sin_leaf(w) * P, tid);
        Vec<T, V> th;
        node_from_vec<V>(vs.v, vc.v, th.v, cur.v);
        T* o = (T*)p.out[node_slot(w)];
        if (o) stg_vec<T, V>(o + g, th);
      }
#pragma unroll 2
      for (int k = nn - 2; k >= 0; --k) {
        const uint64_t w = p.plan.node[k];
        const int leaf_row = (CHAIN == 1) ? node_cos_leaf(w) : node_sin_leaf(w);
        const Vec<T, V> leaf = lds_vec<T, V>(stage + (size_t)leaf_row * P, tid);
        Vec<T, V> th, nxt;
        if constexpr (CHAIN == 1)
          node_from_vec<V, true, false>(cur.v, leaf.v, th.v, nxt.v);  // theta = atan2(norm, leaf)
        else
          node_from_vec<V, false, true>(leaf.v, cur.v, th.v, nxt.v);  // theta = atan2(leaf, norm)
        cur = nxt;
        T* o = (T*)p.out[node_slot(w)];
        if (o) stg_vec<T, V>(o + g, th);
      }
    } else {
      // Any tree: one specialised body per node kind (warp-uniform switch), so that no operand has
      // to be selected at run time and the atan2 can drop the sign logic of operands that are norms.
      Vec<T, V> stack[STACK ? US_MAX_STACK : 1];
      int sp = 0;
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(0);
      for (int k = nn - 1; k >= 0; --k) {
        const uint64_t w = p.plan.node[k];
        Vec<T, V> th, nxt;
        switch (node_kind(w)) {
          case US_NODE_A: {  // two leaves; a finished sibling subtree waits for its `c` parent
            if constexpr (STACK) {
              if (k != nn - 1) stack[sp++] = cur;
            }
            const Vec<T, V> vc = lds_vec<T, V>(stage + (size_t)node_cos_leaf(w) * P, tid);
            const Vec<T, V> vs = lds_vec<T, V>(stage + (size_t)node_sin_leaf(w) * P, tid);
            node_from_vec<V, false, false>(vs.v, vc.v, th.v, nxt.v);
            break;
          }
          case US_NODE_B: {  // cos child leaf, sin child = running norm (>= 0)
            const Vec<T, V> vc = lds_vec<T, V>(stage + (size_t)node_cos_leaf(w) * P, tid);
            node_from_vec<V, true, false>(cur.v, vc.v, th.v, nxt.v);
            break;
          }
          case US_NODE_BP: {  // sin child leaf, cos child = running norm (>= 0)
            const Vec<T, V> vs = lds_vec<T, V>(stage + (size_t)node_sin_leaf(w) * P, tid);
            node_from_vec<V, false, true>(vs.v, cur.v, th.v, nxt.v);
            break;
          }
          default: {  // both children internal: cos = running norm, sin = the deferred norm
            Vec<T, V> vs = cur;
            if constexpr (STACK) vs = stack[--sp];
            node_from_vec<V, true, true>(vs.v, cur.v, th.v, nxt.v);
            break;
          }
        }
        cur = nxt;
        T* o = (T*)p.out[node_slot(w)];
        if (o) stg_vec<T, V>(o + g, th);
      }
    }
    release_stage(sh, s);
  }
}

// =================================================================================================
// host side: eligibility, variant selection, launch, ragged tail
// =================================================================================================
static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

struct Variant {
  const void* fn;
  int v;
};

template <typename T, int V>
static const void* variant(bool to, bool stack, int chain) {
  if (to) return stack ? (const void*)&to_cartesian_tiled<T, V, true> : (const void*)&to_cartesian_tiled<T, V, false>;
  if (stack) return (const void*)&from_cartesian_tiled<T, V, true, 0>;
  if (chain == 1) return (const void*)&from_cartesian_tiled<T, V, false, 1>;
  if (chain == 2) return (const void*)&from_cartesian_tiled<T, V, false, 2>;
  return (const void*)&from_cartesian_tiled<T, V, false, 0>;
}

template <bool TO>
static const void* pick(int dtype, int v, bool stack, int chain) {
  if (dtype == US_F32) {
    if (v == 4) return variant<float, 4>(TO, stack, chain);
    if (v == 2) return variant<float, 2>(TO, stack, chain);
    return variant<float, 1>(TO, stack, chain);
  }
  if (v == 2) return variant<double, 2>(TO, stack, chain);
  return variant<double, 1>(TO, stack, chain);
}

// 1: every node but the last is `b` and the last is `a`; 2: same with `b'`; 0: anything else
static int chain_kind(const us_plan_t& plan) {
  const int nn = plan.n_nodes;
  if (nn < 2 || plan.max_stack != 0 || node_kind(plan.node[nn - 1]) != US_NODE_A) return 0;
  const int first = node_kind(plan.node[0]);
  if (first != US_NODE_B && first != US_NODE_BP) return 0;
  for (int k = 1; k < nn - 1; ++k)
    if (node_kind(plan.node[k]) != first) return 0;
  return first == US_NODE_B ? 1 : 2;
}

static int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
  }
  return cached;
}

// cudaFuncSetAttribute + the occupancy query cost a few microseconds each; their result only
// depends on (kernel, dynamic shared memory, device), so it is computed once and remembered.
int resident_ctas(const void* fn, size_t smem, int block_threads, int* per_sm) {
  struct Entry {
    const void* fn;
    size_t smem;
    int dev, per_sm;
  };
  static Entry cache[128];
  static int used = 0;
  static std::mutex lock;
  int dev = 0;
  US_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  std::lock_guard<std::mutex> guard(lock);
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == fn && cache[i].smem == smem && cache[i].dev == dev) {
      *per_sm = cache[i].per_sm;
      return 0;
    }
  // the attribute is a ceiling: only ever raise it, or a later, smaller tree would lock out a wider one
  size_t ceiling = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == fn && cache[i].dev == dev && cache[i].smem > ceiling) ceiling = cache[i].smem;
  if (smem > ceiling)
    US_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(smem)");
  US_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, fn, block_threads, smem), "occupancy query");
  if (used < 128) {
    cache[used++] = Entry{fn, smem, dev, *per_sm};
  } else {
    // table full (128 distinct kernel/tree-width pairs): keep working, just without the shortcut
    US_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCeiling), "cudaFuncSetAttribute(smem)");
  }
  return 0;
}

template <bool TO>
static int try_tiled(const XformParams& p, int dtype, cudaStream_t st) {
  if (p.ndim != 1) return kNotTaken;
  const int esz = dtype == US_F32 ? 4 : 8;
  const int n_in = TO ? p.plan.n_nodes : p.plan.n_leaves;
  const int n_out = TO ? p.plan.n_leaves : p.plan.n_nodes;
  const int n_rows = n_in + ((TO && p.r_in) ? 1 : 0);
  for (int i = 0; i < n_in; ++i)
    if (p.in_mask[i] != 1u || !aligned16(p.in[i])) return kNotTaken;
  if (TO && p.r_in && (p.r_mask != 1u || !aligned16(p.r_in))) return kNotTaken;
  for (int i = 0; i < n_out; ++i)
    if (p.out[i] && !aligned16(p.out[i])) return kNotTaken;
  if (!TO && p.r_out && !aligned16(p.r_out)) return kNotTaken;
  // widest vector whose two-stage tile still lets two CTAs share an SM
  int v = 16 / esz;
  while (v >= 1 && (int64_t)n_rows * kTileThreads * v * esz > kMaxStageBytes) v >>= 1;
  {
    static int forced_v = -1;  // US_TILE_V=<1|2|4>: tuning override (may leave one CTA per SM)
    if (forced_v < 0) {
      const char* e = getenv("US_TILE_V");
      forced_v = e ? atoi(e) : 0;
    }
    if (forced_v > 0 && forced_v <= 16 / esz) v = forced_v;
  }
  int P = kTileThreads * v;
  int64_t n_tiles = v >= 1 ? p.n / P : 0;
  if (v < 16 / esz) {
    // wide tree: full-width vectors through the row ring instead of narrower resident tiles
    const int ring_p = kTileThreads * (16 / esz);
    if (p.n / ring_p < 2 * sm_count()) return kNotTaken;
    const int rc = launch_ring<TO>(p, dtype, st, &P);
    if (rc != 0) return rc;
    n_tiles = p.n / P;
  } else {
  {
    static int64_t min_tiles = -1;  // US_TILED_MIN_TILES=<n>: tuning override of the small-batch cut-over
    if (min_tiles < 0) {
      const char* e = getenv("US_TILED_MIN_TILES");
      min_tiles = e ? atoll(e) : 4 * sm_count();
    }
    // small batches: below ~4 tiles per SM the pipeline's fill/drain costs more than it hides
    // (measured under CUDA-graph replay, create_spherical float64: 3e5 points 7.2 us strided vs 9.9 us tiled)
    if (n_tiles < min_tiles) return kNotTaken;
  }
  const void* fn = pick<TO>(dtype, v, p.plan.max_stack > 0, chain_kind(p.plan));
  const size_t smem = (size_t)kStages * n_rows * P * esz + sizeof(TileShared);
  int per_sm = 0;
  if (int rc = resident_ctas(fn, smem, kBlockThreads, &per_sm)) return rc;
  if (per_sm < 1) return kNotTaken;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  XformParams body = p;
  int64_t tiles = n_tiles;
  void* args[] = {(void*)&body, (void*)&tiles};
  US_CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kBlockThreads), args, smem, st), "tiled launch");
  }
  // ragged tail (< one tile): the strided kernel on the remaining points
  const int64_t done = n_tiles * P;
  const int64_t rest = p.n - done;
  if (rest > 0) {
    XformParams tail = p;
    const int64_t off = done * esz;
    for (int i = 0; i < n_in; ++i) tail.in[i] = (const char*)p.in[i] + off;
    for (int i = 0; i < n_out; ++i)
      if (p.out[i]) tail.out[i] = (char*)p.out[i] + off;
    if (p.r_in) tail.r_in = (const char*)p.r_in + off;
    if (p.r_out) tail.r_out = (char*)p.r_out + off;
    tail.n = rest;
    tail.shape[0] = rest;
    const int rc = TO ? launch_to_cartesian_strided(tail, dtype, st) : launch_from_cartesian_strided(tail, dtype, st);
    if (rc != 0) return rc;
  }
  return 0;
}

int try_launch_to_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st) { return try_tiled<true>(p, dtype, st); }
int try_launch_from_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st) { return try_tiled<false>(p, dtype, st); }

}  // namespace us
