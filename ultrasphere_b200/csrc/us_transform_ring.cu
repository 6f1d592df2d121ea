// Cartesian <-> polyspherical transforms: contiguous fast path for WIDE trees.
//
// The tile kernels (us_transform_tiled.cu) keep every input row of a tile resident, two stages
// deep: 2 * rows * 16 B of shared memory per consumer thread.  Beyond ~13 rows that no longer
// leaves room for 16-byte vectors and two CTAs per SM (create_random(32) has 33 rows).  Here the
// rows stream through a RING of kRing row slots instead: the producer warp issues one TMA bulk
// copy per (tile, row) in exactly the order the consumers will need them, each slot has its own
// full/empty mbarrier pair, and a consumer warp releases a slot as soon as it has used the row.
// Shared memory per CTA is kRing * 4 KB whatever the tree width, threads keep V = 16 B / sizeof(T).
//
//   to_cartesian    row order per tile: r (if given), then the angle of every node in program order
//   from_cartesian  two passes over a tile's leaves: c_nodes order for r (sequential sum, bit-exact),
//                   then the leaf operands of the nodes in reverse program order.  The second pass
//                   re-reads rows the first one just fetched: they are loaded with an L2 evict-last
//                   hint the first time and evict-first the second, so DRAM still sees each byte once.
#include "us_common.cuh"
#include "us_params.cuh"
#include "us_tma.cuh"

namespace us {

constexpr int kRing = 16;            // row slots (power of two)
constexpr int kRingConsumers = 256;  // consumer threads = points per tile / V
constexpr int kRingBlock = kRingConsumers + 32;

struct RingShared {
  uint64_t full[kRing];
  uint64_t empty[kRing];
};

// Consumer side, on shared-space addresses computed once per thread: `mine` already points at this
// thread's 16 bytes of slot 0, slots are P*sizeof(T) apart, barriers 8 bytes apart.
template <typename T, int V>
struct RingConsumer {
  uint32_t mine, full, empty;
  uint32_t q;
  bool lane0;
  static constexpr int P = kRingConsumers * V;
  __device__ __forceinline__ RingConsumer(const T* rows, RingShared* sh, int tid)
      : mine(smem_u32(rows) + (uint32_t)tid * V * sizeof(T)), full(smem_u32(&sh->full[0])), empty(smem_u32(&sh->empty[0])),
        q(0u), lane0((tid & 31) == 0) {}
  __device__ __forceinline__ Vec<T, V> get(int& slot) {
    slot = (int)(q & (kRing - 1));
    mbar_wait_s(full + (uint32_t)slot * 8u, (q / kRing) & 1u);
    ++q;
    return lds_vec_s<T, V>(mine + (uint32_t)slot * (uint32_t)(P * sizeof(T)));
  }
  __device__ __forceinline__ void release(int slot) {
    __syncwarp();
    if (lane0) mbar_arrive_s(empty + (uint32_t)slot * 8u);
  }
};

struct RingProducer {
  RingShared* sh;
  uint32_t q;
  template <typename T>
  __device__ __forceinline__ void put(T* rows, int P, const T* src, uint64_t policy) {
    const int slot = (int)(q & (kRing - 1));
    if (q >= (uint32_t)kRing) mbar_wait(&sh->empty[slot], ((q / kRing) - 1u) & 1u);
    const uint32_t bytes = (uint32_t)(P * sizeof(T));
    mbar_arrive_expect_tx(&sh->full[slot], bytes);
    tma_load_row(rows + (size_t)slot * P, src, bytes, &sh->full[slot], policy);
    ++q;
  }
};

__device__ __forceinline__ void ring_init(RingShared* sh) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], kRingConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
}

// =================================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(kRingBlock, 3) to_cartesian_ring(const __grid_constant__ XformParams p, int64_t n_tiles) {
  constexpr int P = kRingConsumers * V;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* rows = reinterpret_cast<T*>(smem_raw);
  RingShared* sh = reinterpret_cast<RingShared*>(smem_raw + (size_t)kRing * P * sizeof(T));
  ring_init(sh);
  const int nn = p.plan.n_nodes;
  const bool has_r = p.r_in != nullptr;
  const int tid = threadIdx.x;
  if (tid >= kRingConsumers) {
    if (tid != kRingConsumers) return;
    const uint64_t policy = policy_evict_first();
    RingProducer prod{sh, 0u};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t first = tile * P;
      if (has_r) prod.put<T>(rows, P, (const T*)p.r_in + first, policy);
#pragma unroll 1
      for (int k = 0; k < nn; ++k) prod.put<T>(rows, P, (const T*)p.in[node_slot(p.plan.node[k])] + first, policy);
    }
    return;
  }
  RingConsumer<T, V> ring(rows, sh, tid);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g = tile * P + (int64_t)tid * V;
    Vec<T, V> cur;
    if (has_r) {
      int slot;
      cur = ring.get(slot);
      // the row is in registers once it has been used; `cur` is consumed by the first product
      // below, but the slot can go back right away: force the use with a register barrier
#pragma unroll
      for (int v = 0; v < V; ++v) asm volatile("" ::"r"(*reinterpret_cast<const uint32_t*>(&cur.v[v])));
      ring.release(slot);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(1);
    }
    Vec<T, V> stack[US_MAX_STACK];
    int sp = 0;
    for (int k = 0; k < nn; ++k) {
      const uint64_t w = p.plan.node[k];
      int slot;
      const Vec<T, V> th = ring.get(slot);
      Vec<T, V> pc, ps, sn, cs;
      sincos_vec<V>(th.v, sn.v, cs.v);
      ring.release(slot);  // after the sincos consumed the angles
      mul_vec<V>(cur.v, cs.v, pc.v);
      mul_vec<V>(cur.v, sn.v, ps.v);
      const int kind = node_kind(w);
      T* const oc = (kind == US_NODE_A || kind == US_NODE_B) ? (T*)p.out[node_cos_leaf(w)] : nullptr;
      T* const os = (kind == US_NODE_A || kind == US_NODE_BP) ? (T*)p.out[node_sin_leaf(w)] : nullptr;
      if (kind == US_NODE_C) stack[sp++] = ps;
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = (kind == US_NODE_B) ? ps.v[v] : pc.v[v];
      if (kind == US_NODE_A && sp > 0) cur = stack[--sp];
      if (oc) stg_vec<T, V>(oc + g, pc);
      if (os) stg_vec<T, V>(os + g, ps);
    }
  }
}

// =================================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(kRingBlock, 3) from_cartesian_ring(const __grid_constant__ XformParams p, int64_t n_tiles) {
  constexpr int P = kRingConsumers * V;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* rows = reinterpret_cast<T*>(smem_raw);
  RingShared* sh = reinterpret_cast<RingShared*>(smem_raw + (size_t)kRing * P * sizeof(T));
  ring_init(sh);
  const int nn = p.plan.n_nodes;
  const int n_leaves = p.plan.n_leaves;
  T* const r_out = (T*)p.r_out;
  const int tid = threadIdx.x;
  if (tid >= kRingConsumers) {
    if (tid != kRingConsumers) return;
    const uint64_t keep = policy_evict_last(), drop = policy_evict_first();
    RingProducer prod{sh, 0u};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t first = tile * P;
      if (r_out) {
#pragma unroll 1
        for (int l = 0; l < n_leaves; ++l) prod.put<T>(rows, P, (const T*)p.in[l] + first, keep);
      }
#pragma unroll 1
      for (int k = nn - 1; k >= 0; --k) {
        const uint64_t w = p.plan.node[k];
        const int kind = node_kind(w);
        if (kind == US_NODE_A || kind == US_NODE_B) prod.put<T>(rows, P, (const T*)p.in[node_cos_leaf(w)] + first, drop);
        if (kind == US_NODE_A || kind == US_NODE_BP) prod.put<T>(rows, P, (const T*)p.in[node_sin_leaf(w)] + first, drop);
      }
    }
    return;
  }
  RingConsumer<T, V> ring(rows, sh, tid);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g = tile * P + (int64_t)tid * V;
    if (r_out) {
      Vec<T, V> acc;
      for (int l = 0; l < n_leaves; ++l) {
        int slot;
        const Vec<T, V> x = ring.get(slot);
        sumsq_step<V>(acc.v, x.v, l == 0);
        ring.release(slot);
      }
      sqrt_vec<V>(acc.v);
      stg_vec<T, V>(r_out + g, acc);
    }
    Vec<T, V> stack[US_MAX_STACK];
    int sp = 0;
    Vec<T, V> cur;
#pragma unroll
    for (int v = 0; v < V; ++v) cur.v[v] = T(0);
    for (int k = nn - 1; k >= 0; --k) {
      const uint64_t w = p.plan.node[k];
      Vec<T, V> th, nxt;
      switch (node_kind(w)) {  // one specialised body per node kind (see from_cartesian_tiled)
        case US_NODE_A: {
          if (k != nn - 1) stack[sp++] = cur;
          int slot_c, slot_s;
          const Vec<T, V> vc = ring.get(slot_c);
          const Vec<T, V> vs = ring.get(slot_s);
          node_from_vec<V, false, false>(vs.v, vc.v, th.v, nxt.v);
          ring.release(slot_c);
          ring.release(slot_s);
          break;
        }
        case US_NODE_B: {
          int slot_c;
          const Vec<T, V> vc = ring.get(slot_c);
          node_from_vec<V, true, false>(cur.v, vc.v, th.v, nxt.v);
          ring.release(slot_c);
          break;
        }
        case US_NODE_BP: {
          int slot_s;
          const Vec<T, V> vs = ring.get(slot_s);
          node_from_vec<V, false, true>(vs.v, cur.v, th.v, nxt.v);
          ring.release(slot_s);
          break;
        }
        default: {
          const Vec<T, V> vs = stack[--sp];
          node_from_vec<V, true, true>(vs.v, cur.v, th.v, nxt.v);
          break;
        }
      }
      cur = nxt;
      T* o = (T*)p.out[node_slot(w)];
      if (o) stg_vec<T, V>(o + g, th);
    }
  }
}

// =================================================================================================
template <bool TO>
static const void* ring_fn(int dtype) {
  if (dtype == US_F32) return TO ? (const void*)&to_cartesian_ring<float, 4> : (const void*)&from_cartesian_ring<float, 4>;
  return TO ? (const void*)&to_cartesian_ring<double, 2> : (const void*)&from_cartesian_ring<double, 2>;
}

template <bool TO>
int launch_ring(const XformParams& p, int dtype, cudaStream_t st, int* points_per_tile) {
  const int esz = dtype == US_F32 ? 4 : 8;
  const int P = kRingConsumers * (16 / esz);
  *points_per_tile = P;
  const int64_t n_tiles = p.n / P;
  const void* fn = ring_fn<TO>(dtype);
  const size_t smem = (size_t)kRing * P * esz + sizeof(RingShared);
  int dev = 0, sms = 148;
  US_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  US_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
  int per_sm = 0;
  if (int rc = resident_ctas(fn, smem, kRingBlock, &per_sm)) return rc;
  if (per_sm < 1) return kNotTaken;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  XformParams body = p;
  int64_t tiles = n_tiles;
  void* args[] = {(void*)&body, (void*)&tiles};
  US_CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kRingBlock), args, smem, st), "ring launch");
  return 0;
}

template int launch_ring<true>(const XformParams&, int, cudaStream_t, int*);
template int launch_ring<false>(const XformParams&, int, cudaStream_t, int*);

}  // namespace us
