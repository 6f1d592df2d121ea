// Cartesian <-> polyspherical transforms: contiguous fast path for WIDE trees.
//
// The tile kernels (us_transform_tiled.cu) keep every input row of a tile resident.  Beyond ~13
// rows a pool of whole tiles no longer leaves room for enough consumer warps (create_random(32)
// has 33 rows).  Here the rows stream through a RING of row slots instead: the producer warp
// issues one TMA bulk copy per (tile, row) in exactly the order the consumers will need them, each
// slot has its own full/empty mbarrier pair, and a consumer warp releases a slot as soon as it has
// used the row (RingConsumer::release: not before its loads from the slot have returned).  Shared memory per CTA is (slots + parking rows) * slot size whatever the tree width.
// A thread owns U 16-byte chunks of every row (U * 16 B / sizeof(T) points; chunk u of a slot holds
// the 256 threads' u-th chunks back to back, so shared-memory loads stay conflict-free and every
// global store instruction of a warp covers 512 contiguous bytes): the per-row handshake, the
// descriptor fetch, the kind switch and the loop are paid once per thread and row, so U = 2 halves
// that overhead per point (it was about half of the float32 kernels' instruction stream).
// Deferred branches of `c` nodes are parked in shared memory rows indexed by the depth of the
// deferral (us_walk.cuh): no per-thread stack, no local memory.
//
//   to_cartesian    row order per tile: r (if given), then the angle of every node in program order
//   from_cartesian  two passes over a tile's leaves: c_nodes order for r (sequential sum, bit-exact),
//                   then the leaf operands of the nodes in reverse program order.  The second pass
//                   re-reads rows the first one just fetched: the first fetch leaves them in L2 at
//                   normal priority, the second takes them evict-first, so DRAM sees ~1.07x the
//                   algorithmic bytes.  (Tried and measured worse: the r pass last, the r pass in the
//                   middle of the walk, its rows interleaved with the node steps, evict-last hints.)
#include <mutex>

#include "us_common.cuh"
#include "us_params.cuh"
#include "us_tma.cuh"
#include "us_walk.cuh"

namespace us {

constexpr int kMaxRing = 32;         // row slots at most
// W consumer warps per CTA + the producer warp (four warps in twice the CTAs measured the same as
// eight); a chunk = the 16 bytes of every consumer thread
constexpr int kRingWarps = 8;
__host__ __device__ constexpr int ring_chunk_bytes(int w) { return w * 32 * 16; }

using RingParams = WalkParams<US_MAX_NODES>;

struct RingShared {
  uint64_t full[kMaxRing];
  uint64_t empty[kMaxRing];
};
constexpr int kRingHeader = 512;
static_assert(sizeof(RingShared) <= kRingHeader, "barrier block");

// V-wide values as U chunks of 16 bytes: chunk u lives ring_chunk_bytes(W) further on in a slot and
// 256 * V0 points further on in global memory (V0 = 16 / sizeof(T) points per chunk).
template <typename T, int U, int W>
__device__ __forceinline__ Vec<T, (16 / (int)sizeof(T)) * U> lds_chunks(uint32_t addr) {
  constexpr int V0 = 16 / (int)sizeof(T);
  Vec<T, V0 * U> out;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const Vec<T, V0> c = lds_vec_s<T, V0>(addr + (uint32_t)(u * ring_chunk_bytes(W)));
#pragma unroll
    for (int v = 0; v < V0; ++v) out.v[u * V0 + v] = c.v[v];
  }
  return out;
}
template <typename T, int U, int W>
__device__ __forceinline__ void stg_chunks(T* dst, const Vec<T, (16 / (int)sizeof(T)) * U>& x) {
  constexpr int V0 = 16 / (int)sizeof(T);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    Vec<T, V0> c;
#pragma unroll
    for (int v = 0; v < V0; ++v) c.v[v] = x.v[u * V0 + v];
    stg_vec<T, V0>(dst + u * (W * 32 * V0), c);
  }
}

// Consumer side, on shared-space addresses computed once per thread: `mine` already points at this
// thread's first 16 bytes of slot 0, slots are U * 4 KB apart, barriers 8 bytes apart.
template <typename T, int U, int W>
struct RingConsumer {
  static constexpr int V0 = 16 / (int)sizeof(T);
  static constexpr int V = V0 * U;
  uint32_t mine, full, empty;
  uint32_t slot, parity, n_slots;
  uint32_t zero;  // 0 at run time, unknown at compile time (WalkParams::zero)
  bool lane0;
  __device__ __forceinline__ RingConsumer(const unsigned char* rows, RingShared* sh, int tid, int slots, int zero_)
      : mine(smem_u32(rows) + (uint32_t)tid * 16u), full(smem_u32(&sh->full[0])), empty(smem_u32(&sh->empty[0])),
        slot(0u), parity(0u), n_slots((uint32_t)slots), zero((uint32_t)zero_), lane0((tid & 31) == 0) {}
  __device__ __forceinline__ Vec<T, V> get(uint32_t& used) {
    US_ASSERT(slot < n_slots && n_slots <= (uint32_t)kMaxRing);
    used = slot;
    mbar_wait_s(full + slot * 8u, parity);
    const Vec<T, V> x = lds_chunks<T, U, W>(mine + slot * (uint32_t)(U * ring_chunk_bytes(W)));
    if (++slot == n_slots) {
      slot = 0u;
      parity ^= 1u;
    }
    return x;
  }
  // Hand the slot back — AFTER this warp's loads from it have returned.  Nothing in the memory model
  // orders a plain ld.shared before a later mbarrier.arrive of the same thread in the shared-memory
  // pipeline, and ptxas is free to issue the arrive right behind the loads (it did, for the r rows):
  // with a shallow ring the producer's refill of the slot was then seen to overtake a load still
  // queued — a few dozen points per 10^9 with another row's values (tools/stress.py, three CTAs per
  // SM with four or five slots).  So the barrier address is made to depend on one register of every
  // 16-byte load (`& zero`: always 0, but ptxas cannot know): in-order issue then holds the arrive
  // until the loads' scoreboards clear.  Two integer instructions per chunk.
  __device__ __forceinline__ void release(uint32_t used, const Vec<T, V>& x) {
    uint32_t dep = 0u;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t w;
      memcpy(&w, &x.v[u * V0], 4);
      dep |= w;
    }
    __syncwarp();
    if (lane0) mbar_arrive_s(empty + used * 8u + (dep & zero));
  }
};

// The producer warp runs converged (uniform counters and addresses, us_tma.cuh); only the two
// instructions that post the byte count and issue the copy are predicated on the elected lane.
struct RingProducer {
  RingShared* sh;
  unsigned char* rows;
  uint32_t slot, round, n_slots;
  bool leader;
  template <uint32_t kRowBytes, typename T>
  __device__ __forceinline__ void put(const T* src, uint64_t policy) {
    if (round > 0u) {
      // one lane polls, the warp reconverges behind it: the copy below is a warp-level (uniform
      // datapath) instruction and must be reached exactly once per row
      if (leader) mbar_wait_one(&sh->empty[slot], (round - 1u) & 1u);
      __syncwarp();
    }
    if (leader) {
      mbar_arrive_expect_tx(&sh->full[slot], kRowBytes);
      tma_load_row(rows + (size_t)slot * kRowBytes, src, kRowBytes, &sh->full[slot], policy);
    }
    if (++slot == n_slots) {
      slot = 0u;
      ++round;
    }
  }
};

__device__ __forceinline__ RingShared* ring_init(unsigned char* smem_raw, int n_slots, int consumer_warps) {
  RingShared* sh = reinterpret_cast<RingShared*>(smem_raw);
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_slots; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], consumer_warps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  return sh;
}

// the parking rows: thread-private columns, plain shared-memory accesses
template <typename T, int U, int W>
__device__ __forceinline__ void park_store(uint32_t addr, const Vec<T, (16 / (int)sizeof(T)) * U>& x) {
  US_ASSERT(addr % 16u == 0u);
  uint32_t w[4 * U];
  memcpy(w, &x, 16 * U);
#pragma unroll
  for (int u = 0; u < U; ++u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + (uint32_t)(u * ring_chunk_bytes(W))), "r"(w[4 * u]), "r"(w[4 * u + 1]),
                 "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                 : "memory");
}

// =================================================================================================
template <typename T, int U, int W, int MINB>
__global__ void __launch_bounds__(W * 32 + 32, MINB) to_cartesian_ring(const __grid_constant__ RingParams p, int64_t n_tiles, int n_slots) {
  constexpr int V0 = 16 / (int)sizeof(T), V = V0 * U;
  constexpr int P = W * 32 * V;
  constexpr uint32_t kRowBytes = (uint32_t)(U * ring_chunk_bytes(W));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RingShared* sh = ring_init(smem_raw, n_slots, W);
  unsigned char* rows = smem_raw + kRingHeader;
  const int nn = p.n_nodes;
  const bool has_r = p.has_r != 0;
  const int tid = threadIdx.x;
  if (warp_index_uniform() == W) {
    const uint64_t policy = policy_evict_first();
    RingProducer prod{sh, rows, 0u, 0u, (uint32_t)n_slots, elect_one()};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t first = tile * P;
      if (has_r) prod.put<kRowBytes, T>((const T*)p.in[nn] + first, policy);
#pragma unroll 1
      for (int k = 0; k < nn; ++k) prod.put<kRowBytes, T>((const T*)p.in[p.d[k].row_a] + first, policy);
    }
    return;
  }
  RingConsumer<T, U, W> ring(rows, sh, tid, n_slots, p.zero);
  const uint32_t park = smem_u32(rows) + (uint32_t)n_slots * kRowBytes + (uint32_t)tid * 16u;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g = tile * P + (int64_t)tid * V0;
    Vec<T, V> cur;
    if (has_r) {
      uint32_t slot;
      cur = ring.get(slot);
      ring.release(slot, cur);  // tied to the loaded registers (RingConsumer::release): `cur` is first consumed much later
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(1);
    }
    for (int k = 0; k < nn; ++k) {
      const NodeDesc& d = p.d[k];
      uint32_t slot;
      const Vec<T, V> th = ring.get(slot);
      Vec<T, V> pc, ps, sn, cs;
      sincos_vec<V>(th.v, sn.v, cs.v);
      ring.release(slot, th);  // after the sincos consumed the angles
      mul_vec<V>(cur.v, cs.v, pc.v);
      mul_vec<V>(cur.v, sn.v, ps.v);
      const uint32_t kind = d.kind;
      US_ASSERT(((kind & kNodeKindMask) != US_NODE_C && !(kind & kNodeLinked)) || d.park < (uint32_t)p.max_park * kRowBytes);
      if ((kind & kNodeKindMask) == US_NODE_C) park_store<T, U, W>(park + d.park, ps);
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = ((kind & kNodeKindMask) == US_NODE_B) ? ps.v[v] : pc.v[v];
      if (kind & kNodeLinked) cur = lds_chunks<T, U, W>(park + d.park);
      T* const oc = (T*)d.out_a;
      T* const os = (T*)d.out_b;
      if (oc) stg_chunks<T, U, W>(oc + g, pc);
      if (os) stg_chunks<T, U, W>(os + g, ps);
    }
  }
}

// =================================================================================================
template <typename T, int U, int W, int MINB>
__global__ void __launch_bounds__(W * 32 + 32, MINB) from_cartesian_ring(const __grid_constant__ RingParams p, int64_t n_tiles, int n_slots) {
  constexpr int V0 = 16 / (int)sizeof(T), V = V0 * U;
  constexpr int P = W * 32 * V;
  constexpr uint32_t kRowBytes = (uint32_t)(U * ring_chunk_bytes(W));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RingShared* sh = ring_init(smem_raw, n_slots, W);
  unsigned char* rows = smem_raw + kRingHeader;
  const int nn = p.n_nodes;
  const int n_leaves = p.n_rows;
  T* const r_out = (T*)p.r_out;
  const int tid = threadIdx.x;
  if (warp_index_uniform() == W) {
    // first fetch (the r pass) at normal L2 priority, second fetch (the walk) evict-first so that the
    // row leaves L2 once it has been used twice; measured on create_random(32): DRAM reads 1.15x the
    // leaves' bytes (evict-last for the first fetch: 1.18x; no evict-first on the second: 1.4x)
    const uint64_t keep = policy_evict_normal(), drop = policy_evict_first();
    RingProducer prod{sh, rows, 0u, 0u, (uint32_t)n_slots, elect_one()};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t first = tile * P;
      if (r_out) {
#pragma unroll 1
        for (int l = 0; l < n_leaves; ++l) prod.put<kRowBytes, T>((const T*)p.in[l] + first, keep);
      }
#pragma unroll 1
      for (int k = nn - 1; k >= 0; --k) {
        const NodeDesc& d = p.d[k];
        const uint32_t kind = d.kind & kNodeKindMask;
        if (kind == US_NODE_A || kind == US_NODE_B) prod.put<kRowBytes, T>((const T*)p.in[d.row_a] + first, drop);
        if (kind == US_NODE_A || kind == US_NODE_BP) prod.put<kRowBytes, T>((const T*)p.in[d.row_b] + first, drop);
      }
    }
    return;
  }
  RingConsumer<T, U, W> ring(rows, sh, tid, n_slots, p.zero);
  const uint32_t park = smem_u32(rows) + (uint32_t)n_slots * kRowBytes + (uint32_t)tid * 16u;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g = tile * P + (int64_t)tid * V0;
    if (r_out) {
      Vec<T, V> acc;
      for (int l = 0; l < n_leaves; ++l) {
        uint32_t slot;
        const Vec<T, V> x = ring.get(slot);
        sumsq_step<V>(acc.v, x.v, l == 0);
        ring.release(slot, x);
      }
      sqrt_vec<V>(acc.v);
      stg_chunks<T, U, W>(r_out + g, acc);
    }
    Vec<T, V> cur;
#pragma unroll
    for (int v = 0; v < V; ++v) cur.v[v] = T(0);
    for (int k = nn - 1; k >= 0; --k) {
      const NodeDesc& d = p.d[k];
      const uint32_t kind = d.kind;
      Vec<T, V> th, nxt;
      switch (kind & kNodeKindMask) {  // one specialised body per node kind (see from_cartesian_tiled)
        case US_NODE_A: {
          US_ASSERT(!(kind & kNodeLinked) || d.park < (uint32_t)p.max_park * kRowBytes);
          if (kind & kNodeLinked) park_store<T, U, W>(park + d.park, cur);
          uint32_t slot_c, slot_s;
          const Vec<T, V> vc = ring.get(slot_c);
          const Vec<T, V> vs = ring.get(slot_s);
          node_from_vec<V, false, false>(vs.v, vc.v, th.v, nxt.v);
          ring.release(slot_c, vc);
          ring.release(slot_s, vs);
          break;
        }
        case US_NODE_B: {
          uint32_t slot_c;
          const Vec<T, V> vc = ring.get(slot_c);
          node_from_vec<V, true, false>(cur.v, vc.v, th.v, nxt.v);
          ring.release(slot_c, vc);
          break;
        }
        case US_NODE_BP: {
          uint32_t slot_s;
          const Vec<T, V> vs = ring.get(slot_s);
          node_from_vec<V, false, true>(vs.v, cur.v, th.v, nxt.v);
          ring.release(slot_s, vs);
          break;
        }
        default: {
          US_ASSERT(d.park < (uint32_t)p.max_park * kRowBytes);
          const Vec<T, V> vs = lds_chunks<T, U, W>(park + d.park);
          node_from_vec<V, true, true>(vs.v, cur.v, th.v, nxt.v);
          break;
        }
      }
      cur = nxt;
      T* o = (T*)d.out_a;
      if (o) stg_chunks<T, U, W>(o + g, th);
    }
  }
}

// =================================================================================================
// Geometry: U chunks per thread and CTAs per SM (which also picks the instantiation whose register
// budget allows that many).  Defaults from A/B on create_random(32) under sustained load
// (tools/ab_inproc.py with AB_SETTLE: the kernels run into the 1000 W cap, so what counts is the
// time per call once the clocks have settled): U = 2 everywhere; float32 to_cartesian three CTAs per
// SM (0.89 of the HBM peak against 0.86 with two), from_cartesian two (it wants the deeper ring:
// 0.73 against 0.70); float64 two.  The ring_u / ring_ctas knobs override.
struct RingGeometry {
  int u, ctas;
};
int ring_knob(int which);  // 0: ring_ctas, 1: ring_u, 2: ring_slots (us_transform_tiled.cu: Tuning)

static RingGeometry ring_geometry(int dtype, bool to) {
  RingGeometry g;
  g.u = ring_knob(1);
  if (g.u != 1 && g.u != 2) g.u = 2;
  const bool f32 = dtype == US_F32;
  const int ctas_max = g.u == 1 ? (f32 ? 4 : 3) : (f32 ? 3 : 2);
  g.ctas = ring_knob(0);
  if (g.ctas <= 0) g.ctas = g.u == 1 ? ctas_max : (f32 && to) ? 3 : 2;
  if (g.ctas > ctas_max) g.ctas = ctas_max;
  return g;
}

template <bool TO, typename T, int U, int MINB>
static const void* ring_kernel() {
  return TO ? (const void*)&to_cartesian_ring<T, U, 8, MINB> : (const void*)&from_cartesian_ring<T, U, 8, MINB>;
}
template <bool TO>
static const void* ring_fn(int dtype, const RingGeometry& g) {
  if (dtype == US_F32) {
    if (g.u == 1) return ring_kernel<TO, float, 1, 4>();
    return g.ctas >= 3 ? ring_kernel<TO, float, 2, 3>() : ring_kernel<TO, float, 2, 2>();
  }
  if (g.u == 1) return ring_kernel<TO, double, 1, 3>();
  return ring_kernel<TO, double, 2, 2>();
}

int ring_points_per_tile(int dtype) {  // the same for both directions
  const RingGeometry g = ring_geometry(dtype, true);
  return kRingWarps * 32 * (16 / (dtype == US_F32 ? 4 : 8)) * g.u;
}

template <bool TO>
int launch_ring(const XformParams& p, int dtype, cudaStream_t st, int* points_per_tile) {
  const RingGeometry g = ring_geometry(dtype, TO);
  const int P = kRingWarps * 32 * (16 / (dtype == US_F32 ? 4 : 8)) * g.u;
  *points_per_tile = P;
  const int64_t n_tiles = p.n / P;
  const void* fn = ring_fn<TO>(dtype, g);
  const int block = kRingWarps * 32 + 32;
  // row slots: as many as the CTA's share of shared memory holds after the parking rows
  const int row_bytes = g.u * ring_chunk_bytes(kRingWarps);
  const int park_rows = p.plan.max_stack;
  int n_slots = ((227 * 1024 / g.ctas) - 1024 - kRingHeader) / row_bytes - park_rows;
  if (n_slots > kMaxRing) n_slots = kMaxRing;
  // to_cartesian: four slots are enough and measurably better than more (fewer requests in flight:
  // 0.90-0.91 of the HBM peak against 0.88 with five); from_cartesian wants every slot it can get
  const int cap = ring_knob(2) > 0 ? ring_knob(2) : (TO ? 4 : kMaxRing);
  if (n_slots > cap) n_slots = cap;
  if (n_slots < 4) n_slots = 4;
  const size_t smem = (size_t)kRingHeader + (size_t)(n_slots + park_rows) * row_bytes;
  if (smem > 227u * 1024u) return kNotTaken;
  int per_sm = 0;
  if (int rc = resident_ctas(fn, smem, block, &per_sm)) return rc;
  if (per_sm < 1) return kNotTaken;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  static RingParams body;  // 10 KB parameter block: filled under a lock, copied by the launch
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  if (!build_walk<US_MAX_NODES>(p, TO, 1u, (uint32_t)row_bytes, body)) return kNotTaken;
  int64_t tiles = n_tiles;
  void* args[] = {(void*)&body, (void*)&tiles, (void*)&n_slots};
  US_CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3((unsigned)block), args, smem, st), "ring launch");
  return 0;
}

template int launch_ring<true>(const XformParams&, int, cudaStream_t, int*);
template int launch_ring<false>(const XformParams&, int, cudaStream_t, int*);

}  // namespace us
