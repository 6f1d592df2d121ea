// Cartesian <-> polyspherical transforms: contiguous fast path.
//
// One persistent CTA per SM: one producer warp + G consumer GROUPS of four warps.  A tile is 128*V
// points (V = U * 16 bytes / sizeof(T); U = 1, or 2 for float64 to_cartesian); each input array
// contributes one row of U * 2 KB.  The tiles
// of a CTA travel through a pool of S > G shared-memory buffers: the producer brings tile i into
// buffer i mod S with one TMA bulk copy per row (cp.async.bulk, completing on a "full" mbarrier),
// group i mod G works on it, and each of the group's warps arrives on the buffer's "empty"
// mbarrier when it has read what it needs.  "Full" barriers belong to the GROUP, one per tile it
// may have outstanding (its j-th tile uses barrier j mod 8): successive uses of one buffer go to
// different groups, and a parity wait on a barrier shared between groups can be satisfied by the
// wrong phase once the groups drift two uses apart (seen as stale tiles and hangs on the HBM-bound
// float32 trees); a barrier only one group waits on, in order, cannot alias.  So G tiles are being
// computed on while S - G tiles are in flight or landed, and no warp ever waits for a warp of another
// group.  The pool is cut to the tree (pool_geometry below): G = as many groups as the register file
// holds (7 for float32, 6 for float64, 4-5 with U = 2), S = G + 32 KB worth of buffers — NOT all of
// shared memory: more tiles in flight than the latency-bandwidth product needs slow DRAM down.
// (The first version ran 2-3 CTAs of 8 consumer warps with two stages each: trees of 9+ rows fit
// only two such CTAs, i.e. 4 warps per scheduler for kernels that are latency-bound, and ncu showed
// 9 % of the warp samples sitting in the full-barrier wait.)
//
// Every consumer thread owns U runs of 16 bytes of every row: one 128-bit LDS per input row and
// chunk, the node descriptors of us_walk.cuh (kernel-parameter constant bank, warp-uniform control
// flow), one 128-bit streaming STG per output row and chunk.  Deferred branches of `c` nodes are parked in dead rows of
// the tile (us_walk.cuh): no per-thread stack, no local memory.
//
// HBM traffic = algorithmic traffic: every input element is read once, every output written once
// (2 * c_ndim * sizeof(T) bytes per point and direction).
//
// Reference semantics: src/ultrasphere/_coordinates.py:555-579 (from) and :636-656 (to).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "us_common.cuh"
#include "us_params.cuh"
#include "us_tma.cuh"
#include "us_walk.cuh"

namespace us {

constexpr int kGroupThreads = 128;  // consumer threads per tile = points per tile / V
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kMaxBuffers = 32;
constexpr int kTileCap = 16;              // rows per tile the tile kernels are compiled for
constexpr int kSmemBytes = 227 * 1024;    // dynamic shared memory one CTA may ask for on sm_100
constexpr int kRowBytes = kGroupThreads * 16;  // one 16-byte chunk of every thread of a group: a row of a tile is U of these
// A thread owns U chunks of every row (V = U * 16 / sizeof(T) points).  U = 2 exists for float64
// only: those kernels are bound by instruction issue and, in a long run, by the energy per point
// (DESIGN.md §4.0), and the per-thread work that does not depend on V — descriptor fetches, the
// walk's control flow, addresses, barrier handshakes — halves per point.
// groups per CTA: 1 + 4 G warps must fit the register file (64 K registers) at the kernels'
// register counts — float32 <= 64 registers, float64 <= 80; U = 2: float64 four groups (17 warps, five
// of them on one scheduler's 16 K registers: <= 96), float32 five (six on one scheduler: <= 80)
template <typename T, int U = 1>
constexpr int max_groups() { return U == 2 ? (sizeof(T) == 4 ? 5 : 4) : (sizeof(T) == 4 ? 7 : 6); }

using TileParams = WalkParams<kTileCap>;

constexpr int kMaxGroups = 7;
constexpr int kGroupSlots = 8;  // full barriers per group >= tiles a group can have outstanding + 1 (S / G + 1 <= 8 is checked at launch)
struct PoolShared {
  uint64_t full[kMaxGroups][kGroupSlots];  // producer -> group: the bytes of the group's j-th tile have landed (tx count)
  uint64_t empty[kMaxBuffers];             // group -> producer: its four warps are done with the buffer
};
constexpr int kPoolHeader = 1024;
static_assert(sizeof(PoolShared) <= kPoolHeader, "barrier block");
static_assert((kGroupSlots & (kGroupSlots - 1)) == 0, "slot count is a power of two");

// The producer warp walks this CTA's tiles with all lanes converged (uniform control flow and
// addresses), waits until the buffer it is about to overwrite has been released, and its elected
// lane issues the tile's copies (UBLKCP takes uniform operands; spreading the rows over lanes
// would only serialise them again).
template <typename T, int U>
__device__ __forceinline__ void producer_loop(const TileParams& p, unsigned char* bufs, uint32_t buf_bytes, int n_groups, int n_bufs,
                                              int64_t n_tiles, int tail_points, PoolShared* sh) {
  constexpr int P = kGroupThreads * (16 / (int)sizeof(T)) * U;
  constexpr uint32_t kTileRowBytes = (uint32_t)(kRowBytes * U);
  const bool leader = elect_one();
  const uint64_t policy = policy_evict_first();
  const int n_rows = p.n_rows;
  int b = 0, g = 0, slot = 0;  // buffer, group and the group's barrier slot of the current tile
  uint32_t round = 0;          // how many times the pool has wrapped
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (round > 0) {
      // one lane polls, the warp reconverges behind it: the copies below are warp-level (uniform
      // datapath) instructions and must be reached exactly once per tile
      if (leader) mbar_wait_one(&sh->empty[b], (round - 1u) & 1u);
      __syncwarp();
    }
    unsigned char* dst = bufs + (size_t)b * buf_bytes;
    uint64_t* bar = &sh->full[g][slot];
    // the ragged last tile (tail_points > 0) copies only the bytes that exist: a multiple of 16
    const uint32_t row_bytes = (tail_points > 0 && tile == n_tiles - 1) ? (uint32_t)tail_points * (uint32_t)sizeof(T) : kTileRowBytes;
    if (leader) mbar_arrive_expect_tx(bar, (uint32_t)n_rows * row_bytes);
    const int64_t first = tile * P;
#pragma unroll 4
    for (int row = 0; row < n_rows; ++row) {
      const T* src = (const T*)p.in[row] + first;
      if (leader) tma_load_row(dst + (size_t)row * kTileRowBytes, src, row_bytes, bar, policy);
    }
    if (++b == n_bufs) {
      b = 0;
      ++round;
    }
    if (++g == n_groups) {
      g = 0;
      slot = (slot + 1) & (kGroupSlots - 1);
    }
  }
}

// Writes to shared memory by the consumers (parked branches) must be ordered before the TMA unit
// overwrites the buffer with the next tile: generic proxy -> async proxy.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// V-wide access to a thread's part of a row: chunk u sits kRowBytes further on in the row (shared
// memory) and kGroupThreads * V0 points further on in the arrays (V0 = points per 16-byte chunk), so
// that loads stay conflict-free and every store instruction of a warp covers 512 contiguous bytes.
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lds_at(const unsigned char* mine, uint32_t row_off) {
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
  US_ASSERT(row_off % (kRowBytes * U) == 0 && row_off < (kTileCap + 1) * kRowBytes * U);  // a whole row of this thread's tile
  Vec<T, V> out;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const Vec<T, V0> c = *reinterpret_cast<const Vec<T, V0>*>(mine + row_off + u * kRowBytes);
#pragma unroll
    for (int v = 0; v < V0; ++v) out.v[u * V0 + v] = c.v[v];
  }
  return out;
}
template <typename T, int V>
__device__ __forceinline__ void sts_at(unsigned char* mine, uint32_t row_off, const Vec<T, V>& x) {
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
  US_ASSERT(row_off % (kRowBytes * U) == 0 && row_off < (kTileCap + 1) * kRowBytes * U);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    Vec<T, V0> c;
#pragma unroll
    for (int v = 0; v < V0; ++v) c.v[v] = x.v[u * V0 + v];
    *reinterpret_cast<Vec<T, V0>*>(mine + row_off + u * kRowBytes) = c;
  }
}
// `live`: bit u set = chunk u of this thread lies inside the batch (all ones except in a ragged last tile)
template <typename T, int V>
__device__ __forceinline__ void stg_tile(T* dst, const Vec<T, V>& x, uint32_t live) {
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    Vec<T, V0> c;
#pragma unroll
    for (int v = 0; v < V0; ++v) c.v[v] = x.v[u * V0 + v];
    if (U == 1 || (live >> u) & 1u) stg_vec<T, V0>(dst + u * (kGroupThreads * V0), c);
  }
}
template <typename T, int V>
__device__ __forceinline__ uint32_t live_chunks(int t, int64_t tile, int64_t n_tiles, int tail_points) {
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
  if (tail_points == 0 || tile != n_tiles - 1) return (1u << U) - 1u;
  // in the ragged last tile a chunk holds V0 whole points or none (the tail is a multiple of 16 bytes)
  uint32_t m = 0;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u * (kGroupThreads * V0) + t * V0 < tail_points) m |= 1u << u;
  return m;
}

// Position of a consumer group in the tile sequence of its CTA: its j-th tile is tile
// blockIdx + (group + j G) gridDim, sits in buffer (group + j G) mod S and is announced on the
// group's barrier j mod kGroupSlots, phase parity (j / kGroupSlots) & 1.
struct TileCursor {
  int64_t tile;
  int b, slot;
  uint32_t parity;
  __device__ __forceinline__ TileCursor(int group) : tile((int64_t)blockIdx.x + (int64_t)group * gridDim.x), b(group), slot(0), parity(0u) {}
  __device__ __forceinline__ void next(int n_groups, int n_bufs) {
    tile += (int64_t)n_groups * gridDim.x;
    b += n_groups;  // n_groups < n_bufs: at most one wrap
    if (b >= n_bufs) b -= n_bufs;
    if (++slot == kGroupSlots) {
      slot = 0;
      parity ^= 1u;
    }
  }
};

__device__ __forceinline__ PoolShared* pool_init(unsigned char* smem_raw, int n_groups, int n_bufs) {
  PoolShared* sh = reinterpret_cast<PoolShared*>(smem_raw);
  if (threadIdx.x == 0) {
    for (int g = 0; g < n_groups; ++g)
      for (int s = 0; s < kGroupSlots; ++s) mbar_init(&sh->full[g][s], 1);
    for (int s = 0; s < n_bufs; ++s) mbar_init(&sh->empty[s], kGroupWarps);
    fence_mbar_init();
  }
  __syncthreads();
  return sh;
}

// =================================================================================================
// to_cartesian
// =================================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(32 + kGroupThreads * max_groups<T, V * (int)sizeof(T) / 16>(), 1)
    to_cartesian_tiled(const __grid_constant__ TileParams p, int64_t n_tiles, int n_groups, int n_bufs, int tail_points) {
  constexpr int P = kGroupThreads * V;
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
  constexpr uint32_t kTileRowBytes = (uint32_t)(kRowBytes * U);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PoolShared* sh = pool_init(smem_raw, n_groups, n_bufs);
  unsigned char* bufs = smem_raw + kPoolHeader;
  const int nn = p.n_nodes;
  const uint32_t buf_bytes = (uint32_t)p.n_rows * kTileRowBytes;
  const int tid = threadIdx.x;
  const int warp = warp_index_uniform();
  if (warp == 0) {
    producer_loop<T, U>(p, bufs, buf_bytes, n_groups, n_bufs, n_tiles, tail_points, sh);
    return;
  }
  const int t = (tid - 32) & (kGroupThreads - 1);
  const bool has_r = p.has_r != 0;
  const int group = (warp - 1) / kGroupWarps;
  for (TileCursor at(group); at.tile < n_tiles; at.next(n_groups, n_bufs)) {
    unsigned char* mine = bufs + (size_t)at.b * buf_bytes + t * 16;
    US_ASSERT(group < n_groups && at.b >= 0 && at.b < n_bufs && at.slot >= 0 && at.slot < kGroupSlots);
    US_ASSERT(kPoolHeader + (size_t)(at.b + 1) * buf_bytes <= (size_t)kSmemBytes);
    mbar_wait(&sh->full[group][at.slot], at.parity);
    const int64_t g = at.tile * P + (int64_t)t * V0;
    const uint32_t live = live_chunks<T, V>(t, at.tile, n_tiles, tail_points);
    if (live) {
    Vec<T, V> cur;
    if (has_r) {
      cur = lds_at<T, V>(mine, (uint32_t)nn * kTileRowBytes);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(1);
    }
    Vec<T, V> th = lds_at<T, V>(mine, p.d[0].row_a);
#pragma unroll(sizeof(T) == 4 ? 1 : 2)
    for (int k = 0; k < nn; ++k) {
      const NodeDesc& d = p.d[k];
      // fetch the next node's angles before the long sincos chains of this one
      const Vec<T, V> th_next = lds_at<T, V>(mine, p.d[(k + 1 < nn) ? k + 1 : k].row_a);
      Vec<T, V> pc, ps, sn, cs;
      sincos_vec<V>(th.v, sn.v, cs.v);
      mul_vec<V>(cur.v, cs.v, pc.v);
      mul_vec<V>(cur.v, sn.v, ps.v);
      // data-flow form of the four node kinds (all predicates are warp-uniform):
      //   a : both products are leaves, resume the most recently deferred branch
      //   b : cos product is a leaf, continue along sin      b': sin product is a leaf, continue along cos
      //   c : park the sin product, continue along cos
      const uint32_t kind = d.kind;
      if ((kind & kNodeKindMask) == US_NODE_C) sts_at<T, V>(mine, d.park, ps);
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = ((kind & kNodeKindMask) == US_NODE_B) ? ps.v[v] : pc.v[v];
      if (kind & kNodeLinked) cur = lds_at<T, V>(mine, d.park);
      T* const oc = (T*)d.out_a;
      T* const os = (T*)d.out_b;
      if (oc) stg_tile<T, V>(oc + g, pc, live);
      if (os) stg_tile<T, V>(os + g, ps, live);
      th = th_next;
    }
    }
    if (p.max_park) fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sh->empty[at.b]);
  }
}

// The checked walk of a chain for one thread's V points: every node through node_from_vec (fast path
// where its operands allow, libdevice otherwise — the same per-node choice the generic kernels
// make, so the values do not depend on which kernel ran).  Out of line: points with zero, tiny,
// huge or non-finite coordinates only.
template <typename T, int V, int CHAIN>
__device__ __noinline__ void chain_redo(const TileParams& p, const unsigned char* mine, int64_t g, uint32_t live) {
  const int nn = p.n_nodes;
  Vec<T, V> cur;
  {
    const NodeDesc& d = p.d[nn - 1];
    const Vec<T, V> vc = lds_at<T, V>(mine, d.row_a);
    const Vec<T, V> vs = lds_at<T, V>(mine, d.row_b);
    Vec<T, V> th;
    node_from_vec<V>(vs.v, vc.v, th.v, cur.v);
    T* o = (T*)d.out_a;
    if (o) stg_tile<T, V>(o + g, th, live);
  }
#pragma unroll 1
  for (int k = nn - 2; k >= 0; --k) {
    const NodeDesc& d = p.d[k];
    const Vec<T, V> leaf = lds_at<T, V>(mine, CHAIN == 1 ? d.row_a : d.row_b);
    Vec<T, V> th, nxt;
    if constexpr (CHAIN == 1)
      node_from_vec<V, true, false>(cur.v, leaf.v, th.v, nxt.v);
    else
      node_from_vec<V, false, true>(leaf.v, cur.v, th.v, nxt.v);
    cur = nxt;
    T* o = (T*)d.out_a;
    if (o) stg_tile<T, V>(o + g, th, live);
  }
}

// create_hopf(q): the perfect binary tree — `c` nodes over `a` nodes, 2^q leaves.  Its walk is
// known at compile time, so it is spelled as a recursion over the program position instead of being
// interpreted: the subtree at position I spans 2^Q - 1 nodes, its cos child sits at I + 1 and its
// sin child at I + 2^(Q-1); descriptors become constant-bank operands at fixed offsets, the
// partial norms stay in registers (no parking, no kind switch, no moves at the switch's merge
// point), and the two sub-evaluations are independent chains the scheduler can interleave.
template <typename T, int V, int Q, int I>
__device__ __forceinline__ Vec<T, V> hopf_norm(const TileParams& p, const unsigned char* mine, int64_t g, uint32_t live) {
  const NodeDesc& d = p.d[I];
  Vec<T, V> th, nrm;
  if constexpr (Q == 1) {  // `a`: two leaves
    const Vec<T, V> vc = lds_at<T, V>(mine, d.row_a);
    const Vec<T, V> vs = lds_at<T, V>(mine, d.row_b);
    node_from_vec<V, false, false>(vs.v, vc.v, th.v, nrm.v);
  } else {  // `c`: two subtree norms (>= 0)
    const Vec<T, V> nc = hopf_norm<T, V, Q - 1, I + 1>(p, mine, g, live);
    const Vec<T, V> ns = hopf_norm<T, V, Q - 1, I + (1 << (Q - 1))>(p, mine, g, live);
    node_from_vec<V, true, true>(ns.v, nc.v, th.v, nrm.v);
  }
  T* o = (T*)d.out_a;
  if (o) stg_tile<T, V>(o + g, th, live);
  return nrm;
}

// =================================================================================================
// from_cartesian
// =================================================================================================
// CHAIN (the tree's shape class): 3 / 4 = create_hopf(2) / create_hopf(3), evaluated by hopf_norm;
// 0 = any tree (node kinds decoded per node); 1 = `b...ba` (create_standard), 2 = `b'...b'a`
// (create_standard_prime): every node but the last has one leaf operand and the running norm as the
// other, so the walk needs no kind switch, no operand selection, and the atan2 can skip the sign
// logic of the operand that is known to be a norm.
template <typename T, int V, int CHAIN>
__global__ void __launch_bounds__(32 + kGroupThreads * max_groups<T, V * (int)sizeof(T) / 16>(), 1)
    from_cartesian_tiled(const __grid_constant__ TileParams p, int64_t n_tiles, int n_groups, int n_bufs, int tail_points) {
  constexpr int P = kGroupThreads * V;
  constexpr int V0 = 16 / (int)sizeof(T), U = V / V0;
  constexpr uint32_t kTileRowBytes = (uint32_t)(kRowBytes * U);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PoolShared* sh = pool_init(smem_raw, n_groups, n_bufs);
  unsigned char* bufs = smem_raw + kPoolHeader;
  const int nn = p.n_nodes;
  const int n_rows = p.n_rows;
  const uint32_t buf_bytes = (uint32_t)n_rows * kTileRowBytes;
  const int tid = threadIdx.x;
  const int warp = warp_index_uniform();
  if (warp == 0) {
    producer_loop<T, U>(p, bufs, buf_bytes, n_groups, n_bufs, n_tiles, tail_points, sh);
    return;
  }
  const int t = (tid - 32) & (kGroupThreads - 1);
  T* const r_out = (T*)p.r_out;
  const int group = (warp - 1) / kGroupWarps;
  for (TileCursor at(group); at.tile < n_tiles; at.next(n_groups, n_bufs)) {
    unsigned char* mine = bufs + (size_t)at.b * buf_bytes + t * 16;
    US_ASSERT(group < n_groups && at.b >= 0 && at.b < n_bufs && at.slot >= 0 && at.slot < kGroupSlots);
    US_ASSERT(kPoolHeader + (size_t)(at.b + 1) * buf_bytes <= (size_t)kSmemBytes);
    mbar_wait(&sh->full[group][at.slot], at.parity);
    const int64_t g = at.tile * P + (int64_t)t * V0;
    const uint32_t live = live_chunks<T, V>(t, at.tile, n_tiles, tail_points);
    if (live) {
    if (r_out) {
      // leaves in c_nodes order, sequential, each square and each add rounded once
      Vec<T, V> acc;
      for (int l = 0; l < n_rows; ++l) {
        const Vec<T, V> x = lds_at<T, V>(mine, (uint32_t)l * kTileRowBytes);
        sumsq_step<V>(acc.v, x.v, l == 0);
      }
      sqrt_vec<V>(acc.v);
      stg_tile<T, V>(r_out + g, acc, live);
    }
    Vec<T, V> cur;
    if constexpr (CHAIN >= 3) {
      cur = hopf_norm<T, V, CHAIN - 1, 0>(p, mine, g, live);
    } else if constexpr (CHAIN != 0) {
      // fast paths unchecked; one range test per point at the end (us_common.cuh: chain_in_range)
      Vec<T, V> in_s, in_c;
      {
        const NodeDesc& d = p.d[nn - 1];  // the closing `a` node: two leaves
        in_c = lds_at<T, V>(mine, d.row_a);
        in_s = lds_at<T, V>(mine, d.row_b);
        Vec<T, V> th;
        node_from_fast_vec<V>(in_s.v, in_c.v, th.v, cur.v);
        T* o = (T*)d.out_a;
        if (o) stg_tile<T, V>(o + g, th, live);
      }
#pragma unroll 2
      for (int k = nn - 2; k >= 0; --k) {
        const NodeDesc& d = p.d[k];
        const Vec<T, V> leaf = lds_at<T, V>(mine, CHAIN == 1 ? d.row_a : d.row_b);
        Vec<T, V> th, nxt;
        if constexpr (CHAIN == 1)
          node_from_fast_vec<V, true, false>(cur.v, leaf.v, th.v, nxt.v);  // theta = atan2(norm, leaf)
        else
          node_from_fast_vec<V, false, true>(leaf.v, cur.v, th.v, nxt.v);  // theta = atan2(leaf, norm)
        cur = nxt;
        T* o = (T*)d.out_a;
        if (o) stg_tile<T, V>(o + g, th, live);
      }
      bool ok = true;
#pragma unroll
      for (int v = 0; v < V; ++v) ok = ok && chain_in_range(in_s.v[v], in_c.v[v], cur.v[v]);
      if (!ok) chain_redo<T, V, CHAIN>(p, mine, g, live);
    } else {
      // Any tree: one specialised body per node kind (warp-uniform switch), so that no operand has
      // to be selected at run time and the atan2 can drop the sign logic of operands that are norms.
#pragma unroll
      for (int v = 0; v < V; ++v) cur.v[v] = T(0);
      for (int k = nn - 1; k >= 0; --k) {
        const NodeDesc& d = p.d[k];
        const uint32_t kind = d.kind;
        Vec<T, V> th, nxt;
        switch (kind & kNodeKindMask) {
          case US_NODE_A: {  // two leaves; a finished sibling subtree waits, parked, for its `c` parent
            const Vec<T, V> vc = lds_at<T, V>(mine, d.row_a);
            const Vec<T, V> vs = lds_at<T, V>(mine, d.row_b);
            if (kind & kNodeLinked) sts_at<T, V>(mine, d.park, cur);
            node_from_vec<V, false, false>(vs.v, vc.v, th.v, nxt.v);
            break;
          }
          case US_NODE_B: {  // cos child leaf, sin child = running norm (>= 0)
            const Vec<T, V> vc = lds_at<T, V>(mine, d.row_a);
            node_from_vec<V, true, false>(cur.v, vc.v, th.v, nxt.v);
            break;
          }
          case US_NODE_BP: {  // sin child leaf, cos child = running norm (>= 0)
            const Vec<T, V> vs = lds_at<T, V>(mine, d.row_b);
            node_from_vec<V, false, true>(vs.v, cur.v, th.v, nxt.v);
            break;
          }
          default: {  // both children internal: cos = running norm, sin = the parked norm
            const Vec<T, V> vs = lds_at<T, V>(mine, d.park);
            node_from_vec<V, true, true>(vs.v, cur.v, th.v, nxt.v);
            break;
          }
        }
        cur = nxt;
        T* o = (T*)d.out_a;
        if (o) stg_tile<T, V>(o + g, th, live);
      }
    }
    }
    if (p.max_park) fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sh->empty[at.b]);
  }
}

// =================================================================================================
// host side: eligibility, variant selection, launch, ragged tail
// =================================================================================================
static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

template <typename T, int U>
static const void* variant(bool to, int chain) {
  constexpr int V = U * 16 / (int)sizeof(T);
  if (to) return (const void*)&to_cartesian_tiled<T, V>;
  if (chain == 1) return (const void*)&from_cartesian_tiled<T, V, 1>;
  if (chain == 2) return (const void*)&from_cartesian_tiled<T, V, 2>;
  // the compile-time hopf walk pays where the kernel is bound by instruction issue (float64: +12 %,
  // tools/ab_inproc.py); the float32 kernel sits on HBM and is faster interpreted (fewer registers)
  if constexpr (sizeof(T) == 8) {
    if (chain == 3) return (const void*)&from_cartesian_tiled<T, V, 3>;
    if (chain == 4) return (const void*)&from_cartesian_tiled<T, V, 4>;
  }
  return (const void*)&from_cartesian_tiled<T, V, 0>;
}

// does the program at [i, i + 2^q - 1) spell create_hopf(q): `c` over two copies of create_hopf(q-1), `a` at the bottom?
static bool is_hopf(const us_plan_t& plan, int i, int q) {
  if (q == 1) return node_kind(plan.node[i]) == US_NODE_A;
  return node_kind(plan.node[i]) == US_NODE_C && is_hopf(plan, i + 1, q - 1) && is_hopf(plan, i + (1 << (q - 1)), q - 1);
}

// shape class of the tree: 1: every node but the last is `b` and the last is `a`; 2: same with `b'`;
// 3 / 4: create_hopf(2) / create_hopf(3); 0: anything else
static int chain_kind(const us_plan_t& plan) {
  const int nn = plan.n_nodes;
  if (nn == 3 && is_hopf(plan, 0, 2)) return 3;
  if (nn == 7 && is_hopf(plan, 0, 3)) return 4;
  if (nn < 2 || plan.max_stack != 0 || node_kind(plan.node[nn - 1]) != US_NODE_A) return 0;
  const int first = node_kind(plan.node[0]);
  if (first != US_NODE_B && first != US_NODE_BP) return 0;
  for (int k = 1; k < nn - 1; ++k)
    if (node_kind(plan.node[k]) != first) return 0;
  return first == US_NODE_B ? 1 : 2;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// cudaFuncSetAttribute + the occupancy query cost a few microseconds each; their result only
// depends on (kernel, dynamic shared memory, device), so it is computed once and remembered.
int resident_ctas(const void* fn, size_t smem, int block_threads, int* per_sm) {
  struct Entry {
    const void* fn;
    size_t smem;
    int dev, threads, per_sm;
  };
  static Entry cache[128];
  static int used = 0;
  static std::mutex lock;
  int dev = 0;
  US_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  std::lock_guard<std::mutex> guard(lock);
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == fn && cache[i].smem == smem && cache[i].dev == dev && cache[i].threads == block_threads) {
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
    cache[used++] = Entry{fn, smem, dev, block_threads, *per_sm};
  } else {
    // table full (128 distinct kernel/tree-width pairs): keep working, just without the shortcut
    US_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), "cudaFuncSetAttribute(smem)");
  }
  return 0;
}

static int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return e ? atoi(e) : fallback;
}

// Tuning knobs (none is needed in normal use): environment variables at first use, us_set_tuning()
// at any time (tools/ab_inproc.py compares settings call by call inside one process, because the
// HBM-bound kernels drift by several percent over tens of milliseconds on one and the same box).
struct Tuning {
  int groups = env_int("US_TILE_GROUPS", 0);        // consumer groups per CTA (0: as many as fit)
  int bufs = env_int("US_TILE_BUFS", 0);            // buffers in the pool (0: as many as fit)
  int max_rows = env_int("US_TILE_MAX_ROWS", 13);   // wider trees take the row-ring kernels
  int min_tiles = env_int("US_TILED_MIN_TILES", 8); // per SM: smaller batches take the strided kernels
  int tile_u = env_int("US_TILE_U", 0);             // 16-byte chunks of a row per tile consumer thread (0: 2 for float64 where it fits, else 1)
  int ring_ctas = env_int("US_RING_CTAS", 0);       // resident CTAs per SM the ring kernels size their slots for (0: automatic)
  int ring_u = env_int("US_RING_U", 0);             // 16-byte chunks of a row per ring consumer thread (0: 2)
  int ring_slots = env_int("US_RING_SLOTS", 0);     // cap on the row slots of a ring (0: as many as fit)
};
static Tuning& tuning() {
  static Tuning t;
  return t;
}

// Cut the buffer pool to the tree: G groups computing, S = G + a little in flight.  On the HBM-bound
// float32 trees every buffer the consumers are not working on is a tile in flight, and DRAM runs
// measurably SLOWER under too many outstanding requests: create_standard(9) float32 with the pool
// filling the 227 KB (S = 11, G = 7: ~200 KB requested per SM, five times the latency-bandwidth
// product) ran at 0.86-0.89 of the copy peak right after an idle phase and 0.93-0.95 under sustained
// load; with S = 8 it runs at 0.92-0.98 and 0.92-0.98 (five boxes, tools/ab_inproc.py, AB_BENCH /
// AB_SETTLE modes; the shallower pools of 5-6 groups gain as much on that tree but lose 4-6 % on
// create_hopf(3) float32, whose interpreted walk needs the warps).  So: all the groups the register
// file holds, and 32 KB worth of buffers beyond them (at least one); the float64 trees are
// compute-bound and indifferent.  false: the tree is too wide for tiles.
static bool pool_geometry(int n_rows, int u, int g_max, int* n_groups, int* n_bufs) {
  const Tuning& t = tuning();
  if (n_rows > t.max_rows || n_rows > kTileCap + 1) return false;
  const int buf = n_rows * kRowBytes * u;
  int s_fit = (kSmemBytes - kPoolHeader) / buf;
  if (s_fit > kMaxBuffers) s_fit = kMaxBuffers;
  int extra = 32 * 1024 / buf;
  if (extra < 1) extra = 1;
  int g = g_max;
  if (g + extra > s_fit) g = s_fit - extra;
  if (t.groups > 0 && t.groups <= g_max) g = t.groups;
  int s = g + extra;
  if (t.bufs > 0) s = t.bufs;
  if (s > s_fit) s = s_fit;
  if (g < 1 || s <= g) return false;
  while ((s + g - 1) / g + 1 > kGroupSlots) --s;  // a group never has more tiles outstanding than barriers to announce them on
  *n_groups = g;
  *n_bufs = s;
  return true;
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
  // Two chunks per thread: float64 to_cartesian where the pool still holds four groups and a spare
  // buffer (trees of up to 11 rows at 4 KB a row): +10 % (create_hopf(3), create_standard_prime(8),
  // create_standard(9): 0.81-0.83 -> 0.91-0.93 of the HBM peak after idle, 0.78 -> 0.84-0.85
  // sustained).  float64 from_cartesian loses with it (96 registers spill, a third fewer warps:
  // -7 %) and float32 sits on HBM either way; the tile_u knob (1 / 2) forces the choice for A/B.
  int u = 1, n_groups = 0, n_bufs = 0;
  bool fits = false;
  const bool want_u2 = tuning().tile_u == 2 || (tuning().tile_u == 0 && TO && dtype == US_F64);
  if (want_u2 && (TO || dtype == US_F64)) {
    const int g2 = dtype == US_F32 ? max_groups<float, 2>() : max_groups<double, 2>();
    fits = pool_geometry(n_rows, 2, g2, &n_groups, &n_bufs) && (n_groups >= 4 || tuning().groups > 0);
    if (fits) u = 2;
  }
  if (!fits) fits = pool_geometry(n_rows, 1, dtype == US_F32 ? max_groups<float>() : max_groups<double>(), &n_groups, &n_bufs);
  int P = kGroupThreads * (16 / esz) * u;
  int64_t n_tiles = p.n / P;
  int64_t done = 0;
  if (!fits) {
    // wide tree: the rows stream through a ring instead of whole tiles being resident
    const int ring_p = ring_points_per_tile(dtype);
    if (p.n / ring_p < 2 * sm_count()) return kNotTaken;
    const int rc = launch_ring<TO>(p, dtype, st, &P);
    if (rc != 0) return rc;
    done = (p.n / P) * P;
  } else {
    // small batches: below a few tiles per group the pipeline's fill/drain costs more than it hides
    if (n_tiles < (int64_t)tuning().min_tiles * sm_count()) return kNotTaken;
    // a ragged end whose rows are a multiple of 16 bytes (always the case when the operands are rows
    // of one matrix) rides in the same launch as a partial last tile; otherwise the strided kernel below
    const int64_t rest = p.n - n_tiles * P;
    const int tail_points = (rest > 0 && (rest * esz) % 16 == 0) ? (int)rest : 0;
    if (tail_points) ++n_tiles;
    const int shape = chain_kind(p.plan);
    const void* fn = dtype == US_F32 ? (u == 2 ? (const void*)&to_cartesian_tiled<float, 8> : variant<float, 1>(TO, shape))
                                     : (u == 2 ? variant<double, 2>(TO, shape) : variant<double, 1>(TO, shape));
    const size_t smem = (size_t)kPoolHeader + (size_t)n_bufs * n_rows * kRowBytes * u;
    const int threads = 32 + kGroupThreads * n_groups;
    int per_sm = 0;
    if (int rc = resident_ctas(fn, smem, threads, &per_sm)) return rc;
    if (per_sm < 1) return kNotTaken;
    int64_t grid = sm_count();
    if (grid > n_tiles) grid = n_tiles;
    TileParams body;
    if (!build_walk<kTileCap>(p, TO, (uint32_t)(kRowBytes * u), 0u, body)) return kNotTaken;
    int64_t tiles = n_tiles;
    int tail = tail_points;
    void* args[] = {(void*)&body, (void*)&tiles, (void*)&n_groups, (void*)&n_bufs, (void*)&tail};
    US_CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3((unsigned)threads), args, smem, st), "tiled launch");
    done = tail_points ? p.n : n_tiles * P;
  }
  // ragged tail (< one tile) that could not ride along: the strided kernel on the remaining points
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

int ring_knob(int which) { return which == 0 ? tuning().ring_ctas : which == 1 ? tuning().ring_u : tuning().ring_slots; }

int set_tile_tuning(const char* name, int value) {
  Tuning& t = tuning();
  if (!strcmp(name, "tile_groups")) t.groups = value;
  else if (!strcmp(name, "tile_bufs")) t.bufs = value;
  else if (!strcmp(name, "tile_max_rows")) t.max_rows = value;
  else if (!strcmp(name, "tiled_min_tiles")) t.min_tiles = value;
  else if (!strcmp(name, "tile_u")) t.tile_u = (value == 1 || value == 2) ? value : 0;
  else if (!strcmp(name, "ring_ctas")) t.ring_ctas = value < 0 ? 0 : (value > 4 ? 4 : value);
  else if (!strcmp(name, "ring_u")) t.ring_u = (value == 1 || value == 2) ? value : 0;
  else if (!strcmp(name, "ring_slots")) t.ring_slots = value < 0 ? 0 : value;
  else return fail(US_EINVAL, "us_set_tuning: unknown knob '%s'", name);
  return 0;
}

int try_launch_to_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st) { return try_tiled<true>(p, dtype, st); }
int try_launch_from_cartesian_tiled(const XformParams& p, int dtype, cudaStream_t st) { return try_tiled<false>(p, dtype, st); }

}  // namespace us
