// PTX wrappers shared by the TMA-staged kernels: mbarrier, cp.async.bulk, L2 policies, V-wide
// shared-memory loads and streaming global stores.
#pragma once

#include "us_common.cuh"

namespace us {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for a phase, called by ALL lanes of a converged warp.  The loop condition is a warp vote, so
// the warp leaves the loop as one: with barrier address and parity in uniform registers ptxas
// treats the predicate of a plain `while (!try_wait)` as warp-uniform, keeps the surrounding loop
// state (tile cursor, buffer index, parity) in uniform registers and drops the __syncwarp() before
// the release — but the hardware may answer the lanes of one try_wait differently when the phase
// completes while the instruction is in flight, and a warp that splits there runs the tile twice
// on shared uniform registers (seen as hangs and launch failures at 6 * 10^7 points and more).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
  }
}
// The same on shared-space (32-bit) addresses computed once per kernel: the generic-pointer forms
// above rebuild the shared window base (S2R CgaCtaId + LEA + ...) at every call site, which the
// row-ring kernels pay per row and per thread.
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  while (!__all_sync(0xffffffffu, mbar_try_wait_s(bar, parity))) {
  }
}
// Wait by ONE thread (the elected lane of a producer warp, a single-thread role): no vote.
__device__ __forceinline__ void mbar_wait_one(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp index and leader election in the forms ptxas recognises as warp-uniform: the producer loops
// then keep their tile counters, buffer indices and addresses in uniform registers and issue each
// bulk copy with three or four instructions.  (With a lane picked by `threadIdx.x & 31` the same
// loop ran in vector registers: every UBLKCP was wrapped in R2UR moves and an ELECT retry loop,
// ~17 instructions of a serial chain per row, and ncu showed the consumer warps waiting 10-23 % of
// their time for tiles one starved producer thread had not issued yet.)
__device__ __forceinline__ int warp_index_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// global -> shared bulk copy through the TMA unit; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// ---- V-wide register tiles ------------------------------------------------------------------
template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lds_vec(const T* row, int tid) {
  return *reinterpret_cast<const Vec<T, V>*>(row + tid * V);
}

// 16-byte (or narrower) load from a shared-space address
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lds_vec_s(uint32_t addr) {
  Vec<T, V> out;
  if constexpr (sizeof(T) * V == 16) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    uint32_t w[4] = {a, b, c, d};
    memcpy(&out, w, 16);
  } else if constexpr (sizeof(T) * V == 8) {
    uint32_t a, b;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
    uint32_t w[2] = {a, b};
    memcpy(&out, w, 8);
  } else {
    uint32_t a;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(addr));
    memcpy(&out, &a, 4);
  }
  return out;
}

template <typename T, int V>
__device__ __forceinline__ void stg_vec(T* dst, const Vec<T, V>& x) {
  if constexpr (sizeof(T) * V == 16) {
    const float4 f = *reinterpret_cast<const float4*>(&x);
    __stcs(reinterpret_cast<float4*>(dst), f);
  } else if constexpr (sizeof(T) * V == 8) {
    const float2 f = *reinterpret_cast<const float2*>(&x);
    __stcs(reinterpret_cast<float2*>(dst), f);
  } else {
    __stcs(reinterpret_cast<float*>(dst), *reinterpret_cast<const float*>(&x));
  }
}

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

}  // namespace us
