// Weighted tensor-product reductions behind integrate()
// (reference: src/ultrasphere/_integral.py:245-276).
//
//   dense     out[m] = sum_grid  prod_i w_i[idx_i] * val[idx..., m]
//   separable out_a[m] = sum_j w_a[j] * val_a[j, m]          for every axis a in one launch
//
// Bound: HBM read of `val` (8 B per quadrature node per trailing element for f64); the 1-D weight
// tables (<= a few hundred entries per axis) stay in L1.  Accumulation is float64 throughout.
// Two phases, no atomics: per-block partials (warp shuffle, then a shared-memory tree over the
// warps) are written to caller-provided scratch and a second, single-block kernel adds them in a
// fixed order, so the result is bit-reproducible for a given shape.
#include "us_common.cuh"
#include "us_params.cuh"
#include "us_tma.cuh"

namespace us {

constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;
constexpr int kMaxRedBlocks = 148 * 4;  // one wave at 4 resident CTAs per SM

struct ReduceParams {
  const void* val;
  const void* w[US_MAX_DIMS];
  int64_t extent[US_MAX_DIMS];
  int32_t ndim;
  int32_t m;        // trailing elements per node
  int32_t lanes;    // rows kernel: active lanes per warp = largest multiple of m that is <= 32
  int32_t nseg;     // segments per innermost-axis row
  int64_t rows;     // product of extent[0..ndim-1)
  int64_t inner;    // extent[ndim-1]
  int64_t items;    // rows * nseg
  double* partial;  // [blocks][m]
};

template <typename T>
__device__ __forceinline__ double ldw(const void* w, int64_t i) {
  return (double)__ldg((const T*)w + i);
}

// Odometer over the outer axes: the warp-uniform weight prod_{i<ndim-1} w_i[idx_i] of a grid row.
template <typename T>
struct RowWalker {
  int64_t idx[US_MAX_DIMS];
  const ReduceParams& p;
  __device__ __forceinline__ RowWalker(const ReduceParams& pp, int64_t row) : p(pp) {
    int64_t rem = row;
    for (int d = p.ndim - 2; d >= 0; --d) {
      const int64_t e = p.extent[d];
      const int64_t q = rem / e;
      idx[d] = rem - q * e;
      rem = q;
    }
  }
  __device__ __forceinline__ double weight() const {
    double w = 1.0;
    for (int d = 0; d < p.ndim - 1; ++d) w *= ldw<T>(p.w[d], idx[d]);
    return w;
  }
  __device__ __forceinline__ void next() {
    for (int d = p.ndim - 2; d >= 0; --d) {
      if (++idx[d] < p.extent[d]) break;
      idx[d] = 0;
    }
  }
};

constexpr int kRowGroup = 4;  // rows a warp keeps in flight
constexpr int kChunk = 4;     // strided loads per lane and row in flight (kRowGroup*kChunk loads)

// Rows kernel (m <= 32).  A warp owns a contiguous range of grid rows.  The innermost run of a
// row (inner*m contiguous values) is read by `lanes` lanes with stride `lanes`; because `lanes` is
// a multiple of m, a lane always sees the same trailing index and the inner index advances by
// lanes/m per step: one accumulator per lane, no divisions.  Loads of kRowGroup rows x kChunk
// strides are issued before any is consumed (memory-level parallelism: 16 x 256 B per warp), the
// innermost weights are loaded once per stride and shared by the rows of the group.
template <typename T>
__global__ void __launch_bounds__(kRedThreads) wreduce_rows(const __grid_constant__ ReduceParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * kRedWarps;
  const int64_t gw = (int64_t)blockIdx.x * kRedWarps + warp;
  const int64_t per = (p.rows + nwarps - 1) / nwarps;
  int64_t row = gw * per;
  const int64_t row_end = min(p.rows, row + per);
  const int m = p.m, W = p.lanes;
  const int lw = W / m;             // inner-index advance per stride
  const int l0 = lane / m;          // this lane's inner index in stride 0
  const bool active = lane < W;
  const int64_t run = p.inner * m;  // contiguous values per row
  const T* val = (const T*)p.val;
  const void* wl = p.w[p.ndim - 1];
  double acc = 0.0;
  if (row < row_end) {
    RowWalker<T> walk(p, row);
    for (; row < row_end; row += kRowGroup) {
      double wrow[kRowGroup];
      const T* base[kRowGroup];
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r) {
        const bool live = row + r < row_end;
        wrow[r] = live ? walk.weight() : 0.0;
        base[r] = val + (live ? row + r : row) * run;
        if (live) walk.next();
      }
      double a[kRowGroup];
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r) a[r] = 0.0;
      int64_t lb = l0;  // inner index of this lane's first element in the current chunk
      for (int64_t e0 = 0; e0 < run; e0 += (int64_t)kChunk * W, lb += (int64_t)kChunk * lw) {
        T v[kRowGroup][kChunk];
        double wv[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
          const int64_t e = e0 + (int64_t)j * W + lane;
          const bool ok = active && e < run;
#pragma unroll
          for (int r = 0; r < kRowGroup; ++r) v[r][j] = ok ? ld_stream(base[r] + e) : T(0);
          wv[j] = ok ? ldw<T>(wl, lb + (int64_t)j * lw) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
#pragma unroll
          for (int r = 0; r < kRowGroup; ++r) a[r] = fma(wv[j], (double)v[r][j], a[r]);
      }
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r) acc = fma(wrow[r], a[r], acc);
    }
  }
  // lanes congruent mod m hold the same trailing index: shared-memory tree over lanes and warps
  __shared__ double sm[kRedWarps][32];
  sm[warp][lane] = active ? acc : 0.0;
  __syncthreads();
  if (threadIdx.x < m) {
    double t = 0.0;
    for (int wv = 0; wv < kRedWarps; ++wv)
      for (int l = threadIdx.x; l < W; l += m) t += sm[wv][l];
    p.partial[(int64_t)blockIdx.x * m + threadIdx.x] = t;
  }
}

// Column kernel (m > 32): threads across the trailing index (coalesced), grid.y splits the grid
// nodes; the node weight prod_i w_i[idx_i] is warp-uniform and advanced with an odometer, four
// nodes' loads are in flight per thread.
template <typename T>
__global__ void __launch_bounds__(kRedThreads) wreduce_cols(const __grid_constant__ ReduceParams p,
                                                            int64_t total_nodes, int64_t m64) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (total_nodes + gridDim.y - 1) / gridDim.y;
  int64_t g = (int64_t)blockIdx.y * per;
  const int64_t g1 = min(total_nodes, g + per);
  const bool live_col = col < m64;
  const T* val = (const T*)p.val;
  double acc = 0.0;
  if (g < g1) {
    // odometer over ALL axes (the trailing index is the fastest one in memory)
    int64_t idx[US_MAX_DIMS];
    {
      int64_t rem = g;
      for (int d = p.ndim - 1; d >= 0; --d) {
        const int64_t e = p.extent[d];
        const int64_t q = rem / e;
        idx[d] = rem - q * e;
        rem = q;
      }
    }
    for (; g < g1; g += 4) {
      double wg[4];
      T v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool live = g + r < g1;
        double w = live ? 1.0 : 0.0;
        if (live) {
          for (int d = 0; d < p.ndim; ++d) w *= ldw<T>(p.w[d], idx[d]);
          for (int d = p.ndim - 1; d >= 0; --d) {
            if (++idx[d] < p.extent[d]) break;
            idx[d] = 0;
          }
        }
        wg[r] = w;
        v[r] = (live && live_col) ? ld_stream(val + (g + r) * m64 + col) : T(0);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc = fma(wg[r], (double)v[r], acc);
    }
  }
  if (live_col) p.partial[(int64_t)blockIdx.y * m64 + col] = acc;
}

// TMA-streamed rows kernel (m <= 32, the hot case: real / complex scalar integrands).  `val` is one
// contiguous array, so a persistent CTA's producer warp streams it through a ring of shared-memory
// slots with ONE bulk copy per slab of whole grid rows (cp.async.bulk, mbarrier completion): the
// bytes in flight per SM are set by the ring (4 x 32 KB per CTA), not by registers or occupancy.
// Consumer warps take the rows of a slab round-robin; a row's outer weight needs the row index
// decomposed once per row (warp-uniform), the lanes then stride the row in shared memory.
constexpr int kRedSlots = 4;
constexpr int kRedSlotBytes = 24 * 1024;  // 4 slots x 24 KB (+ weights) per CTA: two CTAs per SM
constexpr int kRedMaxRows = 256;          // rows per slab (their outer weights live next to the slot)

struct RedRing {
  uint64_t full[kRedSlots];
  uint64_t empty[kRedSlots];
  double wrow[kRedSlots][kRedMaxRows];  // outer weight of every row of the slab in each slot
};

template <typename T>
__global__ void __launch_bounds__(kRedThreads + 32, 2) wreduce_tma(const __grid_constant__ ReduceParams p, int rows_per_slab,
                                                                     int64_t n_slabs) {
  extern __shared__ __align__(128) unsigned char red_smem[];
  T* ring = reinterpret_cast<T*>(red_smem);
  RedRing* sh = reinterpret_cast<RedRing*>(red_smem + (size_t)kRedSlots * kRedSlotBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRedSlots; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], kRedWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int m = p.m, W = p.lanes;
  const int64_t run = p.inner * m;  // contiguous values per row
  __shared__ double sm[kRedWarps][32];
  if (warp == kRedWarps) {
    // producer warp: while the slot drains, its lanes compute the slab's row weights (index
    // decomposition with divisions — off the consumers' critical path), then one lane posts the
    // byte count and issues the slab's single bulk copy
    const uint64_t policy = policy_evict_first();
    uint32_t q = 0;
    for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x, ++q) {
      const int slot = (int)(q % kRedSlots);
      if (q >= (uint32_t)kRedSlots) mbar_wait(&sh->empty[slot], ((q / kRedSlots) - 1u) & 1u);
      const int64_t row0 = slab * rows_per_slab;
      const int nrows = (int)min((int64_t)rows_per_slab, p.rows - row0);
      for (int r = lane; r < nrows; r += 32) {
        int64_t rem = row0 + r;
        double wr = 1.0;
        for (int d = p.ndim - 2; d >= 0; --d) {
          const int64_t e = p.extent[d];
          const int64_t qd = rem / e;
          wr *= ldw<T>(p.w[d], rem - qd * e);
          rem = qd;
        }
        sh->wrow[slot][r] = wr;
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)((int64_t)nrows * run * (int64_t)sizeof(T));
        mbar_arrive_expect_tx(&sh->full[slot], bytes);  // release: the weights above are visible to waiters
        tma_load_row(reinterpret_cast<unsigned char*>(ring) + (size_t)slot * kRedSlotBytes, (const T*)p.val + row0 * run, bytes,
                     &sh->full[slot], policy);
      }
    }
  } else {
    const int lw = W / m, l0 = lane / m;
    const bool active = lane < W;
    const void* wl = p.w[p.ndim - 1];
    double acc = 0.0;
    uint32_t q = 0;
    for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x, ++q) {
      const int slot = (int)(q % kRedSlots);
      mbar_wait(&sh->full[slot], (q / kRedSlots) & 1u);
      const T* base = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(ring) + (size_t)slot * kRedSlotBytes);
      const int nrows = (int)min((int64_t)rows_per_slab, p.rows - slab * rows_per_slab);
      for (int r = warp; r < nrows; r += kRedWarps) {
        const T* row = base + (int64_t)r * run;
        double a0 = 0.0, a1 = 0.0;
        int64_t e = lane, l = l0;
        for (; e + W < run; e += 2 * W, l += 2 * lw) {
          if (active) {
            a0 = fma(ldw<T>(wl, l), (double)row[e], a0);
            a1 = fma(ldw<T>(wl, l + lw), (double)row[e + W], a1);
          }
        }
        if (active && e < run) a0 = fma(ldw<T>(wl, l), (double)row[e], a0);
        acc = fma(sh->wrow[slot][r], a0 + a1, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[slot]);
    }
    sm[warp][lane] = active ? acc : 0.0;
  }
  __syncthreads();
  if (tid < m) {
    double t = 0.0;
    for (int wv = 0; wv < kRedWarps; ++wv)
      for (int l = tid; l < W; l += m) t += sm[wv][l];
    p.partial[(int64_t)blockIdx.x * m + tid] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) wreduce_final(const double* __restrict__ partial, int nblocks,
                                                     int64_t m, T* __restrict__ out, int accumulate) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= m) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partial[(int64_t)b * m + col];
  if (accumulate) t += (double)out[col];
  out[col] = (T)t;
}

struct SepParams {
  const void* w[US_MAX_AXES];
  const void* val[US_MAX_AXES];
  void* out[US_MAX_AXES];
  int64_t len[US_MAX_AXES];
  int64_t m[US_MAX_AXES];
};

template <typename T>
__global__ void __launch_bounds__(256) separable_reduce(const __grid_constant__ SepParams p) {
  const int a = blockIdx.x;
  const int64_t m = p.m[a], len = p.len[a];
  const T* val = (const T*)p.val[a];
  const T* w = (const T*)p.w[a];
  T* out = (T*)p.out[a];
  if (m == 1) {
    if (blockIdx.y > 0) return;  // uniform per block
    // one short dot product: whole block, shuffle + shared tree
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < len; j += blockDim.x) acc = fma((double)w[j], (double)val[j], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[k];
      out[0] = (T)t;
    }
    return;
  }
  for (int64_t col = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; col < m;
       col += (int64_t)gridDim.y * blockDim.x) {
    double acc = 0.0;
    for (int64_t j = 0; j < len; ++j) acc = fma((double)w[j], (double)val[j * m + col], acc);
    out[col] = (T)acc;
  }
}

static inline bool rows_ok(int64_t m) { return m >= 1 && m <= 32; }

static int reduce_blocks(int64_t m, int64_t total_nodes, int64_t rows, int* gx, int* gy) {
  if (rows_ok(m)) {
    int64_t b = (rows + (int64_t)kRedWarps * kRowGroup - 1) / ((int64_t)kRedWarps * kRowGroup);
    *gx = (int)(b < 1 ? 1 : (b > kMaxRedBlocks ? kMaxRedBlocks : b));
    *gy = 1;
    return *gx;
  }
  int64_t cx = (m + kRedThreads - 1) / kRedThreads;
  int64_t want = kMaxRedBlocks / cx;
  if (want < 1) want = 1;
  if (want > total_nodes) want = total_nodes < 1 ? 1 : total_nodes;
  *gx = (int)cx;
  *gy = (int)want;
  return (int)want;
}

}  // namespace us

using namespace us;

extern "C" int64_t us_weighted_reduce_scratch_bytes(int64_t m) {
  if (m < 1) return 0;
  // worst case number of partial rows is kMaxRedBlocks for either kernel
  return (int64_t)kMaxRedBlocks * m * (int64_t)sizeof(double);
}

extern "C" int us_weighted_reduce(int dtype, int ndim, const int64_t* extent, const void* const* w_dev,
                                  const void* val_dev, int64_t m, void* out_dev, int accumulate,
                                  void* scratch_dev, int64_t scratch_bytes, void* stream) {
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "us_weighted_reduce: bad dtype %d", dtype);
  if (ndim < 1 || ndim > US_MAX_DIMS) return fail(US_ECAPACITY, "us_weighted_reduce: ndim %d not in 1..%d", ndim, US_MAX_DIMS);
  if (!extent || !w_dev || !val_dev || !out_dev || !scratch_dev || m < 1)
    return fail(US_EINVAL, "us_weighted_reduce: null pointer or m < 1");
  if (scratch_bytes < us_weighted_reduce_scratch_bytes(m))
    return fail(US_EINVAL, "us_weighted_reduce: scratch too small (%lld < %lld)", (long long)scratch_bytes,
                (long long)us_weighted_reduce_scratch_bytes(m));
  ReduceParams p{};
  p.val = val_dev;
  p.ndim = ndim;
  int64_t total = 1;
  for (int d = 0; d < ndim; ++d) {
    if (extent[d] < 1 || !w_dev[d]) return fail(US_EINVAL, "us_weighted_reduce: axis %d empty or null weights", d);
    p.extent[d] = extent[d];
    p.w[d] = w_dev[d];
    total *= extent[d];
  }
  p.inner = extent[ndim - 1];
  p.rows = total / p.inner;
  p.nseg = 1;
  p.items = p.rows;
  p.partial = (double*)scratch_dev;
  cudaStream_t st = (cudaStream_t)stream;
  int gx, gy;
  const int nparts = reduce_blocks(m, total, p.rows, &gx, &gy);
  int nparts_used = nparts;
  const int esz = dtype == US_F32 ? 4 : 8;
  const int64_t run_bytes = p.inner * m * esz;
  if (rows_ok(m) && run_bytes <= kRedSlotBytes / 2 && total * m * esz >= (8 << 20) &&
      (reinterpret_cast<uintptr_t>(val_dev) & 15u) == 0) {
    // TMA-streamed path: slabs of whole rows; a slab's byte count must be a multiple of 16
    p.m = (int32_t)m;
    p.lanes = (int32_t)((32 / m) * m);
    int rows_per_slab = (int)(kRedSlotBytes / run_bytes);
    if (rows_per_slab > kRedMaxRows) rows_per_slab = kRedMaxRows;
    while (rows_per_slab > 1 && (rows_per_slab * run_bytes) % 16 != 0) --rows_per_slab;
    if ((rows_per_slab * run_bytes) % 16 == 0 && (p.rows % rows_per_slab == 0 || ((p.rows % rows_per_slab) * run_bytes) % 16 == 0)) {
      const int64_t n_slabs = (p.rows + rows_per_slab - 1) / rows_per_slab;
      int dev = 0, sms = 148;
      US_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
      US_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
      const void* fn = dtype == US_F32 ? (const void*)&wreduce_tma<float> : (const void*)&wreduce_tma<double>;
      const size_t smem = (size_t)kRedSlots * kRedSlotBytes + sizeof(RedRing);
      int per_sm = 0;
      if (int rc = resident_ctas(fn, smem, kRedThreads + 32, &per_sm)) return rc;
      if (per_sm >= 1) {
        int64_t grid = (int64_t)sms * per_sm;
        if (grid > n_slabs) grid = n_slabs;
        if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
        ReduceParams body = p;
        int rps = rows_per_slab;
        int64_t ns = n_slabs;
        void* args[] = {(void*)&body, (void*)&rps, (void*)&ns};
        US_CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kRedThreads + 32), args, smem, st), "wreduce_tma launch");
        nparts_used = (int)grid;
        goto final_pass;
      }
    }
  }
  if (rows_ok(m)) {
    p.m = (int32_t)m;
    p.lanes = (int32_t)((32 / m) * m);
    if (dtype == US_F32)
      wreduce_rows<float><<<gx, kRedThreads, 0, st>>>(p);
    else
      wreduce_rows<double><<<gx, kRedThreads, 0, st>>>(p);
  } else {
    dim3 grid(gx, gy);
    if (dtype == US_F32)
      wreduce_cols<float><<<grid, kRedThreads, 0, st>>>(p, total, m);
    else
      wreduce_cols<double><<<grid, kRedThreads, 0, st>>>(p, total, m);
  }
  US_CUDA_TRY(cudaGetLastError(), "weighted_reduce launch");
final_pass:
  const int fb = (int)((m + 255) / 256);
  if (dtype == US_F32)
    wreduce_final<float><<<fb, 256, 0, st>>>(p.partial, nparts_used, m, (float*)out_dev, accumulate);
  else
    wreduce_final<double><<<fb, 256, 0, st>>>(p.partial, nparts_used, m, (double*)out_dev, accumulate);
  US_CUDA_TRY(cudaGetLastError(), "weighted_reduce final launch");
  return 0;
}

extern "C" int us_separable_reduce(int dtype, int n_axes, const int64_t* len, const int64_t* m,
                                   const void* const* w_dev, const void* const* val_dev,
                                   void* const* out_dev, void* stream) {
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "us_separable_reduce: bad dtype %d", dtype);
  if (n_axes < 1 || n_axes > US_MAX_AXES)
    return fail(US_ECAPACITY, "us_separable_reduce: n_axes %d not in 1..%d", n_axes, US_MAX_AXES);
  if (!len || !m || !w_dev || !val_dev || !out_dev) return fail(US_EINVAL, "us_separable_reduce: null table");
  SepParams p{};
  int64_t mmax = 1;
  for (int a = 0; a < n_axes; ++a) {
    if (len[a] < 1 || m[a] < 1 || !w_dev[a] || !val_dev[a] || !out_dev[a])
      return fail(US_EINVAL, "us_separable_reduce: axis %d has an empty extent or a null pointer", a);
    p.len[a] = len[a];
    p.m[a] = m[a];
    p.w[a] = w_dev[a];
    p.val[a] = val_dev[a];
    p.out[a] = out_dev[a];
    if (m[a] > mmax) mmax = m[a];
  }
  int64_t gy = (mmax + 255) / 256;
  if (gy > 1024) gy = 1024;
  dim3 grid(n_axes, (unsigned)gy);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == US_F32)
    separable_reduce<float><<<grid, 256, 0, st>>>(p);
  else
    separable_reduce<double><<<grid, 256, 0, st>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "separable_reduce launch");
  return 0;
}
