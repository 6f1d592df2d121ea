// Uniform sampling in the unit ball / on the unit sphere, on the device
// (reference: src/ultrasphere/_random.py:16-57 and :96-192, which draw with NumPy on the host
// and copy c_ndim * N values over PCIe).
//
//   g ~ N(0, 1/2)^dim ;  ball:  x = g / sqrt(|g|^2 + z), z ~ Exp(1)     (Barthe et al. 2005)
//                        sphere: x = g / |g|
//
// Counter-based Philox4x32-10: point i, block b of its variates uses counter (i_lo, i_hi, b, stream)
// under key = seed, so the result depends only on (seed, i) — not on the launch geometry, the GPU
// count or the batch split.  A point's variates are generated twice (norm pass, write pass) instead
// of being stored: dim can be 257 and the kernel is bound by the HBM write of dim*sizeof(T) bytes.
#include "us_common.cuh"

namespace us {

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
  }
  __device__ __forceinline__ void operator()(uint32_t (&c)[4]) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, a, b);
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
  }
};

// two uniforms in (0,1] x [0,1) from four 32-bit words (53-bit resolution: same code for both dtypes)
__device__ __forceinline__ void uniforms(const uint32_t (&w)[4], double& u1, double& u2) {
  const uint64_t a = ((uint64_t)w[0] << 32) | w[1], b = ((uint64_t)w[2] << 32) | w[3];
  u1 = ((double)(a >> 11) + 1.0) * 0x1.0p-53;  // (0, 1]
  u2 = (double)(b >> 11) * 0x1.0p-53;          // [0, 1)
}

// Box–Muller normals with variance 1/2: sqrt(-ln u1) * (cos, sin)(2 pi u2).
// float64: one Philox block -> two 53-bit uniforms -> 2 normals.
__device__ __forceinline__ void normals(const Philox& rng, int64_t i, uint32_t block, double (&g)[2]) {
  uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), block, 0u};
  rng(c);
  double u1, u2;
  uniforms(c, u1, u2);
  const double rad = sqrt(-log(u1));
  double s, co;
  sincospi(2.0 * u2, &s, &co);
  g[0] = rad * co;
  g[1] = rad * s;
}
// float32: one Philox block -> four 24-bit uniforms -> 4 normals, single-precision math.
__device__ __forceinline__ void normals(const Philox& rng, int64_t i, uint32_t block, float (&g)[4]) {
  uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), block, 0u};
  rng(c);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * 0x1.0p-24f;  // (0, 1]
    const float u2 = (float)(c[2 * h + 1] >> 8) * 0x1.0p-24f;       // [0, 1)
    const float rad = sqrtf(-logf(u1));
    float s, co;
    sincospif(2.0f * u2, &s, &co);
    g[2 * h] = rad * co;
    g[2 * h + 1] = rad * s;
  }
}

struct BallParams {
  void* out[US_MAX_LEAVES];
  int64_t n;
  int32_t dim;
  int32_t surface;
  uint32_t k0, k1;
};

// T: output / math type; G = normals per Philox block (2 for double, 4 for float).  Up to kKeep
// normals stay in registers; wider points regenerate their variates in the write pass.
template <typename T, int G>
__global__ void __launch_bounds__(256) random_ball_kernel(const __grid_constant__ BallParams p) {
  constexpr int kKeep = 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const Philox rng{p.k0, p.k1};
  const int blocks = (p.dim + G - 1) / G;
  const bool keep = p.dim <= kKeep;
  T kept[kKeep];
  T q = T(0);
  if (keep) {
#pragma unroll
    for (int b = 0; b < kKeep / G; ++b) {
      if (b < blocks) {
        T g[G];
        normals(rng, i, (uint32_t)b, g);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          kept[b * G + j] = g[j];
          if (b * G + j < p.dim) q += g[j] * g[j];
        }
      }
    }
  } else {
    for (int b = 0; b < blocks; ++b) {
      T g[G];
      normals(rng, i, (uint32_t)b, g);
#pragma unroll
      for (int j = 0; j < G; ++j)
        if (b * G + j < p.dim) q += g[j] * g[j];
    }
  }
  if (!p.surface) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), 0u, 1u};  // stream 1: the exponential
    rng(c);
    double u1, u2;
    uniforms(c, u1, u2);
    q += (T)(-log(u1));
  }
  const T inv = q > T(0) ? (T)(1.0 / sqrt((double)q)) : T(0);  // reference: nan_to_num(..., nan=0)
  if (keep) {
#pragma unroll
    for (int d = 0; d < kKeep; ++d)
      if (d < p.dim) __stcs((T*)p.out[d] + i, kept[d] * inv);
  } else {
    for (int b = 0; b < blocks; ++b) {
      T g[G];
      normals(rng, i, (uint32_t)b, g);
#pragma unroll
      for (int j = 0; j < G; ++j)
        if (b * G + j < p.dim) __stcs((T*)p.out[b * G + j] + i, g[j] * inv);
    }
  }
}

struct UniformParams {
  void* out[US_MAX_LEAVES];
  double lo[US_MAX_LEAVES], span[US_MAX_LEAVES];
  int64_t n;
  int32_t count;
  uint32_t k0, k1;
};

template <typename T>
__global__ void __launch_bounds__(256) random_uniform_kernel(const __grid_constant__ UniformParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const Philox rng{p.k0, p.k1};
  for (int b = 0; 2 * b < p.count; ++b) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)b, 2u};  // stream 2: uniforms
    rng(c);
    double u1, u2;
    uniforms(c, u1, u2);
    __stcs((T*)p.out[2 * b] + i, (T)fma(1.0 - u1, p.span[2 * b], p.lo[2 * b]));  // 1-u1 in [0,1)
    if (2 * b + 1 < p.count) __stcs((T*)p.out[2 * b + 1] + i, (T)fma(u2, p.span[2 * b + 1], p.lo[2 * b + 1]));
  }
}

}  // namespace us

using namespace us;

extern "C" int us_random_ball(int dtype, int dim, int64_t n, uint64_t seed, int surface, void* const* out_dev, void* stream) {
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "us_random_ball: bad dtype %d", dtype);
  if (dim < 1 || dim > US_MAX_LEAVES) return fail(US_ECAPACITY, "us_random_ball: dim %d not in 1..%d", dim, US_MAX_LEAVES);
  if (n < 0 || !out_dev) return fail(US_EINVAL, "us_random_ball: bad arguments");
  if (n == 0) return 0;
  BallParams p{};
  for (int d = 0; d < dim; ++d) {
    if (!out_dev[d]) return fail(US_EINVAL, "us_random_ball: output %d is null", d);
    p.out[d] = out_dev[d];
  }
  p.n = n;
  p.dim = dim;
  p.surface = surface ? 1 : 0;
  p.k0 = (uint32_t)seed;
  p.k1 = (uint32_t)(seed >> 32);
  const int64_t blocks = (n + 255) / 256;
  if (blocks > 0x7FFFFFFFLL) return fail(US_EINVAL, "us_random_ball: n too large for one launch");
  if (dtype == US_F32)
    random_ball_kernel<float, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else
    random_ball_kernel<double, 2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "random_ball launch");
  return 0;
}

extern "C" int us_random_uniform(int dtype, int count, int64_t n, uint64_t seed, const double* lo, const double* hi,
                                 void* const* out_dev, void* stream) {
  if (dtype != US_F32 && dtype != US_F64) return fail(US_EINVAL, "us_random_uniform: bad dtype %d", dtype);
  if (count < 1 || count > US_MAX_LEAVES) return fail(US_ECAPACITY, "us_random_uniform: count %d not in 1..%d", count, US_MAX_LEAVES);
  if (n < 0 || !out_dev || !lo || !hi) return fail(US_EINVAL, "us_random_uniform: bad arguments");
  if (n == 0) return 0;
  UniformParams p{};
  for (int d = 0; d < count; ++d) {
    if (!out_dev[d]) return fail(US_EINVAL, "us_random_uniform: output %d is null", d);
    p.out[d] = out_dev[d];
    p.lo[d] = lo[d];
    p.span[d] = hi[d] - lo[d];
  }
  p.n = n;
  p.count = count;
  p.k0 = (uint32_t)seed;
  p.k1 = (uint32_t)(seed >> 32);
  const int64_t blocks = (n + 255) / 256;
  if (blocks > 0x7FFFFFFFLL) return fail(US_EINVAL, "us_random_uniform: n too large for one launch");
  if (dtype == US_F32)
    random_uniform_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else
    random_uniform_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "random_uniform launch");
  return 0;
}
