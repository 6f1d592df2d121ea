// 1-D quadrature rules of every angle of a coordinate system, built on the device in one launch
// (reference: the per-node loop of roots(), src/ultrasphere/_integral.py:83-109, which calls
// scipy.special.roots_jacobi on the host — Golub–Welsch eigenvalues of the Jacobi matrix plus
// one Newton step, scipy/special/_orthogonal.py:159-205 — and then copies the tables to the
// device).
//
// The nodes are the eigenvalues of the same symmetric tridiagonal Jacobi matrix
//   diag a_k = (b^2-a^2)/((2k+a+b)(2k+a+b+2)),  offdiag b_k = 2/(2k+a+b) sqrt(k(k+a)(k+b)(k+a+b)/((2k+a+b)^2-1))
// but they are found root by root, which suits a GPU: ONE WARP PER ROOT.
//   1. Sturm multisection: the orthonormal polynomials p_0..p_n of the three-term recurrence
//      form a Sturm sequence, the number of sign changes at x = the number of roots above x.
//      The 32 lanes test 32 interior points of the bracket at once, a ballot keeps the
//      sub-bracket holding root j: 6 passes shrink [-1,1] to 2/33^6 = 1.5e-9.
//   2. Newton on p_n with the differentiated recurrence (3 steps, quadratic from 1e-9).
//   3. Christoffel weight w_j = mu0 / sum_{k<n} p_k(x_j)^2 — a sum of squares, no cancellation,
//      no gamma functions beyond mu0 = 2^(a+b+1) B(a+1,b+1), which the host computes.
// Against 50-digit arithmetic the nodes are within 0.6 ulp and the weights within 2e-12 relative
// up to n = 512 (scipy's own: 1.4 ulp and 1.5e-9); see tests/test_gpu_rules.py.
// For a = b only the non-negative roots are solved and mirrored (exact antisymmetry, as scipy's
// symmetrize step gives).
#include <math.h>

#include "us_common.cuh"

namespace us {

struct RuleAxis {
  void* x;
  void* w;
  double alpha, beta, mu0, wdiv;  // wdiv: type c divides its weights by 2^(alpha+beta+2) (_integral.py:104)
  int32_t count;                  // points of this axis (2n for type a)
  uint8_t kind, dtype;
  uint8_t pad[2];
};
struct RuleParams {
  RuleAxis ax[US_MAX_AXES];
};
static_assert(sizeof(RuleParams) < 8192, "rule parameter block");

constexpr int kRuleWarps = 8;  // roots per block

struct Eval {
  double p, dp, ssq;
  int changes;
};

// coef[3k..3k+2] = a_k, rb_{k+1}, b_k * rb_{k+1}   (rb = 1/b; the last row holds a_{n-1}, 1, b_{n-1})
// so that one step is  p_{k+1} = ((x - a_k) rb_{k+1}) p_k - (b_k rb_{k+1}) p_{k-1}:  one FMA on the chain.
template <bool DERIV>
__device__ __forceinline__ Eval eval_rule(const double* __restrict__ coef, int n, double x) {
  double pm = 0.0, p = 1.0, dm = 0.0, d = 0.0, ssq = 1.0;
  int changes = 0;
#pragma unroll 4  // lets the shared-memory loads and (x - a_k) rb_{k+1} run ahead of the p chain
  for (int k = 0; k < n; ++k) {
    const double a = coef[3 * k], rb = coef[3 * k + 1], g = coef[3 * k + 2];
    const double t = (x - a) * rb;
    const double pn = fma(t, p, -(g * pm));
    if (DERIV) {
      const double dn = fma(t, d, fma(rb, p, -(g * dm)));
      dm = d;
      d = dn;
      if (k < n - 1) ssq = fma(pn, pn, ssq);
    }
    changes += (int)((uint32_t)(__double2hiint(pn) ^ __double2hiint(p)) >> 31);
    pm = p;
    p = pn;
  }
  return Eval{p, d, ssq, changes};
}

template <typename T>
__device__ __forceinline__ void put(void* base, int i, double v) {
  static_cast<T*>(base)[i] = (T)v;
}
__device__ __forceinline__ void store_point(const RuleAxis& ax, int i, double x, double w) {
  if (ax.dtype == US_F32) {
    put<float>(ax.x, i, x);
    put<float>(ax.w, i, w);
  } else {
    put<double>(ax.x, i, x);
    put<double>(ax.w, i, w);
  }
}

__global__ void __launch_bounds__(kRuleWarps * 32) quadrature_rules_kernel(const __grid_constant__ RuleParams p) {
  extern __shared__ double coef[];
  const RuleAxis& ax = p.ax[blockIdx.y];
  const int n = ax.count;
  if (ax.kind == US_NODE_A) {
    // periodic trapezoid on [0, 2 pi): x_i = i*pi/n', w = pi/n' with n' = count/2 (_integral.py:86-87),
    // rounded in the output type the way torch's CUDA kernels evaluate `arange(2n) * pi / n`: the
    // division by a host scalar is a multiplication by its rounded reciprocal (<= 1 ulp from i*pi/n')
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int half = n / 2;
    if (ax.dtype == US_F32) {
      const float pi = 3.14159265358979323846f, inv = __fdiv_rn(1.0f, (float)half);
      static_cast<float*>(ax.x)[i] = __fmul_rn(__fmul_rn((float)i, pi), inv);
      static_cast<float*>(ax.w)[i] = __fmul_rn(pi, inv);
    } else {
      const double pi = 3.14159265358979323846, inv = __ddiv_rn(1.0, (double)half);
      static_cast<double*>(ax.x)[i] = __dmul_rn(__dmul_rn((double)i, pi), inv);
      static_cast<double*>(ax.w)[i] = __dmul_rn(pi, inv);
    }
    return;
  }
  const bool symmetric = ax.alpha == ax.beta;
  const int work = symmetric ? (n + 1) / 2 : n;  // roots this axis solves for
  if ((int)blockIdx.x * kRuleWarps >= work) return;

  const double a = ax.alpha, b = ax.beta, s = a + b;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    // a_k and the off-diagonal entries b_k, b_{k+1} of the Jacobi matrix (b_0 := 0)
    auto offdiag = [&](int j) -> double {
      const double dj = (double)j;
      if (j == 1) return 2.0 / (2.0 + s) * sqrt((1.0 + a) * (1.0 + b) / (3.0 + s));
      return 2.0 / (2.0 * dj + s) * sqrt(dj * (dj + a) * (dj + b) * (dj + s) / ((2.0 * dj + s + 1.0) * (2.0 * dj + s - 1.0)));
    };
    const double dk = (double)k;
    const double diag = k == 0 ? (b - a) / (2.0 + s) : (b * b - a * a) / ((2.0 * dk + s) * (2.0 * dk + s + 2.0));
    const double bk = k == 0 ? 0.0 : offdiag(k);
    const double rb = k == n - 1 ? 1.0 : 1.0 / offdiag(k + 1);
    coef[3 * k] = diag;
    coef[3 * k + 1] = rb;
    coef[3 * k + 2] = bk * rb;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kRuleWarps + (threadIdx.x >> 5);  // t = 0 is the largest root
  if (t >= work) return;
  const int j = n - 1 - t;  // ascending index of this warp's root
  double x;
  if (symmetric && (n & 1) && j == n / 2) {
    x = 0.0;
  } else {
    // root j exceeds x  <=>  (roots above x) >= n - j
    double lo = symmetric ? 0.0 : -1.0, hi = 1.0;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {
      const double h = (hi - lo) * (1.0 / 33.0);
      const double xl = fma((double)(lane + 1), h, lo);
      const int above = eval_rule<false>(coef, n, xl).changes;
      const int m = __popc(__ballot_sync(0xffffffffu, above >= n - j));
      hi = fma((double)(m + 1), h, lo);
      lo = fma((double)m, h, lo);
    }
    const double slack = hi - lo;
    x = 0.5 * (lo + hi);
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const Eval e = eval_rule<true>(coef, n, x);
      x = fmin(fmax(x - e.p / e.dp, lo - slack), hi + slack);
    }
  }
  const Eval e = eval_rule<true>(coef, n, x);
  if (x != 0.0) x -= e.p / e.dp;
  if (lane != 0) return;
  double w = ax.mu0 / e.ssq;
  double theta, theta_mirror;
  if (ax.kind == US_NODE_B) {
    theta = acos(x);
    theta_mirror = acos(-x);
  } else if (ax.kind == US_NODE_BP) {
    theta = asin(x);
    theta_mirror = asin(-x);
  } else {
    w /= ax.wdiv;
    theta = acos(x) * 0.5;
    theta_mirror = acos(-x) * 0.5;
  }
  store_point(ax, j, theta, w);
  if (symmetric && j != n - 1 - j) store_point(ax, n - 1 - j, theta_mirror, w);
}

// B(p, q) for p, q positive multiples of 1/2 by the recurrence B(p+1, q) = B(p, q) p / (p+q) from
// B(1/2,1/2) = pi, B(1/2,1) = B(1,1/2) = 2, B(1,1) = 1 (a few roundings); lgamma otherwise.
static double beta_function(double p, double q) {
  const bool halves = p > 0 && q > 0 && 2 * p == floor(2 * p) && 2 * q == floor(2 * q) && p + q < 4000;
  if (!halves) return exp(lgamma(p) + lgamma(q) - lgamma(p + q));
  const double p0 = (fmod(2 * p, 2.0) == 1.0) ? 0.5 : 1.0, q0 = (fmod(2 * q, 2.0) == 1.0) ? 0.5 : 1.0;
  double v = (p0 == 0.5 && q0 == 0.5) ? 3.14159265358979323846 : ((p0 == 1.0 && q0 == 1.0) ? 1.0 : 2.0);
  for (double u = p0; u < p; u += 1.0) v *= u / (u + q0);
  for (double u = q0; u < q; u += 1.0) v *= u / (u + p);
  return v;
}

}  // namespace us

using namespace us;

extern "C" int us_quadrature_rules(int n_axes, const uint8_t* kind, const int64_t* count, const double* alpha,
                                   const double* beta, const int32_t* dtype, void* const* x_dev,
                                   void* const* w_dev, void* stream) {
  static const char* who = "us_quadrature_rules";
  if (n_axes < 0 || n_axes > US_MAX_AXES) return fail(US_ECAPACITY, "%s: %d axes not in 0..%d", who, n_axes, US_MAX_AXES);
  if (n_axes == 0) return 0;
  if (!kind || !count || !alpha || !beta || !dtype || !x_dev || !w_dev) return fail(US_EINVAL, "%s: null argument", who);
  RuleParams p{};
  int64_t blocks = 0, longest = 0;
  for (int i = 0; i < n_axes; ++i) {
    RuleAxis& ax = p.ax[i];
    if (kind[i] > US_NODE_C) return fail(US_EINVAL, "%s: axis %d has kind %d", who, i, (int)kind[i]);
    if (dtype[i] != US_F32 && dtype[i] != US_F64) return fail(US_EINVAL, "%s: axis %d has dtype %d", who, i, dtype[i]);
    if (count[i] < 1) return fail(US_EINVAL, "%s: axis %d has %lld points", who, i, (long long)count[i]);
    if (!x_dev[i] || !w_dev[i]) return fail(US_EINVAL, "%s: axis %d has a null output", who, i);
    ax.x = x_dev[i];
    ax.w = w_dev[i];
    ax.kind = kind[i];
    ax.dtype = (uint8_t)dtype[i];
    ax.count = (int32_t)count[i];
    if (kind[i] == US_NODE_A) {
      if (count[i] % 2 != 0 || count[i] > (int64_t)1 << 30)
        return fail(US_EINVAL, "%s: axis %d (type a) needs an even number of points <= 2^30, got %lld", who, i, (long long)count[i]);
      blocks = blocks > (count[i] + 255) / 256 ? blocks : (count[i] + 255) / 256;
      continue;
    }
    if (count[i] > US_MAX_RULE_POINTS)
      return fail(US_ECAPACITY, "%s: axis %d asks for %lld Gauss points > %d", who, i, (long long)count[i], US_MAX_RULE_POINTS);
    if (!(alpha[i] > -1.0) || !(beta[i] > -1.0)) return fail(US_EINVAL, "%s: axis %d needs alpha, beta > -1", who, i);
    ax.alpha = alpha[i];
    ax.beta = beta[i];
    ax.mu0 = exp2(alpha[i] + beta[i] + 1.0) * beta_function(alpha[i] + 1.0, beta[i] + 1.0);
    ax.wdiv = exp2(alpha[i] + beta[i] + 2.0);
    if (!isfinite(ax.mu0) || ax.mu0 <= 0) return fail(US_EINVAL, "%s: axis %d: weight mass not representable", who, i);
    const int64_t work = alpha[i] == beta[i] ? (count[i] + 1) / 2 : count[i];
    blocks = blocks > (work + kRuleWarps - 1) / kRuleWarps ? blocks : (work + kRuleWarps - 1) / kRuleWarps;
    longest = longest > count[i] ? longest : count[i];
  }
  const size_t smem = (size_t)longest * 3 * sizeof(double);  // <= 48 KB by US_MAX_RULE_POINTS
  quadrature_rules_kernel<<<dim3((unsigned)blocks, (unsigned)n_axes), kRuleWarps * 32, smem, (cudaStream_t)stream>>>(p);
  US_CUDA_TRY(cudaGetLastError(), "quadrature_rules launch");
  return 0;
}
