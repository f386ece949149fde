// ivx_bsiv.cu — kernel C: fused fp64 Black-Scholes implied-volatility solver, one thread per quote.
//
// The reference has no solver; conventions follow reference
// src/interpolation/greeks.py:21-22 (continuous rate r, no dividend, T in years):
//     d1 = (ln(S/K) + (r + sigma^2/2) T) / (sigma sqrt T),  d2 = d1 - sigma sqrt T.
//
// Method.  With Kd = K e^{-rT}, x = -|ln(S/Kd)| and s = sigma sqrt(T) the out-of-the-money price,
// normalised by sqrt(S Kd), is
//     b(x,s) = e^{x/2} Phi(x/s + s/2) - e^{-x/2} Phi(x/s - s/2)      (in-the-money quotes are first
// moved across put-call parity).  b is increasing in s with its inflection at s_c = sqrt(2|x|):
//   * beta <  b(s_c): root in the convex branch — solve ln b(s) = ln beta, evaluated through
//     erfcx so that deep-wing prices never underflow: b = E/2 [erfcx(q1) - erfcx(q2)],
//     E = exp(-(h^2+t^2)/2), h = x/s, t = s/2, q1,2 = (|h| -+ t)/sqrt2, b' = E/sqrt(2 pi);
//   * beta >= b(s_c): root in the concave branch — solve b(s) = beta directly.
// Both branches run bracketed Halley iterations (third order) and fall back to bisection of the
// bracket when a step leaves it ("Newton/Brent" safeguarding); a warp leaves the loop when every
// lane has converged (warp vote).  Loads are structure-of-arrays, one coalesced stream per column.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr double SQRT2 = 1.4142135623730951;
constexpr double INV_SQRT2 = 0.7071067811865476;
constexpr double SQRT_2PI = 2.5066282746310002;
constexpr double INV_SQRT_2PI = 0.3989422804014327;
constexpr double S_MAX = 40.0;   // sigma*sqrt(T) bracket; fp64 prices cannot encode more than ~17
constexpr int MAX_ITERS = 64;

__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * INV_SQRT2); }
__device__ __forceinline__ double norm_pdf(double z) { return INV_SQRT_2PI * exp(-0.5 * z * z); }

// normalised OTM price in the concave branch (s >= s_c, so h + t >= 0)
__device__ __forceinline__ void eval_upper(double x, double ex2, double iex2, double s, double& b, double& db,
                                           double& d2ratio) {
  const double h = x / s, t = 0.5 * s;
  b = ex2 * norm_cdf(h + t) - iex2 * norm_cdf(h - t);
  db = INV_SQRT_2PI * exp(-0.5 * (h * h + t * t));
  d2ratio = h * h / s - 0.25 * s;    // b''/b'
}

// ln b, (ln b)', via scaled complementary error functions (s <= ~s_c, so q1 >= 0 up to overshoot)
__device__ __forceinline__ void eval_lower(double x, double s, double& lnb, double& dlnb, double& d2ratio_b) {
  const double h = -x / s /* = |h| */, t = 0.5 * s;
  const double D = erfcx((h - t) * INV_SQRT2) - erfcx((h + t) * INV_SQRT2);
  lnb = -0.5 * (h * h + t * t) + log(0.5 * D);
  dlnb = 2.0 * INV_SQRT_2PI / D;      // b'/b = (E/sqrt(2pi)) / (E D / 2)
  d2ratio_b = h * h / s - 0.25 * s;   // b''/b'
}

__device__ __forceinline__ double solve_total_vol(double x, double beta, int& iters) {
  // x <= 0, 0 < beta < e^{x/2}
  const double ex2 = exp(0.5 * x), iex2 = 1.0 / ex2;
  const double s_c = sqrt(-2.0 * x);
  const double b_c = 0.5 * ex2 - iex2 * 0.5 * erfc(s_c * INV_SQRT2);
  const bool lower = beta < b_c;
  double lo, hi, s;
  double lnbeta = 0.0;
  if (lower) {
    lnbeta = log(beta);
    lo = 0.0; hi = s_c;
    // match ln b ~ C - x^2/(2 s^2) at s_c  ->  s0 = |x| / sqrt(2 (ln b_c - ln beta) + |x|/2)
    s = -x / sqrt(2.0 * (log(b_c) - lnbeta) - 0.5 * x);
  } else {
    lo = s_c; hi = S_MAX;
    s = s_c + (beta - b_c) * SQRT_2PI * iex2;   // Newton step from the inflection point (stays left of the root)
  }
  if (!(s > lo && s < hi)) s = 0.5 * (lo + hi);
  bool done = false;
  int it = 0;
  while (true) {
    if (!done) {
      double g, dg, ratio;   // objective, its derivative, g''/g'
      if (lower) {
        double lnb, dlnb, r2;
        eval_lower(x, s, lnb, dlnb, r2);
        g = lnb - lnbeta;
        dg = dlnb;
        ratio = r2 - dlnb;                 // (ln b)''/(ln b)' = b''/b' - b'/b
      } else {
        double b, db, r2;
        eval_upper(x, ex2, iex2, s, b, db, r2);
        g = b - beta;
        dg = db;
        ratio = r2;
      }
      ++it;
      if (g > 0.0) hi = s; else lo = s;
      const double newton = -g / dg;
      const double den = 1.0 + 0.5 * newton * ratio;            // Halley: dn / (1 + dn g''/(2 g'))
      double step = (den > 0.5 && den < 2.0) ? newton / den : newton;
      double sn = s + step;
      bool safe = (sn > lo) && (sn < hi);
      if (!safe) { sn = 0.5 * (lo + hi); step = sn - s; }
      // cubic convergence: a step below 1e-6 s leaves an error far below one ulp
      if ((safe && fabs(step) <= 1e-6 * sn) || !(hi - lo > 4.0e-16 * hi) || it >= MAX_ITERS) done = true;
      s = sn;
    }
    if (__all_sync(__activemask(), done)) break;
  }
  iters = it;
  return s;
}

__global__ void __launch_bounds__(256) k_bs_implied_vol(int64_t n, const double* __restrict__ price,
                                                        const double* __restrict__ S, const double* __restrict__ K,
                                                        const double* __restrict__ T, const double* __restrict__ r,
                                                        const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                        int32_t* __restrict__ iters) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // every lane takes part in the convergence vote; lanes past the end carry an already-solved dummy
  const bool inr = i < n;
  double p = 0, s_ = 1, k_ = 1, t_ = 1, r_ = 0;
  bool call = true;
  if (inr) {
    p = ld_stream(price + i); s_ = ld_stream(S + i); k_ = ld_stream(K + i);
    t_ = ld_stream(T + i); r_ = ld_stream(r + i); call = is_call[i] != 0;
  }
  const double kd = k_ * exp(-r_ * t_);
  const double intrinsic = call ? (s_ - kd) : (kd - s_);
  const double lower = fmax(intrinsic, 0.0);
  const double upper = call ? s_ : kd;
  bool ok = inr && isfinite(p) && isfinite(s_) && isfinite(k_) && isfinite(t_) && isfinite(r_) &&
            (s_ > 0.0) && (k_ > 0.0) && (t_ > 0.0) && (p > lower) && (p < upper);
  const double target = intrinsic > 0.0 ? p - intrinsic : p;    // the out-of-the-money twin
  ok = ok && (target > 0.0);
  double out = qnan();
  int it = 0;
  // warp-uniform entry keeps the vote inside solve_total_vol well defined
  const unsigned live = __ballot_sync(IVX_FULL_MASK, ok);
  if (live) {
    double x = ok ? -fabs(log(s_ / kd)) : -0.1;
    double beta = ok ? target / sqrt(s_ * kd) : 0.05;
    const double bmax = exp(0.5 * x);
    if (ok && !(beta < bmax)) { ok = false; x = -0.1; beta = 0.05; }   // rounding pushed it onto the bound
    int its;
    double s = solve_total_vol(x, beta, its);
    if (ok) { out = s / sqrt(t_); it = its; }
  }
  if (inr) {
    st_stream(sigma + i, out);
    if (iters) st_stream(iters + i, it);
  }
}

__global__ void __launch_bounds__(256) k_bs_price(int64_t n, const double* __restrict__ S, const double* __restrict__ K,
                                                  const double* __restrict__ T, const double* __restrict__ r,
                                                  const double* __restrict__ sigma, const uint8_t* __restrict__ is_call,
                                                  double* __restrict__ price) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i], v = sigma[i];
  const double sq = sqrt(t_);
  const double d1 = (log(s_ / k_) + (r_ + 0.5 * v * v) * t_) / (v * sq);   // greeks.py:21
  const double d2 = d1 - v * sq;                                           // greeks.py:22
  const double kd = k_ * exp(-r_ * t_);
  price[i] = is_call[i] ? s_ * norm_cdf(d1) - kd * norm_cdf(d2) : kd * norm_cdf(-d2) - s_ * norm_cdf(-d1);
}

__global__ void __launch_bounds__(256) k_bs_greeks(int64_t n, const double* __restrict__ S, const double* __restrict__ K,
                                                   const double* __restrict__ T, const double* __restrict__ r,
                                                   const double* __restrict__ sigma, const uint8_t* __restrict__ is_call,
                                                   int32_t ref_put_rho, double* delta, double* gamma, double* theta,
                                                   double* vega, double* rho) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i], v = sigma[i];
  const bool call = is_call[i] != 0;
  const double sq = sqrt(t_);
  const double d1 = (log(s_ / k_) + (r_ + 0.5 * v * v) * t_) / (v * sq);   // greeks.py:21
  const double d2 = d1 - v * sq;                                           // greeks.py:22
  const double pdf = norm_pdf(d1);
  const double disc = exp(-r_ * t_);
  if (delta) delta[i] = call ? norm_cdf(d1) : norm_cdf(d1) - 1.0;          // :25,:29
  if (theta)
    theta[i] = call ? (-s_ * pdf * v / (2.0 * sq) - r_ * k_ * disc * norm_cdf(d2)) / 365.0    // :26-27
                    : (-s_ * pdf * v / (2.0 * sq) + r_ * k_ * disc * norm_cdf(-d2)) / 365.0;  // :30-31
  if (gamma) gamma[i] = pdf / (s_ * v * sq);                               // :33
  if (vega) vega[i] = s_ * pdf * sq / 100.0;                               // :34
  if (rho) {
    double rr = k_ * t_ * disc * norm_cdf(call ? d2 : -d2) / 100.0;        // :35 (put sign as shipped)
    if (!call && !ref_put_rho) rr = -rr;
    rho[i] = rr;
  }
}

// ---- roofline probes ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probe_copy(int64_t n2, const double2* __restrict__ src, double2* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_probe_fp64(int32_t iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int k = 0; k < iters; ++k) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double t = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (t == 123.456) sink[0] = t;   // never true; keeps the chain alive
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int ivx_bs_implied_vol(int64_t n, const double* price, const double* S, const double* K, const double* T,
                                  const double* r, const uint8_t* is_call, double* sigma, int32_t* iters,
                                  ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(price && S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol: null pointer");
  k_bs_implied_vol<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, price, S, K, T, r, is_call, sigma, iters);
  IVX_LAUNCH_CHECK("k_bs_implied_vol");
  return IVX_OK;
}

extern "C" int ivx_bs_price(int64_t n, const double* S, const double* K, const double* T, const double* r,
                            const double* sigma, const uint8_t* is_call, double* price, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_price: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(price && S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_price: null pointer");
  k_bs_price<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, S, K, T, r, sigma, is_call, price);
  IVX_LAUNCH_CHECK("k_bs_price");
  return IVX_OK;
}

extern "C" int ivx_bs_greeks(int64_t n, const double* S, const double* K, const double* T, const double* r,
                             const double* sigma, const uint8_t* is_call, int32_t reference_put_rho_sign, double* delta,
                             double* gamma, double* theta, double* vega, double* rho, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_greeks: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_greeks: null pointer");
  k_bs_greeks<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, S, K, T, r, sigma, is_call,
                                                                           reference_put_rho_sign, delta, gamma, theta, vega, rho);
  IVX_LAUNCH_CHECK("k_bs_greeks");
  return IVX_OK;
}

extern "C" int ivx_probe_copy(int64_t n_f64, const double* src, double* dst, ivx_stream_t stream) {
  IVX_REQUIRE(n_f64 >= 0 && (n_f64 % 2) == 0 && src && dst, IVX_ERR_INVALID_ARG, "ivx_probe_copy: need an even count and non-null buffers");
  IVX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0, IVX_ERR_INVALID_ARG, "ivx_probe_copy: buffers must be 16-byte aligned");
  k_probe_copy<<<sm_count() * 16, 256, 0, (cudaStream_t)stream>>>(n_f64 / 2, (const double2*)src, (double2*)dst);
  IVX_LAUNCH_CHECK("k_probe_copy");
  return IVX_OK;
}

extern "C" int ivx_probe_fp64_fma(int64_t n_threads, int32_t iters, double* sink, ivx_stream_t stream) {
  IVX_REQUIRE(n_threads > 0 && iters > 0 && sink, IVX_ERR_INVALID_ARG, "ivx_probe_fp64_fma: bad arguments");
  k_probe_fp64<<<(unsigned)((n_threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(iters, sink);
  IVX_LAUNCH_CHECK("k_probe_fp64");
  return IVX_OK;
}
