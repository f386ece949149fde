// ivx_spline_band.cu — kernel D on shared ladders as a BANDED linear operator.
//
// For a fixed ladder of abscissae x[0..n) and a fixed evaluation grid xq[0..n_q), "fit the natural cubic spline through
// (x, y) and evaluate it on xq" is a fixed LINEAR map  yq = W y.  Row q of W is
//     ta e_{i} + tb e_{i+1} + C row_i(A^{-1} R) + D row_{i+1}(A^{-1} R),        i = bracket of xq[q],
// (A the tridiagonal matrix of the second-derivative system, R the second-difference operator), and the rows of A^{-1}
// decay geometrically away from the diagonal — by 2 - sqrt 3 = 0.268 per knot on a uniform ladder — so BAND_L = 22 rows
// either side of the bracket carry everything above 1e-12 and a slice is n_q short dot products of BAND_W = 48 terms:
// ~25 k fused multiply-adds instead of the ~500 k instructions of the scan-based solve, and the abscissae are not needed
// (8 B/knot instead of 16; they are still read where a slice has to be VERIFIED to sit on its ladder).
//   * k_band_build   one thread per (ladder, grid point): the two local Thomas solves (46 rows), the 48 weights, the
//                    window start.  Truncation is never trusted: k_band_check fits two probe vectors exactly (full Thomas)
//                    and compares; a ladder whose band misses 2e-12 is marked not ok and its slices take the exact solve.
//   * k_band_apply   one CTA per (ladder, batch of slices), one thread per grid point holding its 48 weights in
//                    REGISTERS for the whole batch; a slice's y (and x, when verified) is brought into shared memory by a
//                    3-deep cp.async pipeline, ONE block barrier per slice (it also votes "every knot finite and on the
//                    ladder"); grid points are dealt to lanes 32 apart so that the 16 lanes of a half-warp read 16
//                    different banks; results leave through a shared-memory transpose, coalesced.
//                    Slices that cannot be taken (a missing knot, off the ladder, odd length) are FLAGGED (one int32 per slice).
//   * the flagged slices are then solved by the exact kernel (ivx_spline.cu, k_spline_fit_eval_list), in slice order, same stream.
//   * k_band_apply_mma  the same operator as a BLOCK-BANDED MATRIX PRODUCT on the fp64 tensor path (mma.sync m8n8k4, SASS DMMA:
//                    measured 37.2 TFLOP/s on B200, the DFMA rate, but 256 multiply-adds per instruction instead of 32).  With
//                    all slices of a ladder sharing W, yq[slice][point] = sum_k y[slice][k] W[point][k] IS a contraction: a warp
//                    owns 8 consecutive grid points, whose windows jointly span ~76 knots, keeps their weights as B fragments in
//                    registers for a whole batch of slices, and multiplies 8 slices at a time: 19 DMMA + 19 shared-memory loads
//                    per 8 x 8 outputs, where the dot-product form needs 96 loads and 96 x 32 DFMA lanes (its limiter, per ncu,
//                    was the shared-memory instruction queue).  A CTA = 64 grid points (8 warps); it stages only the ~300 knots
//                    its points touch, 8 slices per stage, two stages of cp.async, ONE block barrier per 8 slices.
//                    Ladders whose windows are too wide for the fragment registers keep k_band_apply.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr int BAND_L = 22;
constexpr int BAND_W = 2 * BAND_L + 4;      // 48
constexpr int BAND_MAXQ = 512;
constexpr int BAND_MAXK = 2048;
constexpr int BAND_NBUF = 3;
constexpr int BAND_THREADS = 512;
constexpr int BAND_BATCH = 48;              // slices per CTA at most: the 200 KB of weights are read once per batch
// tensor-path form
constexpr int MMA_KS = 24;                  // k-steps of 4 knots per 8-point group at most (96 knots): B fragments held in 48 registers
constexpr int MMA_OCT = BAND_MAXQ / 8;      // 64 groups of 8 grid points
constexpr int MMA_SLAB_MAX = 512;           // knots a 64-point CTA may touch
constexpr int MMA_BATCH = 512;              // slices per CTA (64 tiles of 8): the 48 KB of fragments are read once per batch
constexpr int LADDER_OK_BAND = 1, LADDER_OK_MMA = 2, LADDER_OK_FILL = 4;   // bits of ladder_ok
constexpr int FILL_MARGIN = 24;             // knots closer than this to an end of the ladder are not filled (the slice takes the exact solve)
constexpr int FILL_SAMPLES = 32;            // fill rows per ladder checked against the full solve

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- build ---------------------------------------------------------------------------------------------------------
// The 48 weights (and their first knot k0) of "fit the natural spline through (X(i), y[i]), i < n, and evaluate it at xv", for
// xv inside [X(0), X(n-1)] or extrapolated along the end cubic: the two local Thomas solves around the bracket of xv.
// X is an accessor, so that the same code serves a ladder (k_band_build) and a ladder with one knot left out (k_band_fill_build).
template <class XA>
__device__ __forceinline__ void band_row(const XA& X, int n, double xv, double* __restrict__ w, int& k0_out) {
  // bracket: largest i with X(i) <= xv, clamped to a segment
  int lo_ = 0, hi_ = n;
  while (lo_ < hi_) { const int mid = (lo_ + hi_) >> 1; if (X(mid) <= xv) lo_ = mid + 1; else hi_ = mid; }
  int i0 = lo_ - 1;
  i0 = i0 < 0 ? 0 : (i0 > n - 2 ? n - 2 : i0);
  const int lo = max(1, i0 - BAND_L), hi = min(n - 2, i0 + 1 + BAND_L);   // interior equations kept
  const int m = hi - lo + 1;                            // <= 2 BAND_L + 2; n >= 48 -> m >= 1
  double cp[2 * BAND_L + 2], z1[2 * BAND_L + 2], z2[2 * BAND_L + 2];
  {
    double cprev = 0.0, d1 = 0.0, d2 = 0.0;
    for (int j = 0; j < m; ++j) {
      const int i = lo + j;
      const double hm = X(i) - X(i - 1), hp = X(i + 1) - X(i);
      const double inv = 1.0 / (2.0 * (hm + hp) - hm * cprev);
      cprev = hp * inv;
      d1 = ((i == i0 ? 1.0 : 0.0) - hm * d1) * inv;
      d2 = ((i == i0 + 1 ? 1.0 : 0.0) - hm * d2) * inv;
      cp[j] = cprev; z1[j] = d1; z2[j] = d2;
    }
    double m1 = 0.0, m2 = 0.0;
    for (int j = m - 1; j >= 0; --j) {
      m1 = z1[j] - cp[j] * m1; m2 = z2[j] - cp[j] * m2;
      z1[j] = m1; z2[j] = m2;
    }
  }
  const double hseg = X(i0 + 1) - X(i0);
  const double ta = (X(i0 + 1) - xv) / hseg, tb = (xv - X(i0)) / hseg;
  const double C = (ta * ta * ta - ta) * hseg * hseg * (1.0 / 6.0), D = (tb * tb * tb - tb) * hseg * hseg * (1.0 / 6.0);
  int k0 = lo - 1;
  k0 = k0 < 0 ? 0 : (k0 > n - BAND_W ? n - BAND_W : k0);
  k0_out = k0;
  auto zc = [&](int i) -> double { return (i >= lo && i <= hi) ? fma(C, z1[i - lo], D * z2[i - lo]) : 0.0; };   // C z_i0 + D z_{i0+1} at row i
  for (int k = lo - 1; k <= hi + 1; ++k) {
    // d/dy[k] of sum_i zc(i) r[i],  r[i] = 6 ((y[i+1]-y[i])/h[i] - (y[i]-y[i-1])/h[i-1]),  h[i] = X(i+1)-X(i)
    double v = 0.0;
    if (k >= 1) v += zc(k - 1) / (X(k) - X(k - 1));                                       // row k-1: + y[k]/h[k-1]
    if (k >= 1 && k <= n - 2) v -= zc(k) * (1.0 / (X(k) - X(k - 1)) + 1.0 / (X(k + 1) - X(k)));
    if (k <= n - 2) v += zc(k + 1) / (X(k + 1) - X(k));                                   // row k+1: + y[k]/h[k]
    v *= 6.0;
    if (k == i0) v += ta;
    if (k == i0 + 1) v += tb;
    w[k - k0] = v;
  }
}

__global__ void __launch_bounds__(128) k_band_build(int32_t n_ladders, const int64_t* __restrict__ ladder_off, const double* __restrict__ x,
                                                    int32_t n_q, const double* __restrict__ xq, int32_t extrap,
                                                    int32_t* __restrict__ band_k0, double* __restrict__ band_w, uint8_t* __restrict__ ladder_ok) {
  const int l = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  const int64_t a0 = ladder_off[l];
  const int64_t n64 = ladder_off[l + 1] - a0;
  const double* __restrict__ X = x + a0;
  double* __restrict__ w = band_w + ((int64_t)l * n_q + q) * BAND_W;
  int32_t* k0p = band_k0 + (int64_t)l * n_q + q;
  for (int j = 0; j < BAND_W; ++j) w[j] = 0.0;
  *k0p = 0;
  bool usable = n64 >= BAND_W && n64 <= BAND_MAXK && n_q <= BAND_MAXQ;
  if (usable && q == 0) {                   // strictly increasing, finite abscissae
    for (int i = 0; i < (int)n64; ++i) {
      const double xi = X[i];
      if (!(fabs(xi) < INFINITY) || (i > 0 && !(xi > X[i - 1]))) usable = false;
    }
  }
  if (q == 0) ladder_ok[l] = usable ? 1 : 0;           // k_band_check may still clear it
  if (!(n64 >= BAND_W && n64 <= BAND_MAXK)) return;
  const int n = (int)n64;
  const double xv = xq[q];
  if (isnan(xv)) { *k0p = -1; return; }
  const bool below = xv < X[0], above = xv > X[n - 1];
  if ((below || above) && extrap == 0) { *k0p = -1; return; }
  if ((below || above) && extrap == 2) {               // flat: the end knot's value
    const int kk = below ? 0 : n - 1;
    const int k0 = below ? 0 : n - BAND_W;
    *k0p = k0; w[kk - k0] = 1.0;
    return;
  }
  int k0;
  band_row([&](int i) { return X[i]; }, n, xv, w, k0);
  *k0p = k0;
}

// ---- exactness check: two probe vectors per ladder, banded against the full Thomas solve -------------------------------
__device__ __forceinline__ double probe_value(int k, int which) {
  uint32_t h = (uint32_t)k * 2654435761u + 0x9E3779B9u * (uint32_t)(which + 1);
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
  const double noise = (double)h * (1.0 / 4294967296.0) - 0.5;
  return which == 0 ? sin(0.37 * k) + 0.25 * noise : 2.0 * noise;
}

__global__ void __launch_bounds__(256) k_band_check(const int64_t* __restrict__ ladder_off, const double* __restrict__ x, int32_t n_q,
                                                    const double* __restrict__ xq, int32_t extrap, const int32_t* __restrict__ band_k0,
                                                    const double* __restrict__ band_w, uint8_t* __restrict__ ladder_ok) {
  __shared__ double sM[BAND_MAXK], scp[BAND_MAXK];      // the probe y is a function of (k, which): recomputed where needed
  __shared__ int bad;
  const int l = blockIdx.x;
  if (!ladder_ok[l]) return;
  const int64_t a0 = ladder_off[l];
  const int n = (int)(ladder_off[l + 1] - a0);
  const double* __restrict__ X = x + a0;
  if (threadIdx.x == 0) bad = 0;
  for (int which = 0; which < 2; ++which) {
    auto sy = [&](int k) { return probe_value(k, which); };
    __syncthreads();
    // the probe vector is staged by the whole CTA (a sine and a hash per value: inside the serial elimination they were 90 % of
    // a 2.8 ms kernel that a streaming pipeline runs once per chunk); the elimination overwrites it in place behind itself
    for (int k = threadIdx.x; k < n; k += blockDim.x) sM[k] = sy(k);
    __syncthreads();
    if (threadIdx.x == 0) {                       // the reference: plain Thomas over the whole ladder
      double cprev = 0.0, dprev = 0.0;
      double y_prev = sM[0], y_cur = sM[1];
      sM[0] = 0.0;
      for (int i = 1; i <= n - 2; ++i) {
        const double y_next = sM[i + 1];
        const double hm = X[i] - X[i - 1], hp = X[i + 1] - X[i];
        const double inv = 1.0 / (2.0 * (hm + hp) - hm * cprev);
        cprev = hp * inv;
        dprev = (6.0 * ((y_next - y_cur) / hp - (y_cur - y_prev) / hm) - hm * dprev) * inv;
        scp[i] = cprev; sM[i] = dprev;
        y_prev = y_cur; y_cur = y_next;
      }
      sM[n - 1] = 0.0;
      double M = 0.0;
      for (int i = n - 2; i >= 1; --i) { M = sM[i] - scp[i] * M; sM[i] = M; }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
      const int k0 = band_k0[(int64_t)l * n_q + q];
      const double xv = xq[q];
      double exact;
      if (isnan(xv) || ((xv < X[0] || xv > X[n - 1]) && extrap == 0)) exact = qnan();
      else if (xv < X[0] && extrap == 2) exact = sy(0);
      else if (xv > X[n - 1] && extrap == 2) exact = sy(n - 1);
      else {
        int lo_ = 0, hi_ = n;
        while (lo_ < hi_) { const int mid = (lo_ + hi_) >> 1; if (X[mid] <= xv) lo_ = mid + 1; else hi_ = mid; }
        int i0 = lo_ - 1;
        i0 = i0 < 0 ? 0 : (i0 > n - 2 ? n - 2 : i0);
        const double h = X[i0 + 1] - X[i0], A = X[i0 + 1] - xv, B = xv - X[i0];
        exact = (sM[i0] * A * A * A + sM[i0 + 1] * B * B * B) / (6.0 * h) + (sy(i0) / h - sM[i0] * h / 6.0) * A + (sy(i0 + 1) / h - sM[i0 + 1] * h / 6.0) * B;
      }
      double got;
      if (k0 < 0) got = qnan();
      else {
        const double* __restrict__ w = band_w + ((int64_t)l * n_q + q) * BAND_W;
        got = 0.0;
        for (int j = 0; j < BAND_W; ++j) got = fma(w[j], sy(k0 + j), got);
      }
      const bool same = (isnan(exact) && isnan(got)) || fabs(got - exact) <= 2e-12 * (1.0 + fabs(exact));
      if (!same) bad = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && bad) ladder_ok[l] = 0;
}

// ---- missing knots ------------------------------------------------------------------------------------------------------
// A slice with a missing quote has the natural spline through its OTHER knots.  That spline is also a natural spline on the full
// ladder (its third derivative simply does not jump at the missing knot), so it equals W y~ with y~ = y at the quoted knots and
// y~[m] = the value at x[m] of the spline through the others — and THAT is one more row of the same kind: the weights of
// "fit through the ladder without knot m, evaluate at x[m]" (band_row on an accessor that skips m).  A missing knot is thus
// FILLED by a 48-term dot product and the slice stays on the banded path.  Exact up to the band's own 1e-12 as long as no other
// missing knot lies inside the 48-knot window (then the slice is flagged for the exact solve, as before); knots within
// FILL_MARGIN of an end are not filled either (a missing end knot changes where the spline is extrapolated).
// fill_k0[a0 + m] = first knot of row m in the indices of the ladder WITHOUT m (-1: no row), fill_w[(a0 + m) * 48 ..] its weights.
__global__ void __launch_bounds__(128) k_band_fill_build(const int64_t* __restrict__ ladder_off, const double* __restrict__ x,
                                                         uint8_t* __restrict__ ladder_ok, int32_t* __restrict__ fill_k0, double* __restrict__ fill_w) {
  const int l = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t a0 = ladder_off[l];
  const int n = (int)(ladder_off[l + 1] - a0);
  if (m >= n) return;
  const double* __restrict__ X = x + a0;
  double* __restrict__ w = fill_w + (a0 + m) * BAND_W;
  for (int j = 0; j < BAND_W; ++j) w[j] = 0.0;
  fill_k0[a0 + m] = -1;
  const bool usable = (ladder_ok[l] & LADDER_OK_BAND) && n - 1 >= BAND_W;
  if (m == 0 && usable) ladder_ok[l] |= LADDER_OK_FILL;          // k_band_fill_check may still clear it
  if (!usable || m < FILL_MARGIN || m > n - 1 - FILL_MARGIN) return;
  int k0;
  band_row([&](int i) { return X[i + (i >= m ? 1 : 0)]; }, n - 1, X[m], w, k0);
  fill_k0[a0 + m] = k0;
}

// FILL_SAMPLES rows per ladder against the full Thomas solve of the ladder without that knot, two probe vectors each
__global__ void __launch_bounds__(64) k_band_fill_check(const int64_t* __restrict__ ladder_off, const double* __restrict__ x,
                                                        const int32_t* __restrict__ fill_k0, const double* __restrict__ fill_w,
                                                        uint8_t* __restrict__ ladder_ok) {
  __shared__ double sM[BAND_MAXK], scp[BAND_MAXK];
  const int l = blockIdx.y;
  if (!(ladder_ok[l] & LADDER_OK_FILL)) return;
  const int64_t a0 = ladder_off[l];
  const int n = (int)(ladder_off[l + 1] - a0), nr = n - 1;
  const int span = n - 1 - 2 * FILL_MARGIN;
  if (span < 0) return;
  const int m = FILL_MARGIN + (int)(((int64_t)blockIdx.x * span) / (FILL_SAMPLES - 1));
  const double* __restrict__ X = x + a0;
  auto XR = [&](int i) { return X[i + (i >= m ? 1 : 0)]; };
  const int k0 = fill_k0[a0 + m];
  bool bad = k0 != m - 24;                                        // (the evaluation kernel relies on where the row starts)
  for (int which = 0; which < 2; ++which) {
    auto sy = [&](int k) { return probe_value(k, which); };       // values at the knots of the REDUCED ladder
    __syncthreads();
    for (int k = threadIdx.x; k < nr; k += blockDim.x) sM[k] = sy(k);   // staged by the CTA, overwritten in place by the elimination
    __syncthreads();
    if (threadIdx.x != 0 || bad) continue;
    double cprev = 0.0, dprev = 0.0;
    double y_prev = sM[0], y_cur = sM[1];
    sM[0] = 0.0;
    for (int i = 1; i <= nr - 2; ++i) {
      const double y_next = sM[i + 1];
      const double hm = XR(i) - XR(i - 1), hp = XR(i + 1) - XR(i);
      const double inv = 1.0 / (2.0 * (hm + hp) - hm * cprev);
      cprev = hp * inv;
      dprev = (6.0 * ((y_next - y_cur) / hp - (y_cur - y_prev) / hm) - hm * dprev) * inv;
      scp[i] = cprev; sM[i] = dprev;
      y_prev = y_cur; y_cur = y_next;
    }
    sM[nr - 1] = 0.0;
    double M = 0.0;
    for (int i = nr - 2; i >= 1; --i) { M = sM[i] - scp[i] * M; sM[i] = M; }
    const int i0 = m - 1;                                          // x[m] lies between the reduced knots m-1 and m
    const double xv = X[m], h = XR(i0 + 1) - XR(i0), A = XR(i0 + 1) - xv, B = xv - XR(i0);
    const double exact = (sM[i0] * A * A * A + sM[i0 + 1] * B * B * B) / (6.0 * h) + (sy(i0) / h - sM[i0] * h / 6.0) * A + (sy(i0 + 1) / h - sM[i0 + 1] * h / 6.0) * B;
    double got = 0.0;
    for (int j = 0; j < BAND_W; ++j) got = fma(fill_w[(a0 + m) * BAND_W + j], sy(k0 + j), got);
    if (!(fabs(got - exact) <= 2e-12 * (1.0 + fabs(exact)))) bad = true;
  }
  if (threadIdx.x == 0 && bad) ladder_ok[l] &= (uint8_t)~LADDER_OK_FILL;   // (several samples may clear the same bit: same value)
}

// ---- which grid point each thread of k_band_apply takes ------------------------------------------------------------------
// Thread t reads y[k0(q_t) + j], j = 0..47, one 64-bit shared-memory load per j; the 16 lanes of a half-warp are served
// together and collide when two window starts agree modulo 16.  Grid points are therefore dealt to half-warps so that each
// holds 16 different residues k0 mod 16 (possible whenever the residues are evenly spread, which a regular grid over a
// regular ladder gives; leftovers go wherever there is room).  perm[l * 512 + t] = grid point of thread t, -1 = none.
__global__ void __launch_bounds__(32) k_band_perm(int32_t n_q, const int32_t* __restrict__ band_k0, int16_t* __restrict__ perm) {
  __shared__ int16_t bucket[16][BAND_MAXQ];
  __shared__ int cnt[16];
  const int l = blockIdx.x;
  if (threadIdx.x != 0) return;
  for (int r = 0; r < 16; ++r) cnt[r] = 0;
  for (int q = 0; q < n_q; ++q) {
    const int k0 = band_k0[(int64_t)l * n_q + q];
    const int r = (k0 < 0 ? 0 : k0) & 15;
    bucket[r][cnt[r]++] = (int16_t)q;
  }
  int16_t* __restrict__ out = perm + (int64_t)l * BAND_THREADS;
  for (int t = 0; t < BAND_THREADS; ++t) out[t] = -1;
  int taken[16];
  for (int r = 0; r < 16; ++r) taken[r] = 0;
  int spare = 0;                                   // half-warps are filled front to back; slots nobody claimed are filled afterwards
  for (int hw = 0; hw < BAND_THREADS / 16; ++hw)
    for (int r = 0; r < 16; ++r)
      if (taken[r] < cnt[r]) out[hw * 16 + r] = bucket[r][taken[r]++];
  for (int r = 0; r < 16; ++r)
    while (taken[r] < cnt[r]) {
      while (spare < BAND_THREADS && out[spare] >= 0) ++spare;
      if (spare >= BAND_THREADS) break;
      out[spare] = bucket[r][taken[r]++];
    }
}

// ---- tensor-path tables: per (ladder, group of 8 points) the first knot kb (even) and the number of 4-knot steps that cover
// the union of the 8 windows, and the weights laid out as mma B fragments: frag[((l*64 + o)*MMA_KS + ks)*32 + lane] =
// W[point 8o + (lane>>2)][knot kb + 4 ks + (lane&3)].  ladder_ok gets LADDER_OK_MMA when every group fits MMA_KS steps and
// every 64-point slab MMA_SLAB_MAX knots.  meta[l] = {largest slab extent, largest step count}.
__global__ void __launch_bounds__(256) k_band_frag(int32_t n_q, const int64_t* __restrict__ ladder_off, const int32_t* __restrict__ band_k0,
                                                   const double* __restrict__ band_w, uint8_t* __restrict__ ladder_ok,
                                                   double* __restrict__ frag, int32_t* __restrict__ oct, int32_t* __restrict__ meta) {
  __shared__ int s_kb[MMA_OCT], s_ke[MMA_OCT];
  __shared__ int s_bad, s_ext, s_nks;
  const int l = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = (int)(ladder_off[l + 1] - ladder_off[l]);
  if (threadIdx.x == 0) { s_bad = 0; s_ext = 0; s_nks = 0; }
  __syncthreads();
  const bool band_ok = (ladder_ok[l] & LADDER_OK_BAND) != 0;
  for (int o = warp; o < MMA_OCT; o += 8) {
    const int g = lane >> 2, tig = lane & 3;
    const int q = 8 * o + g;
    const int k0 = (band_ok && q < n_q) ? band_k0[(int64_t)l * n_q + q] : -1;
    int lo = k0 >= 0 ? k0 : INT32_MAX, hi = k0 >= 0 ? k0 + BAND_W : 0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { lo = min(lo, __shfl_xor_sync(IVX_FULL_MASK, lo, d)); hi = max(hi, __shfl_xor_sync(IVX_FULL_MASK, hi, d)); }
    int kb = 0, nks = 0;
    if (hi > 0) { kb = lo & ~1; hi = min(hi, n); nks = (hi - kb + 3) >> 2; }
    if (nks > MMA_KS) { if (lane == 0) s_bad = 1; nks = 0; }
    if (lane == 0) {
      oct[((int64_t)l * MMA_OCT + o) * 2] = kb; oct[((int64_t)l * MMA_OCT + o) * 2 + 1] = nks;
      s_kb[o] = nks > 0 ? kb : INT32_MAX; s_ke[o] = nks > 0 ? kb + 4 * nks : 0;
      atomicMax(&s_nks, nks);
    }
    const double* __restrict__ wq = band_w + ((int64_t)l * n_q + (q < n_q ? q : 0)) * BAND_W;
    for (int ks = 0; ks < MMA_KS; ++ks) {
      const int j = kb + 4 * ks + tig - k0;
      const double v = (k0 >= 0 && ks < nks && j >= 0 && j < BAND_W) ? wq[j] : 0.0;
      frag[(((int64_t)l * MMA_OCT + o) * MMA_KS + ks) * 32 + lane] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < MMA_OCT / 8) {                  // slab extents
    int lo = INT32_MAX, hi = 0;
    for (int o = 8 * threadIdx.x; o < 8 * threadIdx.x + 8; ++o) { lo = min(lo, s_kb[o]); hi = max(hi, s_ke[o]); }
    if (hi > 0) { if (hi - lo > MMA_SLAB_MAX) s_bad = 1; atomicMax(&s_ext, hi - lo); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (band_ok && !s_bad) ladder_ok[l] |= LADDER_OK_MMA;
    meta[2 * l] = s_ext; meta[2 * l + 1] = s_nks;
  }
}

// ---- apply -----------------------------------------------------------------------------------------------------------
struct BandParams {
  const int64_t* knot_off;       // [n_slices+1]
  const int32_t* order;          // slices grouped by ladder
  const int64_t* lad_slice_off;  // [n_ladders+1] into order
  const int64_t* ladder_off;     // [n_ladders+1] into x_ladder
  const double* x_ladder;
  const double* x_own;           // per-knot abscissae of the slices (verified against the ladder) or NULL
  const double* y;
  const int32_t* band_k0;
  const double* band_w;
  const uint8_t* ladder_ok;
  const int16_t* perm;           // [n_ladders][512] grid point of every thread (k_band_perm)
  double* yq;
  int32_t* fb_list;
  uint32_t* fb_count;
  int32_t n_q;
  int32_t batch;                 // slices per CTA
  // tensor-path form (k_band_apply_mma); frag == NULL: k_band_apply takes every ladder
  const double* frag;            // [n_ladders][64][MMA_KS][32] B fragments (k_band_frag)
  const int32_t* oct;            // [n_ladders][64][2] first knot and step count of every 8-point group
  int32_t row;                   // doubles per staged slice row (slab extent + padding), = 4 mod 16
  int32_t vec_ok;                // n_q even and yq 16-byte aligned
  const double* fill_w;          // missing-knot rows (k_band_fill_build) or NULL: a slice with a missing knot is flagged
  const int32_t* fill_k0;
  int32_t preflag;               // fb_list already flags slices that must take the exact solve (x verified by the caller): honoured, not cleared
};

template <bool VERIFY>
__global__ void __launch_bounds__(BAND_THREADS, 1) k_band_apply(BandParams P) {
  extern __shared__ __align__(16) double sm[];
  double* ybuf = sm;                                         // [NBUF][MAXK]
  double* obuf = ybuf + BAND_NBUF * BAND_MAXK;               // [2][MAXQ]
  double* xbuf = obuf + 2 * (BAND_MAXQ + 16);                // [NBUF][MAXK]   (VERIFY)
  double* ladx = xbuf + BAND_NBUF * BAND_MAXK;               // [MAXK]         (VERIFY)
  const int l = blockIdx.y;
  const int64_t s_begin = P.lad_slice_off[l] + (int64_t)blockIdx.x * P.batch;
  const int64_t s_end_all = P.lad_slice_off[l + 1];
  if (s_begin >= s_end_all) return;
  const int ns = (int)min((int64_t)P.batch, s_end_all - s_begin);
  const int32_t* __restrict__ ord = P.order + s_begin;
  const int t = threadIdx.x;
  if (P.frag && (P.ladder_ok[l] & LADDER_OK_MMA)) return;    // k_band_apply_mma takes this ladder
  if (!P.ladder_ok[l]) {                                     // no usable band for this ladder: everything to the exact solve
    for (int i = t; i < ns; i += BAND_THREADS)
      if (atomicExch(P.fb_list + ord[i], 1) == 0 || !(P.preflag || P.frag)) atomicAdd(P.fb_count, 1u);
    return;
  }
  const int n = (int)(P.ladder_off[l + 1] - P.ladder_off[l]);      // BAND_W <= n <= BAND_MAXK (k_band_build)
  const int n_q = P.n_q;
  // grid point of this thread: dealt so that the 16 lanes of a half-warp start their windows in 16 different banks
  const int q_raw = P.perm[(int64_t)l * BAND_THREADS + t];
  const bool has_q = q_raw >= 0;
  const int q = has_q ? q_raw : 0;
  double w[BAND_W];
  int k0 = 0;
  if (has_q) {
    k0 = P.band_k0[(int64_t)l * n_q + q];
    const double* __restrict__ wp = P.band_w + ((int64_t)l * n_q + q) * BAND_W;
#pragma unroll
    for (int j = 0; j < BAND_W; ++j) w[j] = __ldg(wp + j);
  } else {
#pragma unroll
    for (int j = 0; j < BAND_W; ++j) w[j] = 0.0;
  }
  const bool nan_q = k0 < 0;
  if (nan_q) k0 = 0;
  if (VERIFY) {
    for (int k = t; k < n; k += BAND_THREADS) ladx[k] = __ldg(P.x_ladder + P.ladder_off[l] + k);
  }
  // per-slice bookkeeping once per batch: first knot of slice i (or -1: another length than the ladder -> exact solve) and its id
  __shared__ int64_t sbase[BAND_BATCH];
  __shared__ int32_t sid[BAND_BATCH];
  if (t < ns) {
    const int32_t s = ord[t];
    const int64_t ko = P.knot_off[s];
    sid[t] = s;
    sbase[t] = ((P.knot_off[s + 1] - ko) == (int64_t)n && !(P.preflag && P.fb_list[s] != 0)) ? ko : (int64_t)-1;
  }
  __syncthreads();
  // knots this thread moves: k = t + r * 512, r = 0..3
  bool pr[BAND_MAXK / BAND_THREADS];
#pragma unroll
  for (int r = 0; r < BAND_MAXK / BAND_THREADS; ++r) pr[r] = t + r * BAND_THREADS < n;
  auto issue = [&](int i, int buf) {
    if (i < ns) {
      const int64_t ko = sbase[i];
      if (ko >= 0) {
        const double* __restrict__ ysrc = P.y + ko + t;
        double* yb = ybuf + buf * BAND_MAXK + t;
#pragma unroll
        for (int r = 0; r < BAND_MAXK / BAND_THREADS; ++r)
          if (pr[r]) cp_async8(yb + r * BAND_THREADS, ysrc + r * BAND_THREADS);
        if (VERIFY) {
          const double* __restrict__ xsrc = P.x_own + ko + t;
          double* xb = xbuf + buf * BAND_MAXK + t;
#pragma unroll
          for (int r = 0; r < BAND_MAXK / BAND_THREADS; ++r)
            if (pr[r]) cp_async8(xb + r * BAND_THREADS, xsrc + r * BAND_THREADS);
        }
      }
    }
    cp_async_commit();
  };
  issue(0, 0);
  issue(1, 1);
  bool prev_ok = false;                                      // CTA-uniform: slice i-1 was computed here (its results sit in obuf)
  int buf = 0, buf_next2 = 2;                                // buffers of slice i and of slice i+2 (= the one slice i-1 used)
  // obuf is written at the thread's grid point and read back in thread order: one pad double per 32 keeps both conflict-free
  const int oq = q + (q >> 5), ot = t + (t >> 5);
  for (int i = 0; i < ns; ++i) {
    cp_async_wait<1>();                                      // this thread's part of slice i has landed
    const bool len_ok = sbase[i] >= 0;
    const double* yb = ybuf + buf * BAND_MAXK;
    bool ok = len_ok;
    if (len_ok) {                                            // every thread checks the knots it moved itself
      const double* xb = xbuf + buf * BAND_MAXK;
#pragma unroll
      for (int r = 0; r < BAND_MAXK / BAND_THREADS; ++r) {
        const int k = t + r * BAND_THREADS;
        if (pr[r]) {
          ok = ok && (fabs(yb[k]) < INFINITY);
          if (VERIFY) ok = ok && (xb[k] == ladx[k]);        // ladx[k] was written by this same thread
        }
      }
    }
    const bool all_ok = __syncthreads_and(ok) != 0;          // the one barrier per slice: data visible, vote, buffers free
    issue(i + 2, buf_next2);                                 // into the buffer slice i-1 was read from
    if (i > 0 && prev_ok && t < n_q)                         // slice i-1: transposed through shared memory, stored coalesced
      st_stream(P.yq + (int64_t)sid[i - 1] * n_q + t, obuf[((i - 1) & 1) * (BAND_MAXQ + 16) + ot]);
    if (all_ok) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const double* yw = yb + k0;
#pragma unroll
      for (int j = 0; j < BAND_W; j += 4) {
        a0 = fma(w[j], yw[j], a0);
        a1 = fma(w[j + 1], yw[j + 1], a1);
        a2 = fma(w[j + 2], yw[j + 2], a2);
        a3 = fma(w[j + 3], yw[j + 3], a3);
      }
      if (has_q) obuf[(i & 1) * (BAND_MAXQ + 16) + oq] = nan_q ? qnan() : (a0 + a1) + (a2 + a3);
    }
    if (t == 0) {                                            // a missing knot, off the ladder, another length: flagged for the exact solve
      if (P.preflag || P.frag) { if (!all_ok && atomicExch(P.fb_list + sid[i], 1) == 0) atomicAdd(P.fb_count, 1u); }   // the list was zeroed or pre-set
      else { P.fb_list[sid[i]] = all_ok ? 0 : 1; if (!all_ok) atomicAdd(P.fb_count, 1u); }
    }
    prev_ok = all_ok;
    buf_next2 = buf;                                         // slice i+3 reuses slice i's buffer
    buf = buf + 1 == BAND_NBUF ? 0 : buf + 1;
  }
  __syncthreads();
  if (prev_ok && t < n_q) st_stream(P.yq + (int64_t)sid[ns - 1] * n_q + t, obuf[((ns - 1) & 1) * (BAND_MAXQ + 16) + ot]);
  cp_async_wait<0>();
}


// ---- apply, tensor path ----------------------------------------------------------------------------------------------
// The gaps of one staged row, by the warp that owns it: every NaN below nvalid is replaced by its fill row's 48-term dot product over
// the slice's other knots (weights and knots straight from global memory: the window may reach outside the slab).  Returns false when
// the rows cannot take the slice — an infinity, a gap near an end of the ladder, another gap inside a window — and the slice is left
// to the exact solve.  The row of knot m starts 24 knots before it (band_row on the ladder without m; k_band_fill_check holds the table
// to that), so nothing is looked up on this path, which is all latency.
__device__ __noinline__ bool fill_row_gaps(double* __restrict__ ybw, int n_chunks, int nvalid, int KB, int n, int64_t ko, int64_t lad_a0,
                                           const double* __restrict__ y, const double* __restrict__ fill_w) {
  const int lane = threadIdx.x & 31;
  const double2* yb = reinterpret_cast<const double2*>(ybw);
  bool row_ok = true;
  for (int r = 0; r < n_chunks; ++r) {
    const double2 yv = yb[lane + 32 * r];
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      const double v = comp ? yv.y : yv.x;
      uint32_t mask = __ballot_sync(IVX_FULL_MASK, ((uint32_t)__double2hiint(v) & 0x7ff00000u) == 0x7ff00000u);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int k = 2 * (src + 32 * r) + comp, m = KB + k;
        const double cur = ybw[k];                               // broadcast
        if (!(k < nvalid && isnan(cur) && m >= FILL_MARGIN && m <= n - 1 - FILL_MARGIN)) { row_ok = false; continue; }
        const int k0f = m - 24;
        const double* __restrict__ fw = fill_w + (lad_a0 + m) * BAND_W;
        double acc = 0.0;
        bool fin = true;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int j = lane + 32 * rr;
          if (j < BAND_W) {
            const int xi = k0f + j;
            const double yy = __ldg(y + ko + xi + (xi >= m ? 1 : 0));
            fin = fin && (((uint32_t)__double2hiint(yy) & 0x7ff00000u) != 0x7ff00000u);
            acc = fma(__ldg(fw + j), yy, acc);
          }
        }
        if (!__all_sync(IVX_FULL_MASK, fin)) { row_ok = false; continue; }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(IVX_FULL_MASK, acc, d);
        if (lane == 0) ybw[k] = acc;
      }
    }
  }
  return row_ok;
}

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8u(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// grid: x = 64-point slab, y = batch of MMA_BATCH slices, z = ladder; 256 threads; a staged row holds 32 NIT knots (+36 of padding:
// rows 4 mod 16 doubles apart keep the fragment loads conflict-free, and the multiply runs 20 or 24 steps whatever the group
// needs — the fragments past its last step are zero, and so is the row past its last knot).  NST stages of cp.async.  Warp w stages row w of every tile (one slice's slab of knots:
// cp.async, 16 bytes at a time where the slice's address allows), checks that row — every y finite, every x equal to the ladder's
// bit for bit, both as integer operations on the words — votes inside the warp, and multiplies the 8-point group 8 slab + w
// against all 8 rows.  Knots that no slab touches (the grid may be narrower than the ladder) are dealt evenly to the slabs and
// checked straight from global memory.
template <bool VERIFY, int NIT, int NST>
__global__ void __launch_bounds__(256, 2) k_band_apply_mma(BandParams P) {
  constexpr int ROW = 32 * NIT + 36;   // 20 steps from a group that starts 48 knots before the end of the slab read 31 knots past it
  extern __shared__ __align__(16) double sm[];
  double* ybuf = sm;                                         // [NST][8][ROW]
  double* xbuf = ybuf + NST * 8 * ROW;                       // [NST][8][ROW]   (VERIFY)
  double* ladx = xbuf + (VERIFY ? NST * 8 * ROW : 0);        // [ROW]           (VERIFY)
  __shared__ int s_kb[8], s_ke[8], s_kfirst, s_klast;
  __shared__ int32_t s_sid[NST][8];                          // slice of every row, -1 = no store (no such row, or flagged)
  __shared__ int64_t s_base[MMA_BATCH];                      // first knot of every slice of the batch, -1 = another length than the ladder
  __shared__ int32_t s_id[MMA_BATCH];
  const int l = blockIdx.z, slab = blockIdx.x;
  if (!(P.ladder_ok[l] & LADDER_OK_MMA)) return;
  const int64_t s_begin = P.lad_slice_off[l] + (int64_t)blockIdx.y * MMA_BATCH;
  const int64_t s_end_all = P.lad_slice_off[l + 1];
  if (s_begin >= s_end_all) return;
  const int ns = (int)min((int64_t)MMA_BATCH, s_end_all - s_begin);
  const int32_t* __restrict__ ord = P.order + s_begin;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const int n = (int)(P.ladder_off[l + 1] - P.ladder_off[l]);
  const int n_q = P.n_q;
  const int o = slab * 8 + w;                                // this warp's group of 8 grid points
  const int32_t* __restrict__ octl = P.oct + (int64_t)l * MMA_OCT * 2;
  const int kb = octl[o * 2], nks = octl[o * 2 + 1];
  if (lane == 0) { s_kb[w] = nks > 0 ? kb : INT32_MAX; s_ke[w] = nks > 0 ? kb + 4 * nks : 0; }
  if (w == 0) {                                              // knots before the first window and after the last one
    const unsigned b0 = __ballot_sync(IVX_FULL_MASK, octl[lane * 2 + 1] > 0);
    const unsigned b1 = __ballot_sync(IVX_FULL_MASK, octl[(32 + lane) * 2 + 1] > 0);
    if (lane == 0) {
      const int of = b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : -1);
      const int ol = b1 ? 63 - __clz(b1) : (b0 ? 31 - __clz(b0) : -1);
      s_kfirst = of >= 0 ? octl[of * 2] : n;                // no window at all: every knot is "before the first"
      s_klast = ol >= 0 ? min(n, octl[ol * 2] + 4 * octl[ol * 2 + 1]) : n;
    }
  }
  for (int i = threadIdx.x; i < ns; i += 256) {
    const int32_t sl = ord[i];
    const int64_t ko = P.knot_off[sl];
    s_id[i] = sl;
    s_base[i] = ((P.knot_off[sl + 1] - ko) == (int64_t)n && !(P.preflag && P.fb_list[sl] != 0)) ? ko : (int64_t)-1;
  }
  __syncthreads();
  int KB = INT32_MAX, KE = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { KB = min(KB, s_kb[i]); KE = max(KE, s_ke[i]); }
  if (KE == 0) KB = 0;                                       // every point of this slab is NaN by construction: nothing staged, NaNs written
  const int klen = KE - KB;                                  // <= MMA_SLAB_MAX (k_band_frag); the host chose NIT from the same tables
  if (klen > 32 * NIT) {                                     // the caller's hint was too small: the batch goes to the exact solve
    for (int i = threadIdx.x; i < ns; i += 256)
      if (atomicExch(P.fb_list + s_id[i], 1) == 0) atomicAdd(P.fb_count, 1u);
    return;
  }
  const int nvalid = min(klen, n - KB);                      // knots of the slab that exist; the rest of the row stays 0
  const int64_t lad_a0 = P.ladder_off[l];
  const bool fill_on = P.fill_w != nullptr && (P.ladder_ok[l] & LADDER_OK_FILL) != 0;
  // this slab's share of the knots no slab stages: [0, kfirst) and [klast, n), dealt evenly
  const int kfirst = s_kfirst, klast = s_klast;
  const int n_out = kfirst + (n - klast), per = (n_out + (int)gridDim.x - 1) / (int)gridDim.x;
  const int out_lo = min(n_out, slab * per), out_hi = min(n_out, out_lo + per);
  // B fragments of this warp's 8 points, for the whole batch
  double wreg[MMA_KS];
  {
    const double* __restrict__ fp = P.frag + (((int64_t)l * MMA_OCT + o) * MMA_KS) * 32 + lane;
#pragma unroll
    for (int ks = 0; ks < MMA_KS; ++ks) wreg[ks] = __ldg(fp + ks * 32);
  }
  // the two points of this lane's accumulator pair: 8o + 2 tig, + 1
  const int q0 = 8 * o + 2 * tig;
  const bool st0 = q0 < n_q, st1 = q0 + 1 < n_q;
  const bool nan0 = st0 && P.band_k0[(int64_t)l * n_q + q0] < 0, nan1 = st1 && P.band_k0[(int64_t)l * n_q + q0 + 1] < 0;
  const bool vec_out = P.vec_ok && st1;                       // 16-byte stores (n_q even, yq aligned)
  // zero the tail of every staged row once (cp.async never writes past nvalid, and the multiply reads up to 31 knots past the slab
  // against zero fragments: those knots must be finite); what lies below nvalid is overwritten by every tile — a row that never
  // receives a slice keeps whatever it held, feeds only its own outputs and is not stored
  for (int i = threadIdx.x; i < (VERIFY ? 2 : 1) * NST * 8 * (ROW - (nvalid & ~1)); i += 256) {
    const int row = i / (ROW - (nvalid & ~1)), k = (nvalid & ~1) + i % (ROW - (nvalid & ~1));
    ybuf[row * ROW + k] = 0.0;
  }
  if (VERIFY)
    for (int k = threadIdx.x; k < ROW; k += 256) ladx[k] = k < nvalid ? __ldg(P.x_ladder + P.ladder_off[l] + KB + k) : 0.0;
  __syncthreads();
  const int ntiles = (ns + 7) >> 3;
  // how many of this lane's 16-byte chunks (knots 2(lane + 32 r), +1) lie wholly inside the row, and whether one 8-byte knot follows
  const int full16 = nvalid / 2 > lane ? (nvalid / 2 - lane + 31) / 32 : 0;
  const bool tail8 = (nvalid & 1) && ((nvalid / 2) & 31) == lane;
  const int tail_k = nvalid - 1;
  const int full8 = nvalid > lane ? (nvalid - lane + 31) / 32 : 0;
  auto issue = [&](int tile, int buf) {                      // row w of the tile
    const int idx = tile * 8 + w;
    const int64_t ko = idx < ns ? s_base[idx] : (int64_t)-1;
    if (ko >= 0) {
      const double* __restrict__ ysrc = P.y + ko + KB;
      const double* __restrict__ xsrc = VERIFY ? P.x_own + ko + KB : nullptr;
      const uint32_t yb = smem_u32(ybuf + (buf * 8 + w) * ROW), xb = smem_u32(xbuf + (buf * 8 + w) * ROW);
      const bool vec = (((uintptr_t)ysrc | (VERIFY ? (uintptr_t)xsrc : 0)) & 15u) == 0;      // warp-uniform
      if (vec) {
#pragma unroll
        for (int r = 0; r < NIT / 2; ++r)
          if (r < full16) {
            const int k = 2 * (lane + 32 * r);
            cp_async16(yb + 8 * k, ysrc + k);
            if (VERIFY) cp_async16(xb + 8 * k, xsrc + k);
          }
        if (tail8) { cp_async8u(yb + 8 * tail_k, ysrc + tail_k); if (VERIFY) cp_async8u(xb + 8 * tail_k, xsrc + tail_k); }
      } else {
#pragma unroll
        for (int r = 0; r < NIT; ++r)
          if (r < full8) {
            const int k = lane + 32 * r;
            cp_async8u(yb + 8 * k, ysrc + k);
            if (VERIFY) cp_async8u(xb + 8 * k, xsrc + k);
          }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int t0 = 0; t0 < NST - 1; ++t0) issue(t0, t0);        // (a tile past the end commits an empty group)
  // the lane's first knot outside every slab, if it has one: loaded one tile ahead so that the load travels under the multiply
  const bool has_out = out_lo + lane < out_hi;
  const int out_k = has_out ? (out_lo + lane < kfirst ? out_lo + lane : klast + (out_lo + lane - kfirst)) : 0;
  const double out_lx = (VERIFY && has_out) ? __ldg(P.x_ladder + P.ladder_off[l] + out_k) : 0.0;
  double ey = 0.0, ex = 0.0;
  auto load_out = [&](int tile) {
    const int idx = tile * 8 + w;
    const int64_t ko = idx < ns ? s_base[idx] : (int64_t)-1;
    if (has_out && ko >= 0) { ey = __ldg(P.y + ko + out_k); if (VERIFY) ex = __ldg(P.x_own + ko + out_k); }
  };
  load_out(0);
  int buf = 0;
  for (int tile = 0; tile < ntiles; ++tile) {
    {
      const int idx = tile * 8 + w;
      const int64_t ko = idx < ns ? s_base[idx] : (int64_t)-1;
      // knots outside every slab: the lane's first one was loaded a tile ago (ey, ex), any further ones are loaded here
      uint32_t bad = 0;
      if (ko >= 0) {
        if (has_out) {
          if (!fill_on) bad |= ((uint32_t)__double2hiint(ey) & 0x7ff00000u) == 0x7ff00000u;   // (no window reaches such a knot: with the fill rows a gap there is nobody's business)
          if (VERIFY) bad |= (uint32_t)((__double2hiint(ex) ^ __double2hiint(out_lx)) | (__double2loint(ex) ^ __double2loint(out_lx)));
        }
        for (int i = out_lo + lane + 32; i < out_hi; i += 32) {
          const int k = i < kfirst ? i : klast + (i - kfirst);
          const double yv = __ldg(P.y + ko + k);
          if (!fill_on) bad |= ((uint32_t)__double2hiint(yv) & 0x7ff00000u) == 0x7ff00000u;
          if (VERIFY) {
            const double xv = __ldg(P.x_own + ko + k), lv = __ldg(P.x_ladder + P.ladder_off[l] + k);
            bad |= (uint32_t)((__double2hiint(xv) ^ __double2hiint(lv)) | (__double2loint(xv) ^ __double2loint(lv)));
          }
        }
      }
      cp_async_wait<NST - 2>();                              // the lanes' parts of row w have landed ...
      __syncwarp();                                          // ... and are visible to the whole warp
      if (ko >= 0) {
        const double2* yb = reinterpret_cast<const double2*>(ybuf + (buf * 8 + w) * ROW);
        const double2* xb = reinterpret_cast<const double2*>(xbuf + (buf * 8 + w) * ROW);
        const double2* lx = reinterpret_cast<const double2*>(ladx);
        uint32_t emax = 0;                                   // largest exponent field seen: 0x7ff = NaN or infinity (zeros past nvalid pass)
#pragma unroll
        for (int r = 0; r < NIT / 2; ++r) {
          const int c = lane + 32 * r;
          const double2 yv = yb[c];
          emax = max(emax, max((uint32_t)__double2hiint(yv.x) & 0x7ff00000u, (uint32_t)__double2hiint(yv.y) & 0x7ff00000u));
          if (VERIFY) {
            const double2 xv = xb[c], lv = lx[c];
            bad |= (uint32_t)((__double2hiint(xv.x) ^ __double2hiint(lv.x)) | (__double2loint(xv.x) ^ __double2loint(lv.x)) |
                              (__double2hiint(xv.y) ^ __double2hiint(lv.y)) | (__double2loint(xv.y) ^ __double2loint(lv.y)));
          }
        }
        bool nonfin = emax == 0x7ff00000u;
        // a missing quote (NaN) in the staged row: filled, out of line (the clean day never goes there and its loop keeps its registers)
        if (fill_on && __any_sync(IVX_FULL_MASK, nonfin)) nonfin = !fill_row_gaps(ybuf + (buf * 8 + w) * ROW, NIT / 2, nvalid, KB, n, ko, lad_a0, P.y, P.fill_w);
        bad |= nonfin;
      }
      const bool ok = (ko >= 0) && __all_sync(IVX_FULL_MASK, bad == 0);
      if (lane == 0) {
        const int32_t sl = idx < ns ? s_id[idx] : -1;
        if (idx < ns && !ok && atomicExch(P.fb_list + sl, 1) == 0) atomicAdd(P.fb_count, 1u);   // for the exact solve, counted once
        s_sid[buf][w] = ok ? sl : -1;
      }
    }
    __syncthreads();                                         // rows and verdicts of this tile visible; the stage of tile - 1 is free
    {
      const int nb = buf == 0 ? NST - 1 : buf - 1;           // (tile + NST - 1) mod NST
      issue(tile + NST - 1, nb);
      load_out(tile + 1);
    }
    // 8 slices x 8 points: two accumulator chains, 20 or 24 steps (zero fragments and zero knots past the group's own)
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    const double* arow = ybuf + (buf * 8 + g) * ROW + (nks > 0 ? kb - KB : 0) + tig;   // (a group without a point multiplies zeros)
#pragma unroll
    for (int ks = 0; ks < 20; ks += 2) {
      dmma_884(c0, c1, arow[4 * ks], wreg[ks]);
      dmma_884(d0, d1, arow[4 * ks + 4], wreg[ks + 1]);
    }
    if (nks > 20) {
#pragma unroll
      for (int ks = 20; ks < MMA_KS; ks += 2) {
        dmma_884(c0, c1, arow[4 * ks], wreg[ks]);
        dmma_884(d0, d1, arow[4 * ks + 4], wreg[ks + 1]);
      }
    }
    const int32_t sl = s_sid[buf][g];
    if (sl >= 0) {
      double r0 = c0 + d0, r1 = c1 + d1;
      if (nan0) r0 = qnan();
      if (nan1) r1 = qnan();
      double* dst = P.yq + (int64_t)sl * n_q + q0;
      if (vec_out) __stcs(reinterpret_cast<double2*>(dst), make_double2(r0, r1));
      else { if (st0) st_stream(dst, r0); if (st1) st_stream(dst + 1, r1); }
    }
    buf = buf + 1 == NST ? 0 : buf + 1;
  }
  cp_async_wait<0>();
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_spline_band_len(int64_t n_ladders, int64_t n_q) {
  if (n_ladders < 0 || n_q < 0) return -1;
  return n_ladders * n_q * BAND_W;
}

extern "C" int64_t ivx_spline_band_frag_len(int64_t n_ladders) {
  if (n_ladders < 0) return -1;
  return n_ladders * MMA_OCT * MMA_KS * 32;
}

// int32 entries of band_oct: {first knot, steps} per (ladder, group of 8 points), then {largest slab extent, largest step count} per ladder
extern "C" int64_t ivx_spline_band_oct_len(int64_t n_ladders) {
  if (n_ladders < 0) return -1;
  return n_ladders * (MMA_OCT * 2 + 2);
}

// doubles of band_fill_w for ladders with n_ladder_knots knots in all (band_fill_k0: one int32 per knot)
extern "C" int64_t ivx_spline_band_fill_len(int64_t n_ladder_knots) {
  if (n_ladder_knots < 0) return -1;
  return n_ladder_knots * BAND_W;
}

extern "C" int ivx_spline_band_build(int64_t n_ladders, const int64_t* ladder_off, const double* x_ladder, int64_t n_q, const double* xq,
                                     int32_t extrap, int32_t* band_k0, double* band_w, uint8_t* ladder_ok, int16_t* band_perm,
                                     double* band_frag, int32_t* band_oct, double* band_fill_w, int32_t* band_fill_k0, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_ladders >= 0 && n_ladders < 65536 && n_q >= 0 && n_q < ((int64_t)1 << 30), IVX_ERR_INVALID_ARG, "ivx_spline_band_build: bad sizes");
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_band_build: extrap must be 0, 1 or 2");
  if (n_ladders == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(ladder_off && x_ladder && xq && band_k0 && band_w && ladder_ok && band_perm, IVX_ERR_INVALID_ARG, "ivx_spline_band_build: null pointer");
  k_band_build<<<dim3((unsigned)((n_q + 127) / 128), (unsigned)n_ladders), 128, 0, stream>>>((int32_t)n_ladders, ladder_off, x_ladder, (int32_t)n_q, xq, extrap,
                                                                                             band_k0, band_w, ladder_ok);
  IVX_LAUNCH_CHECK("k_band_build");
  if (n_q <= BAND_MAXQ) {
    k_band_check<<<(unsigned)n_ladders, 256, 0, stream>>>(ladder_off, x_ladder, (int32_t)n_q, xq, extrap, band_k0, band_w, ladder_ok);
    IVX_LAUNCH_CHECK("k_band_check");
    k_band_perm<<<(unsigned)n_ladders, 32, 0, stream>>>((int32_t)n_q, band_k0, band_perm);
    IVX_LAUNCH_CHECK("k_band_perm");
    if (band_frag && band_oct) {
      k_band_frag<<<(unsigned)n_ladders, 256, 0, stream>>>((int32_t)n_q, ladder_off, band_k0, band_w, ladder_ok, band_frag, band_oct,
                                                          band_oct + n_ladders * MMA_OCT * 2);
      IVX_LAUNCH_CHECK("k_band_frag");
    }
    if (band_fill_w && band_fill_k0) {
      // ladders hold at most BAND_MAXK knots where the band is usable; longer ones have no band and their threads return at once
      k_band_fill_build<<<dim3((BAND_MAXK + 127) / 128, (unsigned)n_ladders), 128, 0, stream>>>(ladder_off, x_ladder, ladder_ok, band_fill_k0, band_fill_w);
      IVX_LAUNCH_CHECK("k_band_fill_build");
      k_band_fill_check<<<dim3(FILL_SAMPLES, (unsigned)n_ladders), 64, 0, stream>>>(ladder_off, x_ladder, band_fill_k0, band_fill_w, ladder_ok);
      IVX_LAUNCH_CHECK("k_band_fill_check");
    }
  }
  return IVX_OK;
}

extern "C" int ivx_spline_fit_eval_banded(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, const int32_t* slice_ladder,
                                          const int32_t* order, const int64_t* lad_slice_off, int64_t n_ladders, int64_t max_slices_per_ladder,
                                          const int64_t* ladder_off, int64_t max_knots, const double* x_ladder, const double* x_own,
                                          const double* lu_cp, const double* lu_inv, const double* lu_ih, const double* y, int64_t n_q,
                                          const double* xq, int32_t extrap, const int32_t* band_k0, const double* band_w,
                                          const uint8_t* ladder_ok, const int16_t* band_perm, const double* band_frag, const int32_t* band_oct,
                                          const double* band_fill_w, const int32_t* band_fill_k0, int32_t slab_knots_hint, int32_t all_tensor_path, int32_t x_verified, double* yq, int32_t* fb_list, uint32_t* fb_count,
                                          ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_slices >= 0 && n_slices < ((int64_t)1 << 31) && n_ladders >= 0 && n_ladders < 65536 && max_slices_per_ladder >= 0,
              IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval_banded: bad sizes");
  IVX_REQUIRE(n_q >= 0 && n_q <= BAND_MAXQ, IVX_ERR_UNSUPPORTED, "ivx_spline_fit_eval_banded: at most %d grid points (got %lld)", BAND_MAXQ, (long long)n_q);
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval_banded: extrap must be 0, 1 or 2");
  if (n_slices == 0 || n_q == 0 || n_ladders == 0) return IVX_OK;
  IVX_REQUIRE(knot_off && x_off && slice_ladder && order && lad_slice_off && ladder_off && x_ladder && lu_cp && lu_inv && lu_ih && y && xq && band_k0 &&
                  band_w && ladder_ok && band_perm && yq && fb_list && fb_count,
              IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval_banded: null pointer");
  // x_verified: the caller (ivx_bs_implied_vol_ladder) compared every slice's abscissae with its ladder's already and left the
  // verdicts in fb_list / fb_count — kept, added to, and x_own is read by the exact solve of the flagged slices only
  if (!x_verified) IVX_CUDA(cudaMemsetAsync(fb_count, 0, 4, stream));
  const double* x_check = x_verified ? nullptr : x_own;
  IVX_REQUIRE((band_frag == nullptr) == (band_oct == nullptr) && slab_knots_hint >= 0, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval_banded: band_frag and band_oct go together");
  BandParams P;
  P.knot_off = knot_off; P.order = order; P.lad_slice_off = lad_slice_off; P.ladder_off = ladder_off; P.x_ladder = x_ladder; P.x_own = x_check;
  P.preflag = x_verified ? 1 : 0;
  P.y = y; P.band_k0 = band_k0; P.band_w = band_w; P.ladder_ok = ladder_ok; P.perm = band_perm; P.yq = yq; P.fb_list = fb_list; P.fb_count = fb_count; P.n_q = (int32_t)n_q;
  P.fill_w = (band_fill_w && band_fill_k0) ? band_fill_w : nullptr; P.fill_k0 = band_fill_k0;
  P.frag = band_frag; P.oct = band_oct; P.row = 0; P.vec_ok = ((n_q & 1) == 0 && ((uintptr_t)yq & 15u) == 0) ? 1 : 0;
  if (band_frag) {
    // tensor-path form for the ladders whose tables allow it (LADDER_OK_MMA); k_band_apply below skips those.  A row holds the
    // knots one 64-point slab touches: the caller's hint (meta of ivx_spline_band_build) or the most the tables can ask for.
    const int64_t ext = slab_knots_hint > 0 && slab_knots_hint <= MMA_SLAB_MAX ? slab_knots_hint : MMA_SLAB_MAX;
    const int nit = ext <= 320 ? 10 : 16;
    P.row = 32 * nit + 36;
    if (!x_verified) IVX_CUDA(cudaMemsetAsync(fb_list, 0, (size_t)n_slices * sizeof(int32_t), stream));
    const dim3 mgrid((unsigned)((n_q + 63) / 64), (unsigned)((max_slices_per_ladder + MMA_BATCH - 1) / MMA_BATCH), (unsigned)n_ladders);
    if (mgrid.y > 0 && mgrid.y < 65536) {
      // stages: as many as let two CTAs share an SM
#define IVX_MMA_LAUNCH(V, N, S)                                                                                                       \
  {                                                                                                                                   \
    const size_t smem = (size_t)(((V) ? 2 * (S) * 8 + 1 : (S) * 8) * P.row) * sizeof(double);                                         \
    static DeviceOnce once;                                                                                                           \
    if (once_per_device(once)) IVX_CUDA(cudaFuncSetAttribute(k_band_apply_mma<V, N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
    k_band_apply_mma<V, N, S><<<mgrid, 256, smem, stream>>>(P);                                                                       \
  }
      if (x_check) { if (nit == 10) IVX_MMA_LAUNCH(true, 10, 2) else IVX_MMA_LAUNCH(true, 16, 2) }
      else { if (nit == 10) IVX_MMA_LAUNCH(false, 10, 4) else IVX_MMA_LAUNCH(false, 16, 3) }
#undef IVX_MMA_LAUNCH
      IVX_LAUNCH_CHECK("k_band_apply_mma");
    } else {
      P.frag = nullptr;                                // more batches than a grid dimension holds: the dot-product form takes everything
    }
  }
  // slices per CTA: as many as 48 (the weights are loaded once per CTA), but no fewer than ~3 CTAs per SM in the launch
  int64_t batch = (max_slices_per_ladder * n_ladders + 3 * (int64_t)sm_count() - 1) / (3 * (int64_t)sm_count());
  batch = batch < 4 ? 4 : (batch > BAND_BATCH ? BAND_BATCH : batch);
  P.batch = (int32_t)batch;
  const dim3 grid((unsigned)((max_slices_per_ladder + batch - 1) / batch), (unsigned)n_ladders);
  // all_tensor_path: the caller read ladder_ok back after the build and every ladder has its tensor-path tables — the
  // dot-product kernel would only launch CTAs that leave at once
  if (grid.x > 0 && !(P.frag && all_tensor_path)) {
    if (x_check) {
      const size_t smem = (size_t)(2 * BAND_NBUF * BAND_MAXK + 2 * (BAND_MAXQ + 16) + BAND_MAXK) * sizeof(double);
      static DeviceOnce once;
      if (once_per_device(once)) IVX_CUDA(cudaFuncSetAttribute(k_band_apply<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_band_apply<true><<<grid, BAND_THREADS, smem, stream>>>(P);
    } else {
      const size_t smem = (size_t)(BAND_NBUF * BAND_MAXK + 2 * (BAND_MAXQ + 16)) * sizeof(double);
      static DeviceOnce once;
      if (once_per_device(once)) IVX_CUDA(cudaFuncSetAttribute(k_band_apply<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_band_apply<false><<<grid, BAND_THREADS, smem, stream>>>(P);
    }
    IVX_LAUNCH_CHECK("k_band_apply");
  }
  // whatever the banded kernel listed: the exact solve (ladder tables where a slice may use them), same stream
  return spline_fit_eval_list(n_slices, fb_list, fb_count, max_knots, knot_off, x_off, x_ladder, lu_cp, lu_inv, lu_ih, slice_ladder, x_own, ladder_off, y,
                              (int)n_q, xq, extrap, yq, stream);
}
