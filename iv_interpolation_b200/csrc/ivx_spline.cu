// ivx_spline.cu — kernel D: batched natural cubic spline fit + grid evaluation.
//
// North_star kernel 2 (the reference has no spline code): per (timestamp, expiry)
// slice, fit the natural cubic spline through the quoted (strike, IV) knots and evaluate it on a
// shared strike grid; then the same across maturities for every grid strike.
//
// ivx_spline_fit_eval — one WARP per slice, slice staged in shared memory (x, y, c', d'/M: 32 B per
//   knot, so a 2000-knot slice is 64 KB and three slices are resident per SM).  The second
//   derivatives M solve the tridiagonal system  h[i-1] M[i-1] + 2(h[i-1]+h[i]) M[i] + h[i] M[i+1] =
//   6 (dy[i]/h[i] - dy[i-1]/h[i-1]),  M[0] = M[n-1] = 0.  The system is strictly diagonally
//   dominant, so both Thomas recurrences forget their start geometrically (c' contracts by
//   <= 0.083 per row, the right-hand-side recurrence by a_i/den_i, 0.27 on a uniform grid).  Each
//   lane therefore owns a contiguous chunk of rows and warms its recurrences up over a HALO of rows
//   before (forward) / after (backward) the chunk: 32-way parallel with O(n) work per slice —
//   a cyclic-reduction solve would need 2 fp64 divisions per row per level, ~10x the flops, and
//   this kernel is fp64-pipe bound, not HBM bound, at 2000 knots.
// ivx_spline_strided — one THREAD per system for the many tiny systems of the maturity direction,
//   data laid out [knot][system] so every load/store is coalesced; when no knot is missing the LU
//   factors depend only on the shared abscissae and are computed once per CTA in shared memory.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr int HALO = 64;
constexpr int SPL_MAX_WARPS = 8;
constexpr int STRIDED_MAXK = 64;

__device__ __forceinline__ double spline_eval(double xl, double xr, double yl, double yr, double Ml, double Mr, double xq) {
  const double h = xr - xl;
  const double A = xr - xq, B = xq - xl;
  const double ih = 1.0 / h;
  return (Ml * A * A * A + Mr * B * B * B) * (ih * (1.0 / 6.0)) + (yl * ih - Ml * h * (1.0 / 6.0)) * A +
         (yr * ih - Mr * h * (1.0 / 6.0)) * B;
}

__device__ __forceinline__ double apply_extrap(double v, double xq, double xlo, double xhi, double ylo, double yhi, int extrap) {
  if (isnan(xq)) return qnan();
  if (xq < xlo) return extrap == 0 ? qnan() : (extrap == 2 ? ylo : v);
  if (xq > xhi) return extrap == 0 ? qnan() : (extrap == 2 ? yhi : v);
  return v;
}

__global__ void __launch_bounds__(SPL_MAX_WARPS * 32) k_spline_fit_eval(int64_t n_slices, const int64_t* __restrict__ knot_off,
                                                                       const double* __restrict__ x, const double* __restrict__ y,
                                                                       int n_q, const double* __restrict__ xq, int extrap,
                                                                       double* __restrict__ yq, int mk,
                                                                       unsigned long long* counter) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* xs = smem + (size_t)warp * 4 * mk;
  double* ys = xs + mk;
  double* cp = ys + mk;
  double* dp = cp + mk;

  while (true) {
    unsigned long long sl = 0;
    if (lane == 0) sl = atomicAdd(counter, 1ull);
    sl = __shfl_sync(IVX_FULL_MASK, sl, 0);
    if ((int64_t)sl >= n_slices) break;
    const int64_t k0 = knot_off[sl], k1 = knot_off[sl + 1];
    double* out = yq + (int64_t)sl * n_q;

    // ---- stage + compact the usable knots ---------------------------------------------------------
    int n = 0;
    for (int64_t base = k0; base < k1; base += 32) {
      const int64_t i = base + lane;
      double xv = qnan(), yv = qnan();
      if (i < k1) { xv = ld_stream(x + i); yv = ld_stream(y + i); }
      const bool ok = isfinite(xv) && isfinite(yv);
      const uint32_t m = __ballot_sync(IVX_FULL_MASK, ok);
      if (ok) {
        const int j = n + __popc(m & ((1u << lane) - 1u));
        xs[j] = xv; ys[j] = yv;
      }
      n += __popc(m);
    }
    __syncwarp();

    if (n < 2) {
      for (int q = lane; q < n_q; q += 32) st_stream(out + q, qnan());
      __syncwarp();
      continue;
    }
    const double xlo = xs[0], xhi = xs[n - 1], ylo = ys[0], yhi = ys[n - 1];

    if (n > 2) {
      // rows 1..n-2, split in contiguous chunks over the lanes
      const int m = n - 2;
      const int chunk = (m + 31) / 32;
      const int first = 1 + lane * chunk;
      const int last = min(first + chunk - 1, n - 2);   // inclusive; empty when first > n-2
      if (first <= n - 2) {
        int i = max(1, first - HALO);
        double c_prev = 0.0, d_prev = 0.0;
        double xm = xs[i - 1], x0 = xs[i], ym = ys[i - 1], y0 = ys[i];
        double hm = x0 - xm, sm = (y0 - ym) / hm;
        for (; i <= last; ++i) {
          const double xp = xs[i + 1], yp = ys[i + 1];
          const double hp = xp - x0;
          const double sp = (yp - y0) / hp;
          const double inv = 1.0 / (2.0 * (hm + hp) - hm * c_prev);
          c_prev = hp * inv;
          d_prev = (6.0 * (sp - sm) - hm * d_prev) * inv;
          if (i >= first) { cp[i] = c_prev; dp[i] = d_prev; }
          x0 = xp; y0 = yp; hm = hp; sm = sp;
        }
      }
      __syncwarp();
      // backward: warm up over the rows after the chunk (owned by higher lanes), then own rows in place
      double M = 0.0;
      if (first <= n - 2) {
        for (int i = min(n - 2, last + HALO); i > last; --i) M = dp[i] - cp[i] * M;
      }
      __syncwarp();
      if (first <= n - 2) {
        for (int i = last; i >= first; --i) { M = dp[i] - cp[i] * M; dp[i] = M; }
      }
      if (lane == 0) { dp[0] = 0.0; dp[n - 1] = 0.0; }
    } else {
      if (lane == 0) { dp[0] = 0.0; dp[1] = 0.0; }     // two knots: the straight line
    }
    __syncwarp();

    // ---- evaluate on the shared grid ---------------------------------------------------------------
    for (int q = lane; q < n_q; q += 32) {
      const double xv = __ldg(xq + q);
      int lo = 0, hi = n - 1;                            // last i in [0, n-2] with xs[i] <= xv
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xs[mid] <= xv) lo = mid; else hi = mid;
      }
      const double v = spline_eval(xs[lo], xs[lo + 1], ys[lo], ys[lo + 1], dp[lo], dp[lo + 1], xv);
      st_stream(out + q, apply_extrap(v, xv, xlo, xhi, ylo, yhi, extrap));
    }
    __syncwarp();
  }
}

// ---- many small systems, [knot][system] layout ----------------------------------------------------------
// blockIdx.y = batch: batch b has its own abscissae x[b*x_batch_stride + i*x_knot_stride] and its own
// [knot][system] / [query][system] planes of y / yq
__global__ void __launch_bounds__(128) k_spline_strided(int64_t n_sys, int nk, const double* __restrict__ x,
                                                        int64_t x_batch_stride, int64_t x_knot_stride,
                                                        const double* __restrict__ y, int n_q,
                                                        const double* __restrict__ xq, int extrap,
                                                        double* __restrict__ yq) {
  __shared__ double s_x[STRIDED_MAXK], s_cp[STRIDED_MAXK], s_inv[STRIDED_MAXK];
  __shared__ int s_x_ok;
  x += (int64_t)blockIdx.y * x_batch_stride;
  y += (int64_t)blockIdx.y * nk * n_sys;
  yq += (int64_t)blockIdx.y * n_q * n_sys;
  // LU factors of the shared abscissae (valid when a system has every knot)
  if (threadIdx.x == 0) {
    bool okx = true;
    for (int i = 0; i < nk; ++i) { s_x[i] = x[i * x_knot_stride]; okx = okx && isfinite(s_x[i]); }
    s_x_ok = okx;
    double c_prev = 0.0;
    for (int i = 1; i + 1 < nk; ++i) {
      const double hm = s_x[i] - s_x[i - 1], hp = s_x[i + 1] - s_x[i];
      const double inv = 1.0 / (2.0 * (hm + hp) - hm * c_prev);
      c_prev = hp * inv;
      s_cp[i] = c_prev; s_inv[i] = inv;
    }
  }
  __syncthreads();
  const int64_t sys = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (sys >= n_sys) return;

  double yk[STRIDED_MAXK];     // y, then compacted y
  double Mk[STRIDED_MAXK];
  bool all_ok = s_x_ok != 0;
  for (int i = 0; i < nk; ++i) {
    yk[i] = ld_stream(y + (int64_t)i * n_sys + sys);
    all_ok = all_ok && isfinite(yk[i]);
  }

  if (all_ok && nk >= 3) {
    // shared-LU path: only the right-hand side is per system
    double d_prev = 0.0;
    double sm = (yk[1] - yk[0]) / (s_x[1] - s_x[0]);
    for (int i = 1; i + 1 < nk; ++i) {
      const double hm = s_x[i] - s_x[i - 1];
      const double sp = (yk[i + 1] - yk[i]) / (s_x[i + 1] - s_x[i]);
      d_prev = (6.0 * (sp - sm) - hm * d_prev) * s_inv[i];
      Mk[i] = d_prev;
      sm = sp;
    }
    double M = 0.0;
    Mk[nk - 1] = 0.0;
    for (int i = nk - 2; i >= 1; --i) { M = Mk[i] - s_cp[i] * M; Mk[i] = M; }
    Mk[0] = 0.0;
    int j = 0;
    for (int q = 0; q < n_q; ++q) {
      const double xv = __ldg(xq + q);
      while (j < nk - 2 && xv >= s_x[j + 1]) ++j;
      const double v = spline_eval(s_x[j], s_x[j + 1], yk[j], yk[j + 1], Mk[j], Mk[j + 1], xv);
      st_stream(yq + (int64_t)q * n_sys + sys, apply_extrap(v, xv, s_x[0], s_x[nk - 1], yk[0], yk[nk - 1], extrap));
    }
    return;
  }

  // generic path: drop unusable knots, private abscissae
  double xk[STRIDED_MAXK], ck[STRIDED_MAXK];
  int n = 0;
  for (int i = 0; i < nk; ++i) {
    const double xv = s_x[i], yv = yk[i];
    if (isfinite(xv) && isfinite(yv)) { xk[n] = xv; yk[n] = yv; ++n; }
  }
  if (n < 2) {
    for (int q = 0; q < n_q; ++q) st_stream(yq + (int64_t)q * n_sys + sys, qnan());
    return;
  }
  Mk[0] = 0.0; Mk[n - 1] = 0.0;
  if (n > 2) {
    double c_prev = 0.0, d_prev = 0.0;
    double sm = (yk[1] - yk[0]) / (xk[1] - xk[0]);
    for (int i = 1; i + 1 < n; ++i) {
      const double hm = xk[i] - xk[i - 1], hp = xk[i + 1] - xk[i];
      const double sp = (yk[i + 1] - yk[i]) / hp;
      const double inv = 1.0 / (2.0 * (hm + hp) - hm * c_prev);
      c_prev = hp * inv;
      d_prev = (6.0 * (sp - sm) - hm * d_prev) * inv;
      ck[i] = c_prev; Mk[i] = d_prev;
      sm = sp;
    }
    double M = 0.0;
    for (int i = n - 2; i >= 1; --i) { M = Mk[i] - ck[i] * M; Mk[i] = M; }
  }
  int j = 0;
  for (int q = 0; q < n_q; ++q) {
    const double xv = __ldg(xq + q);
    while (j < n - 2 && xv >= xk[j + 1]) ++j;
    const double v = spline_eval(xk[j], xk[j + 1], yk[j], yk[j + 1], Mk[j], Mk[j + 1], xv);
    st_stream(yq + (int64_t)q * n_sys + sys, apply_extrap(v, xv, xk[0], xk[n - 1], yk[0], yk[n - 1], extrap));
  }
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_spline_workspace_bytes(int64_t, int64_t) { return 256; }

extern "C" int ivx_spline_fit_eval(int64_t n_slices, const int64_t* knot_off, int64_t max_knots, const double* x,
                                   const double* y, int64_t n_q, const double* xq, int32_t extrap, double* yq,
                                   void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_slices >= 0 && max_knots >= 0 && n_q >= 0 && n_q < ((int64_t)1 << 30), IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: bad sizes");
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: extrap must be 0, 1 or 2");
  if (n_slices == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(knot_off && x && y && xq && yq && workspace && workspace_bytes >= 256, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: null pointer or workspace < 256 B");
  int mk = (int)((max_knots + 3) / 2 * 2);   // room for row n+1 reads, even for 16-byte alignment
  if (mk < 4) mk = 4;
  const size_t per_warp = (size_t)mk * 4 * sizeof(double);
  IVX_REQUIRE(per_warp <= 227 * 1024, IVX_ERR_UNSUPPORTED, "ivx_spline_fit_eval: max_knots %lld needs %zu B of shared memory per slice (limit 227 KB)", (long long)max_knots, per_warp);
  int warps = (int)((227 * 1024) / per_warp);
  if (warps > SPL_MAX_WARPS) warps = SPL_MAX_WARPS;
  const size_t smem = per_warp * warps;
  IVX_CUDA(cudaFuncSetAttribute(k_spline_fit_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm * warps > 48) ctas_per_sm = 48 / warps;
  int64_t grid = (int64_t)sm_count() * ctas_per_sm;
  const int64_t need = (n_slices + warps - 1) / warps;
  if (grid > need) grid = need;
  unsigned long long* counter = (unsigned long long*)workspace;
  IVX_CUDA(cudaMemsetAsync(counter, 0, 8, stream));
  k_spline_fit_eval<<<(unsigned)grid, warps * 32, smem, stream>>>(n_slices, knot_off, x, y, (int)n_q, xq, extrap, yq, mk, counter);
  IVX_LAUNCH_CHECK("k_spline_fit_eval");
  return IVX_OK;
}

extern "C" int ivx_spline_strided(int64_t n_systems, int32_t n_knots, const double* x, const double* y, int32_t n_q,
                                  const double* xq, int32_t extrap, double* yq, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_systems >= 0 && n_knots >= 0 && n_q >= 0, IVX_ERR_INVALID_ARG, "ivx_spline_strided: bad sizes");
  IVX_REQUIRE(n_knots <= STRIDED_MAXK, IVX_ERR_UNSUPPORTED, "ivx_spline_strided: at most %d knots per system (got %d)", STRIDED_MAXK, n_knots);
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_strided: extrap must be 0, 1 or 2");
  if (n_systems == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(x && y && xq && yq, IVX_ERR_INVALID_ARG, "ivx_spline_strided: null pointer");
  k_spline_strided<<<(unsigned)((n_systems + 127) / 128), 128, 0, stream>>>(n_systems, n_knots, x, 0, 1, y, n_q, xq, extrap, yq);
  IVX_LAUNCH_CHECK("k_spline_strided");
  return IVX_OK;
}

extern "C" int ivx_spline_strided_batched(int64_t n_batches, int64_t n_systems, int32_t n_knots, const double* x,
                                          int64_t x_batch_stride, int64_t x_knot_stride, const double* y, int32_t n_q,
                                          const double* xq, int32_t extrap, double* yq, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_batches >= 0 && n_systems >= 0 && n_knots >= 0 && n_q >= 0, IVX_ERR_INVALID_ARG, "ivx_spline_strided_batched: bad sizes");
  IVX_REQUIRE(n_knots <= STRIDED_MAXK, IVX_ERR_UNSUPPORTED, "ivx_spline_strided_batched: at most %d knots per system (got %d)", STRIDED_MAXK, n_knots);
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_strided_batched: extrap must be 0, 1 or 2");
  IVX_REQUIRE(x_batch_stride >= 0 && x_knot_stride >= 1, IVX_ERR_INVALID_ARG, "ivx_spline_strided_batched: bad abscissa strides");
  if (n_batches == 0 || n_systems == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(x && y && xq && yq, IVX_ERR_INVALID_ARG, "ivx_spline_strided_batched: null pointer");
  const int64_t gx = (n_systems + 127) / 128;
  for (int64_t b0 = 0; b0 < n_batches; b0 += 65535) {   // gridDim.y limit
    const int64_t nb = n_batches - b0 < 65535 ? n_batches - b0 : 65535;
    dim3 grid((unsigned)gx, (unsigned)nb);
    k_spline_strided<<<grid, 128, 0, stream>>>(n_systems, n_knots, x + b0 * x_batch_stride, x_batch_stride, x_knot_stride,
                                               y + b0 * n_knots * n_systems, n_q, xq, extrap, yq + b0 * n_q * n_systems);
    IVX_LAUNCH_CHECK("k_spline_strided(batched)");
  }
  return IVX_OK;
}
