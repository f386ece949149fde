// ivx_spline.cu — kernel D: batched natural cubic spline fit + grid evaluation.
//
// North_star kernel 2 (the reference has no spline code): per (timestamp, expiry)
// slice, fit the natural cubic spline through the quoted (strike, IV) knots and evaluate it on a
// shared strike grid; then the same across maturities for every grid strike.
//
// ivx_spline_fit_eval — a GROUP of G threads per slice (G = 32 ... 1024 chosen from max_knots; 256 for
//   the 2000-strike chains), 8 consecutive rows per thread held in REGISTERS; the slice is staged once
//   into (padded, conflict-free) shared memory with coalesced loads, which is also where non-finite
//   knots are squeezed out.  The second derivatives M solve the tridiagonal system
//       h[i-1] M[i-1] + 2(h[i-1]+h[i]) M[i] + h[i] M[i+1] = 6 (dy[i]/h[i] - dy[i-1]/h[i-1]),  M[0] = M[n-1] = 0.
//   Thomas elimination is a chain of 2n dependent divisions; written in homogeneous form it is LINEAR:
//       q[i] = b[i] q[i-1] - h[i-1]^2 q[i-2]   (q = leading principal minors, c'[i] = h[i] q[i-1]/q[i])
//       D[i] = r[i] q[i-1] - h[i-1] D[i-1]     (d'[i] = D[i]/q[i])
//   so the state (q[i], q[i-1], D[i]) advances by a 3x3 matrix per row.  Each thread multiplies out its
//   8 rows (no division), the group runs a parallel prefix over those matrices (warp shuffles, then
//   one shared-memory hop between warps), and every thread re-walks its rows from the exact incoming
//   state — the divisions that are left (one per row, for 1/q[i]) are independent of each other.
//   Back-substitution M[i] = d'[i] - c'[i] M[i+1] is an affine map per row: same pattern, suffix scan
//   of 2-vectors.  The whole solve is exact (no truncated halo), O(n) work, depth 8 + log2(G); prefix
//   products are renormalised by powers of two (the recurrences are homogeneous), so no range limit.
//   Each thread then evaluates the grid points that fall inside its own 8 segments (the grid is sorted:
//   two binary searches give its range) straight from registers.
//   HBM traffic: 16 B per knot read once, 8 B per grid point written once.
// ivx_spline_strided — one THREAD per system for the many tiny systems of the maturity direction,
//   data laid out [knot][system] so every load/store is coalesced; when no knot is missing the LU
//   factors depend only on the shared abscissae and are computed once per CTA in shared memory.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr int SPL_RMAX = 8;       // rows per thread
constexpr int STRIDED_MAXK = 64;
constexpr int STRIDED_SPT = 4;      // systems per thread in the strided kernel
constexpr int STRIDED_MAXQ = 256;   // grid points whose segment, offset and class are tabulated once per CTA

// 1/a for a normal a (1e-300 < |a| < 1e300): 20-bit hardware seed and two Newton steps, without the special-case tail
// of the compiler's division (a zero or denormal a gives inf/NaN, like the division it replaces would give inf)
__device__ __forceinline__ double rcp_fast(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ double spline_eval(double xl, double xr, double yl, double yr, double Ml, double Mr, double xq) {
  const double h = xr - xl;
  const double A = xr - xq, B = xq - xl;
  const double ih = rcp_fast(h);
  return (Ml * A * A * A + Mr * B * B * B) * (ih * (1.0 / 6.0)) + (yl * ih - Ml * h * (1.0 / 6.0)) * A +
         (yr * ih - Mr * h * (1.0 / 6.0)) * B;
}

__device__ __forceinline__ double apply_extrap(double v, double xq, double xlo, double xhi, double ylo, double yhi, int extrap) {
  if (isnan(xq)) return qnan();
  if (xq < xlo) return extrap == 0 ? qnan() : (extrap == 2 ? ylo : v);
  if (xq > xhi) return extrap == 0 ? qnan() : (extrap == 2 ? yhi : v);
  return v;
}

// state transition of the homogeneous elimination: [[A,B,0],[C,E,0],[F,G,H]]
struct Mat3 {
  double A, B, C, E, F, G, H;
};
__device__ __forceinline__ Mat3 mat_identity() { return Mat3{1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0}; }
// L * R (R is applied first)
__device__ __forceinline__ Mat3 mat_mul(const Mat3& L, const Mat3& R) {
  Mat3 o;
  o.A = L.A * R.A + L.B * R.C;
  o.B = L.A * R.B + L.B * R.E;
  o.C = L.C * R.A + L.E * R.C;
  o.E = L.C * R.B + L.E * R.E;
  o.F = L.F * R.A + L.G * R.C + L.H * R.F;
  o.G = L.F * R.B + L.G * R.E + L.H * R.G;
  o.H = L.H * R.H;
  return o;
}
// 2^-e for the largest of the magnitudes whose high words (sign stripped) have maximum hw: 1.f x 2^e  (1.0 for zero,
// subnormal, inf, NaN).  Magnitudes compare like their high words, so the maximum is taken on integers.
__device__ __forceinline__ double pow2_inv_hw(int hw) {
  const int e = hw >> 20;
  return (e == 0 || e == 0x7ff) ? 1.0 : __hiloint2double((2046 - e) << 20, 0);
}
__device__ __forceinline__ int abs_hw(double v) { return __double2hiint(v) & 0x7fffffff; }
__device__ __forceinline__ double pow2_inv3(double a, double b, double c) {
  return pow2_inv_hw(max(max(abs_hw(a), abs_hw(b)), abs_hw(c)));
}
__device__ __forceinline__ void mat_normalize(Mat3& P) {
  const int hw = max(max(max(abs_hw(P.A), abs_hw(P.B)), max(abs_hw(P.C), abs_hw(P.E))),
                     max(max(abs_hw(P.F), abs_hw(P.G)), abs_hw(P.H)));
  const double sc = pow2_inv_hw(hw);
  P.A *= sc; P.B *= sc; P.C *= sc; P.E *= sc; P.F *= sc; P.G *= sc; P.H *= sc;
}
__device__ __forceinline__ Mat3 mat_shfl_up(const Mat3& P, int d) {
  Mat3 o;
  o.A = __shfl_up_sync(IVX_FULL_MASK, P.A, d); o.B = __shfl_up_sync(IVX_FULL_MASK, P.B, d);
  o.C = __shfl_up_sync(IVX_FULL_MASK, P.C, d); o.E = __shfl_up_sync(IVX_FULL_MASK, P.E, d);
  o.F = __shfl_up_sync(IVX_FULL_MASK, P.F, d); o.G = __shfl_up_sync(IVX_FULL_MASK, P.G, d);
  o.H = __shfl_up_sync(IVX_FULL_MASK, P.H, d);
  return o;
}

// one pad double per R rows: thread g reads doubles (R+1) g + j, and R+1 is odd -> conflict-free chunk reads
template <int R>
__device__ __forceinline__ int spl_pad_t(int i) { return i + i / R; }

template <int G>
__device__ __forceinline__ void group_sync(int grp) {
  if (G == 32) __syncwarp();
  else if (G >= 256) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(G) : "memory");
}

// first q in [0, n_q) with xq[q] >= v (n_q if none)
__device__ __forceinline__ int lower_bound_q(const double* __restrict__ xq, int n_q, double v) {
  int lo = 0, hi = n_q;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(xq + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// the same with a starting guess: on a (near-)uniform grid the answer is within one of (v - xq[0]) * scale, so four
// independent loads around the guess settle it in one L1 round trip; anything else falls back to the bisection
__device__ __forceinline__ int lower_bound_q_guess(const double* __restrict__ xq, int n_q, double v, double xq0, double scale) {
  const double t = (v - xq0) * scale;
  const int gi = (t > 0.0) ? ((t < (double)n_q) ? (int)t : n_q) : 0;      // NaN -> 0
  const int base = gi - 1;
  double wv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = base + k;
    const double ld = __ldg(xq + min(max(i, 0), n_q - 1));
    wv[k] = i < 0 ? -INFINITY : (i >= n_q ? INFINITY : ld);
  }
  if ((wv[0] < v) && (wv[3] >= v)) return base + 1 + (wv[1] < v ? 1 : 0) + (wv[2] < v ? 1 : 0);
  return lower_bound_q(xq, n_q, v);
}

template <int G, int SPL_R>
struct SplLayout {
  static constexpr int BT = G < 256 ? 256 : G;
  static constexpr int SPB = BT / G;
  static constexpr int NW = G / 32;
  static constexpr int CAP = G * SPL_R;
  static constexpr int PADCAP = CAP + CAP / SPL_R + 8;                 // rows CAP, CAP+1 exist (read as neighbours)
  static constexpr int PER_SLICE = 2 * PADCAP + NW * 7 + NW * 2 + NW + (G + 2) / 2;   // doubles (the tail holds NW + G+1 ints)
  static constexpr size_t SMEM = (size_t)SPB * PER_SLICE * sizeof(double);
};

// LADDER: the slices' abscissae come from shared ladders whose Thomas factors c'[i], 1/den[i], 1/h[i] were formed once
// (k_spline_ladder_factor) and sit in lu_cp/lu_inv/lu_ih, parallel to x.  A slice with every knot usable then only
// has its right-hand side to eliminate: d'[i] = (r[i] - h[i-1] d'[i-1]) / den[i] is an affine map per row — a prefix scan
// of 2-vectors, no matrices, no division.  Slices with missing knots take the general path below.
template <int G, int SPL_R, int MODE>   // MODE 0: general; 1: slices sit on tabulated ladders; 2: presumed to — verified per slice
__device__ __forceinline__ void spline_slice(const int64_t sl, const int64_t* __restrict__ knot_off,
                                             const int64_t* __restrict__ x_off,
                                             const double* __restrict__ x,
                                             const double* __restrict__ lu_cp, const double* __restrict__ lu_inv,
                                             const double* __restrict__ lu_ih,
                                             const int32_t* __restrict__ slice_ladder,
                                             const double* __restrict__ x_own,
                                             const int64_t* __restrict__ ladder_off,
                                             const double* __restrict__ y,
                                             int n_q, const double* __restrict__ xq, int extrap,
                                             double* __restrict__ yq) {
  using L = SplLayout<G, SPL_R>;
  constexpr bool LADDER = MODE != 0;
  auto spl_pad = [](int i) { return spl_pad_t<SPL_R>(i); };
  extern __shared__ __align__(16) double smem[];
  const int grp = threadIdx.x / G, g = threadIdx.x % G;
  const int lane = threadIdx.x & 31, w = g >> 5;
  double* xs = smem + (size_t)grp * L::PER_SLICE;
  double* ys = xs + L::PADCAP;
  double* wm = ys + L::PADCAP;          // [NW][7] warp totals of the forward scan
  double* wa = wm + L::NW * 7;          // [NW][2] warp totals of the backward scan
  int* wc = reinterpret_cast<int*>(wa + L::NW * 2);   // [NW] valid-knot counts
  int* qs = reinterpret_cast<int*>(wa + L::NW * 3);   // [G+1] first grid point of every thread
  const int64_t k0 = knot_off[sl];
  const int64_t xk0 = x_off ? x_off[sl] : k0;      // abscissae may live elsewhere (a strike ladder shared by many slices)
  int m = (int)min((int64_t)L::CAP + 1, knot_off[sl + 1] - k0);
  if (m > L::CAP) m = 0;                // more knots than the launch was sized for: NaN row (host checks max_knots)
  double* out = yq + sl * (int64_t)n_q;

  // ---- stage the slice (coalesced), NaN beyond its end ------------------------------------------------
  // All of a thread's loads are issued before the first one is used (clamped addresses, no branch around them): one
  // round trip to HBM per slice instead of one per row of the loop.
  // VERIFY mode (LADDER with x_own): the slice brings its own abscissae and is only PRESUMED to sit on ladder
  // slice_ladder[sl]; its x is compared with the ladder's (L2-resident) bit for bit, and the tabulated factors are used
  // only when every knot matches — a caller that does not know whether its slices share ladders loses nothing.
  constexpr bool verify = MODE == 2;
  bool fin = true;
  {
    double xl[SPL_R + 1], yl[SPL_R + 1];
#pragma unroll
    for (int r = 0; r < SPL_R + 1; ++r) { xl[r] = qnan(); yl[r] = qnan(); }
    if (m > 0) {                           // group-uniform; an empty slice may sit at the very end of the arrays
#pragma unroll
      for (int r = 0; r < SPL_R + 1; ++r) {
        const int jc = min(r * G + g, m - 1);
        xl[r] = verify ? ld_stream(x_own + k0 + jc) : (x_off ? __ldg(x + xk0 + jc) : ld_stream(x + xk0 + jc));
        yl[r] = ld_stream(y + k0 + jc);
      }
    }
#pragma unroll
    for (int r = 0; r < SPL_R + 1; ++r) {
      const int j = r * G + g;
      if (j < L::CAP + 2) {
        const bool in = j < m;
        const double xv = in ? xl[r] : qnan(), yv = in ? yl[r] : qnan();
        fin = fin && (!in || ((fabs(xv) < INFINITY) && (fabs(yv) < INFINITY)));
        xs[spl_pad(j)] = xv; ys[spl_pad(j)] = yv;
      }
    }
  }
  // VERIFY: the ladder's abscissae are requested before the barrier (the staging registers are free by now), so their L2
  // round trip overlaps with the wait for the other warps' staging
  double xd[SPL_R + 1];
  int64_t llen = 0;
  if (verify && G >= 256 && m > 0) {                       // group-uniform
    const int lad = slice_ladder[sl];
    llen = ladder_off[lad + 1] - ladder_off[lad];
#pragma unroll
    for (int r = 0; r < SPL_R + 1; ++r) xd[r] = __ldg(x + xk0 + max((int64_t)0, min((int64_t)min(r * G + g, m - 1), llen - 1)));
  }
  // a slice with every knot usable (the common case) is known at the barrier itself when the group is the whole CTA
  bool clean = false;
  if (G >= 256) clean = __syncthreads_and(fin) != 0;
  else group_sync<G>(grp);
  bool on_ladder = LADDER && !verify;                      // told so by the caller ...
  if (verify && G >= 256 && clean) {                       // ... or found out (group-uniform branch)
    bool same = (int64_t)m <= llen;
#pragma unroll
    for (int r = 0; r < SPL_R + 1; ++r) {
      const int j = r * G + g;
      if (j < m) same = same && (xd[r] == xs[spl_pad(j)]);
    }
    on_ladder = __syncthreads_and(same) != 0;
  }

  // ---- squeeze out unusable knots if there are any (chunks of SPL_R knots per thread, stable) ------------------------
  const int i0 = g * SPL_R;
  int n = m;
  if (!clean) {                          // group-uniform
    double xv[SPL_R], yv[SPL_R];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < SPL_R; ++j) {
      xv[j] = xs[spl_pad(i0 + j)]; yv[j] = ys[spl_pad(i0 + j)];
      cnt += (isfinite(xv[j]) && isfinite(yv[j])) ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(IVX_FULL_MASK, incl, d);
      if (lane >= d) incl += u;
    }
    int before = incl - cnt;
    n = __shfl_sync(IVX_FULL_MASK, incl, 31);
    if (L::NW > 1) {
      if (lane == 31) wc[w] = incl;
      group_sync<G>(grp);
      n = 0;
#pragma unroll
      for (int ww = 0; ww < L::NW; ++ww) {
        const int c = wc[ww];
        if (ww < w) before += c;
        n += c;
      }
    }
    if (n != m) {                          // group-uniform
      group_sync<G>(grp);                  // every thread holds its rows in registers by now
      int o = before;
#pragma unroll
      for (int j = 0; j < SPL_R; ++j) {
        if (isfinite(xv[j]) && isfinite(yv[j])) { xs[spl_pad(o)] = xv[j]; ys[spl_pad(o)] = yv[j]; ++o; }
      }
#pragma unroll
      for (int r = 0; r < SPL_R; ++r) {
        const int j = r * G + g;
        if (j >= n && j < m) { xs[spl_pad(j)] = qnan(); ys[spl_pad(j)] = qnan(); }
      }
      group_sync<G>(grp);
    }
  }
  if (n < 2) {
    for (int q = g; q < n_q; q += G) st_stream(out + q, qnan());
    return;
  }

  // Thread g owns the equations (rows) i = R g + 1 + j, j = 0..R-1 — row 0 is not an equation, so every thread's rows
  // are equations up to the end of the slice — and the segments [x[i], x[i+1]] with the same numbers; thread 0 also
  // owns segment 0.  Rows past n-2 are computed like the others on whatever the NaN padding gives and zeroed afterwards:
  // the elimination only flows forward, so they cannot reach a valid row, and no per-row test is left in the loops.
  double cv[SPL_R], dv[SPL_R];
  const int nvalid = min(max(n - 2 - i0, 0), SPL_R);      // rows i0+1 .. i0+nvalid are equations
  if (LADDER && n == m && on_ladder) {
    // ---- shared-ladder path: slopes through the tabulated 1/h, right-hand side as affine maps of d' ------------------
    // tables are stored the way they are read: row i = R g + 1 + j of ladder l sits at l CAP + j G + g, so that the 32
    // lanes of a warp read 32 consecutive doubles; ih at row i is 1 / (x[i] - x[i-1])
    const int64_t tb = (int64_t)slice_ladder[sl] * L::CAP;
    const double* __restrict__ cp_t = lu_cp + tb;
    const double* __restrict__ inv_t = lu_inv + tb;
    const double* __restrict__ ih_t = lu_ih + tb;
    // table positions are always inside the table (j G + g < CAP): the loads are unconditional, issued together, and
    // entries past the end of the ladder (never written) are discarded by the selects below
    double ihv[SPL_R + 1], invv[SPL_R];
#pragma unroll
    for (int k = 0; k <= SPL_R; ++k) ihv[k] = __ldg(ih_t + (k < SPL_R ? k * G + g : (g + 1 < G ? g + 1 : g)));
#pragma unroll
    for (int j = 0; j < SPL_R; ++j) invv[j] = __ldg(inv_t + j * G + g);
    double slp[SPL_R + 1];                // slp[k] = slope of segment i0+k
    {
      double yl = ys[spl_pad(i0)];
#pragma unroll
      for (int k = 0; k <= SPL_R; ++k) {
        const double yr = ys[spl_pad(i0 + k + 1)];
        slp[k] = (i0 + k <= n - 2) ? (yr - yl) * ihv[k] : 0.0;
        yl = yr;
      }
    }
    double al[SPL_R], bt[SPL_R];
    double A = 0.0, B = 1.0;              // d'[last row of the chunk] = A + B d'[last row of the previous chunk]
    {
      double xl = xs[spl_pad(i0)];
#pragma unroll
      for (int j = 0; j < SPL_R; ++j) {
        const double xr = xs[spl_pad(i0 + j + 1)];
        const bool rowok = j < nvalid;
        al[j] = rowok ? 6.0 * (slp[j + 1] - slp[j]) * invv[j] : 0.0;
        bt[j] = rowok ? -(xr - xl) * invv[j] : 0.0;
        if (rowok) { A = fma(bt[j], A, al[j]); B *= bt[j]; }
        xl = xr;
      }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double A2 = __shfl_up_sync(IVX_FULL_MASK, A, d), B2 = __shfl_up_sync(IVX_FULL_MASK, B, d);
      if (lane >= d) { A = fma(B, A2, A); B *= B2; }
    }
    double XA = __shfl_up_sync(IVX_FULL_MASK, A, 1), XB = __shfl_up_sync(IVX_FULL_MASK, B, 1);
    if (lane == 0) { XA = 0.0; XB = 1.0; }
    double d = 0.0;                       // d' before row 1
    if (L::NW > 1) {
      if (lane == 31) { wa[w * 2] = A; wa[w * 2 + 1] = B; }
      group_sync<G>(grp);
      for (int ww = 0; ww < w; ++ww) d = fma(wa[ww * 2 + 1], d, wa[ww * 2]);
      group_sync<G>(grp);                 // wa is reused by the backward scan
    }
    d = fma(XB, d, XA);
#pragma unroll
    for (int j = 0; j < SPL_R; ++j) cv[j] = __ldg(cp_t + j * G + g);
#pragma unroll
    for (int j = 0; j < SPL_R; ++j) {
      const bool rowok = j < nvalid;
      d = fma(bt[j], d, al[j]);            // rows past the end: bt = al = 0
      dv[j] = rowok ? d : 0.0;
      cv[j] = rowok ? cv[j] : 0.0;
    }
  } else {
    // ---- row coefficients.  Row i (1 <= i <= n-2) is multiplied by h[i-1] h[i], which clears both slope
    //      divisions:  a = hm^2 hp,  b = 2 (hm+hp) hm hp,  c = hm hp^2,  r = 6 (dy[i] hm - dy[i-1] hp), hm = h[i-1], hp = h[i];
    //      the q recurrence needs a[i] c[i-1], and c is carried from row to row.
    // The differences are formed from shared memory twice (before the scan and after it) instead of being kept in
    // registers across the scan: the live set during the scan is then the two matrices only, which is what lets four
    // CTAs share an SM.   hh[k] = x[i0+k] - x[i0+k-1], k = 0..R+1;  dy[k] = y[i0+k+1] - y[i0+k], k = 0..R
    auto load_diffs = [&](double (&hh)[SPL_R + 2], double (&dy)[SPL_R + 1]) {
      double xc = g > 0 ? xs[spl_pad(i0 - 1)] : qnan();
      double yc = qnan();
#pragma unroll
      for (int k = 0; k <= SPL_R + 1; ++k) {
        const double xk = xs[spl_pad(i0 + k)], yk = ys[spl_pad(i0 + k)];
        hh[k] = xk - xc;
        if (k > 0) dy[k - 1] = yk - yc;
        xc = xk; yc = yk;
      }
    };
    // row j (i = i0+1+j): hm = hh[j+1], hp = hh[j+2], dy[i-1] = dy[j], dy[i] = dy[j+1]; c of row i0 (before j = 0) = hh[0] hh[1]^2

    // ---- forward: chunk transition matrix, prefix over the group ----------------------------------------------
    Mat3 P = mat_identity();
    {
      double hh[SPL_R + 2], dy[SPL_R + 1];
      load_diffs(hh, dy);
      double c_prev = g > 0 ? hh[0] * hh[1] * hh[1] : 0.0;           // multiplies q[-1] = 0 in row 1
#pragma unroll
      for (int j = 0; j < SPL_R; ++j) {
        const double hm = hh[j + 1], hp = hh[j + 2], hmp = hm * hp;
        const double a = hm * hmp, b = (hm + hp) * (hmp + hmp), r = 6.0 * (dy[j + 1] * hm - dy[j] * hp);
        const double ag = a * c_prev;
        c_prev = hmp * hp;
        const double nA = b * P.A - ag * P.C, nB = b * P.B - ag * P.E;
        const double nF = r * P.A - a * P.F, nG = r * P.B - a * P.G;
        P.C = P.A; P.E = P.B; P.A = nA; P.B = nB; P.F = nF; P.G = nG; P.H = -a * P.H;
      }
    }
    mat_normalize(P);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {       // products of normalised matrices grow by at most 3x per level: no rescaling inside
      const Mat3 Q = mat_shfl_up(P, d);
      if (lane >= d) P = mat_mul(P, Q);
    }
    mat_normalize(P);
    Mat3 X = mat_shfl_up(P, 1);            // exclusive prefix inside the warp
    if (lane == 0) X = mat_identity();
    double q = 1.0, qp = 0.0, D = 0.0;     // state before row 1: q[0] = 1, q[-1] = 0, D[0] = 0
    if (L::NW > 1) {
      if (lane == 31) {
        double* t = wm + w * 7;
        t[0] = P.A; t[1] = P.B; t[2] = P.C; t[3] = P.E; t[4] = P.F; t[5] = P.G; t[6] = P.H;
      }
      group_sync<G>(grp);
      for (int ww = 0; ww < w; ++ww) {     // normalised warp totals (largest entry in [1,2)) move the state by at most 6x each: no rescaling inside
        const double* t = wm + ww * 7;
        const double nq = t[0] * q + t[1] * qp, nqp = t[2] * q + t[3] * qp, nD = t[4] * q + t[5] * qp + t[6] * D;
        q = nq; qp = nqp; D = nD;
      }
    }
    {
      const double nq = X.A * q + X.B * qp, nqp = X.C * q + X.E * qp, nD = X.F * q + X.G * qp + X.H * D;
      const double sc = pow2_inv3(nq, nqp, nD);
      q = nq * sc; qp = nqp * sc; D = nD * sc;
    }
    // ---- re-walk the rows from the exact incoming state: Thomas' c', d' ---------------------------------------
    {
      double hh[SPL_R + 2], dy[SPL_R + 1];
      load_diffs(hh, dy);
      double c_prev = g > 0 ? hh[0] * hh[1] * hh[1] : 0.0;
#pragma unroll
      for (int j = 0; j < SPL_R; ++j) {
        const double hm = hh[j + 1], hp = hh[j + 2], hmp = hm * hp;
        const double a = hm * hmp, b = (hm + hp) * (hmp + hmp), r = 6.0 * (dy[j + 1] * hm - dy[j] * hp);
        const double ag = a * c_prev;
        c_prev = hmp * hp;
        const double qn = b * q - ag * qp, Dn = r * q - a * D;
        const double inv = rcp_fast(qn);     // the state was renormalised on entry and grows by h^3 per row: normal range
        cv[j] = c_prev * q * inv;
        dv[j] = Dn * inv;
        qp = q; q = qn; D = Dn;
      }
    }
    if (nvalid < SPL_R) {                  // the thread that holds the end of the slice, and those after it
#pragma unroll
      for (int j = 0; j < SPL_R; ++j)
        if (j >= nvalid) { cv[j] = 0.0; dv[j] = 0.0; }
    }
  }
  // ---- backward: M[i] = d'[i] - c'[i] M[i+1] as affine maps, suffix scan ---------------------------------------
  double al = 0.0, be = 1.0;             // M[first row of the chunk] = al + be * M[first row of the next chunk]
#pragma unroll
  for (int j = SPL_R - 1; j >= 0; --j) { al = dv[j] - cv[j] * al; be = -cv[j] * be; }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double al2 = __shfl_down_sync(IVX_FULL_MASK, al, d), be2 = __shfl_down_sync(IVX_FULL_MASK, be, d);
    if (lane + d < 32) { al = al + be * al2; be = be * be2; }
  }
  double z = 0.0;                        // M at the first row of the next warp
  if (L::NW > 1) {
    if (lane == 0) { wa[w * 2] = al; wa[w * 2 + 1] = be; }
    group_sync<G>(grp);
    for (int ww = L::NW - 1; ww > w; --ww) z = wa[ww * 2] + wa[ww * 2 + 1] * z;
  }
  const double m_first = al + be * z;
  double Mn = __shfl_down_sync(IVX_FULL_MASK, m_first, 1);
  if (lane == 31) Mn = z;
#pragma unroll
  for (int j = SPL_R - 1; j >= 0; --j) { dv[j] = dv[j] - cv[j] * (j == SPL_R - 1 ? Mn : dv[j + 1]); }   // dv[j] = M[i0+1+j] now

  // ---- evaluate the grid points inside my segments ------------------------------------------------------------
  // the grid is sorted: thread g's points start at the first one >= x[i0+1] (thread 0: at the first point) and end
  // where thread g+1's start; the owner of the last segment (n-2) also takes everything above it
  const bool has_seg = i0 + 1 <= n - 2;
  {
    const double xq0 = __ldg(xq), xqN = __ldg(xq + n_q - 1);
    const double scale = (double)(n_q - 1) * rcp_fast(xqN - xq0);    // inf/NaN on a degenerate grid: the guess fails, bisection runs
    qs[g] = (g == 0) ? 0 : (has_seg ? lower_bound_q_guess(xq, n_q, xs[spl_pad(i0 + 1)], xq0, scale) : n_q);
  }
  if (g == 0) qs[G] = n_q;
  group_sync<G>(grp);
  if (!has_seg && g > 0) return;
  const double xlo = xs[spl_pad(0)], xhi = xs[spl_pad(n - 1)], ylo = ys[spl_pad(0)], yhi = ys[spl_pad(n - 1)];
  const bool owns_last = (has_seg && n - 2 <= i0 + SPL_R) || !has_seg;   // !has_seg here: thread 0 of a 2- or 3-knot... (n <= 2)
  const bool edge = g == 0 || owns_last;
  int left = (owns_last ? n_q : qs[g + 1]) - qs[g];
  // the current grid point rides in a register (+inf once the thread's range is exhausted); a segment takes points while
  // they lie left of its right knot — the last segment's right knot counts as +inf, so it takes whatever is left
  const double* __restrict__ xqp = xq + qs[g];
  double* __restrict__ op = out + qs[g];
  double xq_cur = left > 0 ? __ldg(xqp) : INFINITY;
  auto take = [&](double xl_, double xr_, double yl_, double yr_, double Ml_, double Mr_, bool last) {
    const double xr_take = last ? INFINITY : xr_;
    while (!(xq_cur >= xr_take)) {                                  // a NaN grid point is taken where it stands and yields NaN
      // most segments receive no grid point (8 segments, ~2 points per thread): the opaque move keeps the segment's
      // reciprocal and products inside the loop instead of being hoisted in front of every segment's (empty) loop
      double xl_in = xl_;
      asm volatile("" : "+d"(xl_in));
      const double v = spline_eval(xl_in, xr_, yl_, yr_, Ml_, Mr_, xq_cur);
      // only the first thread can see points below the first knot, only the last segment's owner points above the last
      st_stream(op, edge ? apply_extrap(v, xq_cur, xlo, xhi, ylo, yhi, extrap) : v);
      ++op; ++xqp; --left;
      xq_cur = left > 0 ? __ldg(xqp) : INFINITY;
    }
  };
  double xl_ = xs[spl_pad(i0 + 1)], yl_ = ys[spl_pad(i0 + 1)];
  if (g == 0) take(xlo, xl_, ylo, yl_, 0.0, n > 2 ? dv[0] : 0.0, n <= 2);   // segment 0 (M[0] = 0; M[1] = 0 too when n == 2)
#pragma unroll
  for (int j = 0; j < SPL_R; ++j) {
    const double xr_ = xs[spl_pad(i0 + j + 2)], yr_ = ys[spl_pad(i0 + j + 2)];
    const double Mr_ = j < SPL_R - 1 ? dv[j + 1 < SPL_R ? j + 1 : j] : Mn;
    take(xl_, xr_, yl_, yr_, dv[j], Mr_, i0 + 1 + j >= n - 2);        // segments past n-2 are never reached with points left
    xl_ = xr_; yl_ = yr_;
  }
}

#define IVX_SPL_PARAMS const int64_t* __restrict__ knot_off, const int64_t* __restrict__ x_off, const double* __restrict__ x,                      \
                       const double* __restrict__ lu_cp, const double* __restrict__ lu_inv, const double* __restrict__ lu_ih,                 \
                       const int32_t* __restrict__ slice_ladder, const double* __restrict__ x_own, const int64_t* __restrict__ ladder_off,    \
                       const double* __restrict__ y, int n_q, const double* __restrict__ xq, int extrap, double* __restrict__ yq
#define IVX_SPL_ARGS knot_off, x_off, x, lu_cp, lu_inv, lu_ih, slice_ladder, x_own, ladder_off, y, n_q, xq, extrap, yq

// one group of G threads per slice, slices in order
template <int G, int SPL_R, int MODE>
__global__ void __launch_bounds__(SplLayout<G, SPL_R>::BT, (G <= 256 ? 4 : 1)) k_spline_fit_eval(int64_t n_slices, IVX_SPL_PARAMS) {
  const int64_t sl = (int64_t)blockIdx.x * SplLayout<G, SPL_R>::SPB + threadIdx.x / G;
  if (sl >= n_slices) return;           // a whole group leaves together (named barriers / warp sync inside)
  spline_slice<G, SPL_R, MODE>(sl, IVX_SPL_ARGS);
}

// the slices flagged in need[0 .. n_slices): what the banded kernel (ivx_spline_band.cu) could not take — a missing knot, a
// slice off its ladder, a ladder whose band failed its exactness check.  Fixed grid; the groups walk the slices IN ORDER and
// skip the ones that are done, so that the CTAs running at any moment work on neighbouring slices (an atomically appended
// list visits them in a scattered order: every slice in another DRAM page and TLB entry — measured 1.45x slower).
template <int G, int SPL_R, int MODE>
__global__ void __launch_bounds__(SplLayout<G, SPL_R>::BT, (G <= 256 ? 4 : 1)) k_spline_fit_eval_list(int64_t n_slices, const int32_t* __restrict__ need,
                                                                                                  const uint32_t* __restrict__ n_need, IVX_SPL_PARAMS) {
  if (n_need && *n_need == 0) return;               // nothing was flagged (the usual day): do not walk 700 000 flags to find that out
  constexpr int SPB = SplLayout<G, SPL_R>::SPB;
  constexpr int BT = SplLayout<G, SPL_R>::BT;
  __shared__ int32_t s_need[BT];
  const int grp = threadIdx.x / G;
  // Group j of CTA b takes the slices (b SPB + j) + k (gridDim SPB), k = 0, 1, ...: the CTAs running at any moment work on
  // neighbouring slices (DRAM and TLB locality) and every CTA gets the same share.  The flags of BT rounds are fetched at once,
  // one per thread (a flag per iteration would cost an L2 round trip per slice: 0.6 ms per day for nothing).
  const int64_t stride = (int64_t)gridDim.x * SPB;
  for (int64_t k0 = 0; (int64_t)blockIdx.x * SPB + k0 * stride < n_slices; k0 += BT / SPB) {
    __syncthreads();
    {
      const int64_t sl = (int64_t)blockIdx.x * SPB + (threadIdx.x % SPB) + (k0 + threadIdx.x / SPB) * stride;    // entry [round][group]
      s_need[threadIdx.x] = sl < n_slices ? need[sl] : 0;
    }
    __syncthreads();
    for (int r = 0; r < BT / SPB; ++r) {
      if (!s_need[r * SPB + grp]) continue;         // group-uniform
      spline_slice<G, SPL_R, MODE>((int64_t)blockIdx.x * SPB + grp + (k0 + r) * stride, IVX_SPL_ARGS);
      group_sync<G>(grp);               // the group's shared memory is rewritten by its next slice
    }
  }
}

// ---- many small systems, [knot][system] layout ----------------------------------------------------------
// generic path of the strided kernel: some knots of this system are unusable -> private abscissae (rare)
__device__ __noinline__ void strided_generic(int64_t n_sys, int64_t sys, int nk, const double* s_x, const double* __restrict__ y,
                                             int n_q, const double* __restrict__ xq, int extrap, double* __restrict__ yq) {
  double xk[STRIDED_MAXK], yk[STRIDED_MAXK], ck[STRIDED_MAXK], Mk[STRIDED_MAXK];
  int n = 0;
  for (int i = 0; i < nk; ++i) {
    const double xv = s_x[i], yv = __ldg(y + (int64_t)i * n_sys + sys);
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

// blockIdx.y = batch: batch b has its own abscissae x[b*x_batch_stride + i*x_knot_stride] and its own
// [knot][system] / [query][system] planes of y / yq.
// Fast path (every knot usable): the LU factors depend only on the shared abscissae and are formed once per CTA;
// a thread streams its system's y forward (coalesced loads), parks Thomas' d' in shared memory [knot][thread]
// (conflict-free), and evaluates the grid points of each segment on the way back down — no per-thread arrays,
// no division.  A non-finite y sends the thread to the generic path.
__global__ void __launch_bounds__(128) k_spline_strided(int64_t n_sys, int nk, const double* __restrict__ x,
                                                        int64_t x_batch_stride, int64_t x_knot_stride,
                                                        const double* __restrict__ y, int n_q,
                                                        const double* __restrict__ xq, int extrap,
                                                        double* __restrict__ yq) {
  __shared__ double s_x[STRIDED_MAXK], s_cp[STRIDED_MAXK], s_inv[STRIDED_MAXK], s_ih[STRIDED_MAXK];
  __shared__ int s_x_ok;
  // Every system of the CTA shares the abscissae AND the grid, so which segment a grid point falls in, its offset from the
  // segment's left knot and whether it lies outside the knots are the same for all 512 systems: tabulated once (ascending
  // grid, as the walk below has always assumed).  The evaluation loop then has CTA-uniform bounds and no compare per point.
  __shared__ double s_B[STRIDED_MAXQ];          // xq - x[segment]
  __shared__ uint8_t s_cls[STRIDED_MAXQ];       // 0 inside, 1 below the first knot, 2 above the last, 3 NaN
  __shared__ int s_qf[STRIDED_MAXK + 1];        // segment i takes grid points s_qf[i] .. s_qf[i+1]-1
  const bool tab = n_q <= STRIDED_MAXQ;
  extern __shared__ __align__(16) double s_d[];          // [nk][128]
  x += (int64_t)blockIdx.y * x_batch_stride;
  y += (int64_t)blockIdx.y * nk * n_sys;
  yq += (int64_t)blockIdx.y * n_q * n_sys;
  if (threadIdx.x < STRIDED_MAXK && threadIdx.x < nk) s_x[threadIdx.x] = x[threadIdx.x * x_knot_stride];
  __syncthreads();
  if (threadIdx.x + 1 < nk && threadIdx.x < STRIDED_MAXK) s_ih[threadIdx.x] = 1.0 / (s_x[threadIdx.x + 1] - s_x[threadIdx.x]);
  if (threadIdx.x == 0) {
    bool okx = true;
    for (int i = 0; i < nk; ++i) okx = okx && isfinite(s_x[i]);
    s_x_ok = okx;
    double c_prev = 0.0;
    for (int i = 1; i + 1 < nk; ++i) {
      const double hm = s_x[i] - s_x[i - 1], hp = s_x[i + 1] - s_x[i];
      const double inv = 1.0 / (2.0 * (hm + hp) - hm * c_prev);
      c_prev = hp * inv;
      s_cp[i] = c_prev; s_inv[i] = inv;
    }
  }
  if (tab && nk >= 3) {
    // s_qf[i] = number of grid points below x[i] (i = 1 .. nk-2); segment 0 also takes what lies below x[0], segment nk-2 what lies above
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
      int lo = 0, hi = n_q;                               // first q with !(xq[q] < x[i]); NaN grid points count as "not below"
      const double xi = s_x[i];
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(xq + mid) < xi) lo = mid + 1; else hi = mid; }
      s_qf[i] = i == 0 ? 0 : (i == nk - 1 ? n_q : lo);
    }
  }
  __syncthreads();
  if (tab && nk >= 3) {
    for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
      const double xv = __ldg(xq + q);
      int lo = 0, hi = nk - 1;                            // segment: last i in [0, nk-2] with s_qf[i] <= q
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_qf[mid] <= q) lo = mid; else hi = mid; }
      s_B[q] = xv - s_x[lo];
      s_cls[q] = isnan(xv) ? 3 : (xv < s_x[0] ? 1 : (xv > s_x[nk - 1] ? 2 : 0));
    }
  }
  __syncthreads();
  double* sd = s_d + threadIdx.x;                         // sd[i * 128]
  constexpr int PF = 8;                                   // loads in flight per thread: y is fetched 8 knots ahead
  // STRIDED_SPT systems per thread, one after the other: the CTA's factorisation is shared by 128 * STRIDED_SPT systems
  for (int rep = 0; rep < STRIDED_SPT; ++rep) {
  const int64_t sys = ((int64_t)blockIdx.x * STRIDED_SPT + rep) * blockDim.x + threadIdx.x;
  if (sys >= n_sys) return;
  if (!s_x_ok || nk < 3) { strided_generic(n_sys, sys, nk, s_x, y, n_q, xq, extrap, yq); continue; }

  // ---- forward: d'[i] = (6 (slope[i] - slope[i-1]) - h[i-1] d'[i-1]) / den[i], y streamed ------------------------
  const double* yp = y + sys;
  double y_prev = __ldg(yp), y_cur = __ldg(yp + n_sys);
  const double y_first = y_prev;
  bool ok = isfinite(y_prev) && isfinite(y_cur);
  double sm = (y_cur - y_prev) * s_ih[0], d_prev = 0.0;
  for (int ib = 1; ib + 1 < nk; ib += PF) {
    double yb[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) yb[u] = (ib + u + 1 < nk) ? __ldg(yp + (int64_t)(ib + u + 1) * n_sys) : 0.0;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int i = ib + u;
      if (i + 1 < nk) {
        const double y_next = yb[u];
        ok = ok && isfinite(y_next);
        const double sp = (y_next - y_cur) * s_ih[i];
        d_prev = (6.0 * (sp - sm) - (s_x[i] - s_x[i - 1]) * d_prev) * s_inv[i];
        sd[i * 128] = d_prev;
        sm = sp; y_cur = y_next;
      }
    }
  }
  if (!ok) { strided_generic(n_sys, sys, nk, s_x, y, n_q, xq, extrap, yq); continue; }
  const double y_last = y_cur;
  // ---- backward: M[i] = d'[i] - c'[i] M[i+1]; the grid points of segment [x[i], x[i+1]) are evaluated as soon as its
  //      two moments are known (descending); segment 0 also takes the points below x[0], segment nk-2 those above ------
  const double xlo = s_x[0], xhi = s_x[nk - 1];
  int q = n_q - 1;
  double* outp = yq + (int64_t)q * n_sys + sys;
  double M_hi = 0.0, y_hi = y_last;
  for (int ib = nk - 2; ib >= 0; ib -= PF) {
    double yb[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) yb[u] = (ib - u >= 0) ? __ldg(yp + (int64_t)(ib - u) * n_sys) : 0.0;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int i = ib - u;
      if (i >= 0) {
        const double M_lo = i >= 1 ? sd[i * 128] - s_cp[i] * M_hi : 0.0;
        const double y_lo = yb[u];
        // the segment's cubic about its left knot: y = a0 + B (a1 + B (a2 + B a3)),  B = x - x[i]
        const double xl = s_x[i], h = s_x[i + 1] - xl, ih = s_ih[i];
        const double a1 = (y_hi - y_lo) * ih - h * (2.0 * M_lo + M_hi) * (1.0 / 6.0);
        const double a2 = 0.5 * M_lo, a3 = (M_hi - M_lo) * ih * (1.0 / 6.0);
        if (tab) {
          const int q_lo = s_qf[i], q_hi = s_qf[i + 1];    // CTA-uniform
          double* op = yq + (int64_t)(q_hi - 1) * n_sys + sys;
          for (int qq = q_hi - 1; qq >= q_lo; --qq, op -= n_sys) {
            const double B = s_B[qq];
            double v = fma(B, fma(B, fma(B, a3, a2), a1), y_lo);
            const int c = s_cls[qq];
            if (c) v = (c == 3 || extrap == 0) ? qnan() : (extrap == 2 ? (c == 1 ? y_first : y_last) : v);
            st_stream(op, v);
          }
        } else
        while (q >= 0) {
          const double xv = __ldg(xq + q);
          if (!(i == 0 || xv >= xl || isnan(xv))) break;
          const double B = xv - xl;
          const double v = fma(B, fma(B, fma(B, a3, a2), a1), y_lo);
          st_stream(outp, apply_extrap(v, xv, xlo, xhi, y_first, y_last, extrap));
          --q; outp -= n_sys;
        }
        M_hi = M_lo; y_hi = y_lo;
      }
    }
  }
  }  // rep
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_spline_workspace_bytes(int64_t, int64_t) { return 256; }

template <int G, int SPL_R, int MODE>
static int launch_spline_t(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, const double* x, const double* lu_cp,
                           const double* lu_inv, const double* lu_ih, const int32_t* slice_ladder, const double* x_own,
                           const int64_t* ladder_off, const double* y, int n_q,
                         const double* xq, int extrap, double* yq, cudaStream_t stream) {
  using L = SplLayout<G, SPL_R>;
  static DeviceOnce attr_once;                       // the attribute is per device, not per process
  if (once_per_device(attr_once)) IVX_CUDA(cudaFuncSetAttribute(k_spline_fit_eval<G, SPL_R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
  const int64_t blocks = (n_slices + L::SPB - 1) / L::SPB;
  IVX_REQUIRE(blocks < ((int64_t)1 << 31), IVX_ERR_TOO_LARGE, "ivx_spline_fit_eval: too many slices for one launch; split the batch");
  k_spline_fit_eval<G, SPL_R, MODE><<<(unsigned)blocks, L::BT, L::SMEM, stream>>>(n_slices, IVX_SPL_ARGS);
  IVX_LAUNCH_CHECK("k_spline_fit_eval");
  return IVX_OK;
}

// threads per slice for a given max_knots (the kernels are instantiated for these widths, 8 rows per thread)
static int spline_group_width(int64_t max_knots) {
  for (int g = 32; g < 1024; g <<= 1)
    if (max_knots <= (int64_t)g * SPL_RMAX) return g;
  return 1024;
}

// one CTA per ladder: the x-only half of Thomas' elimination, once for every slice that uses the ladder.  The abscissae
// are staged into shared memory by the whole CTA (the recurrence itself is a serial chain of divisions, run by thread 0
// out of shared memory), results leave through coalesced stores.
// Row i >= 1 of ladder l goes to l*cap + ((i-1) % 8)*G + (i-1)/8 — the order in which the evaluation kernel's threads
// read it (thread g owns rows 8g+1 .. 8g+8).
__global__ void __launch_bounds__(256) k_spline_ladder_factor(int64_t n_ladders, const int64_t* __restrict__ ladder_off,
                                                             const double* __restrict__ x, int G, double* __restrict__ lu_cp,
                                                             double* __restrict__ lu_inv, double* __restrict__ lu_ih) {
  extern __shared__ __align__(16) double sm[];      // x[n], cp[n], inv[n]
  const int64_t l = blockIdx.x;
  const int64_t a = ladder_off[l], cap = (int64_t)G * SPL_RMAX;
  int64_t n64 = ladder_off[l + 1] - a;
  if (n64 > cap) n64 = cap;
  const int n = (int)n64;
  double* sx = sm;
  double* scp = sm + cap;
  double* sinv = sm + 2 * cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = x[a + i];
  __syncthreads();
  if (threadIdx.x == 0) {
    double c_prev = 0.0;
    for (int i = 0; i < n; ++i) {
      double cp = 0.0, inv = 0.0;
      if (i >= 1 && i + 1 < n) {
        const double hm = sx[i] - sx[i - 1], hp = sx[i + 1] - sx[i];
        inv = 1.0 / (2.0 * (hm + hp) - hm * c_prev);
        cp = hp * inv;
        c_prev = cp;
      }
      scp[i] = cp; sinv[i] = inv;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x + 1; i < n; i += blockDim.x) {      // row i >= 1; ih at row i belongs to the segment left of knot i
    const int64_t o = l * cap + (int64_t)((i - 1) % SPL_RMAX) * G + (i - 1) / SPL_RMAX;
    lu_ih[o] = 1.0 / (sx[i] - sx[i - 1]);
    lu_cp[o] = scp[i]; lu_inv[o] = sinv[i];
  }
}

static int spline_dispatch(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, int64_t max_knots, const double* x,
                           const double* lu_cp, const double* lu_inv, const double* lu_ih, const int32_t* slice_ladder,
                           const double* x_own, const int64_t* ladder_off, const double* y, int64_t n_q, const double* xq, int32_t extrap, double* yq, cudaStream_t stream) {
  IVX_REQUIRE(n_slices >= 0 && max_knots >= 0 && n_q >= 0 && n_q < ((int64_t)1 << 30), IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: bad sizes");
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: extrap must be 0, 1 or 2");
  if (n_slices == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(knot_off && x && y && xq && yq, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval: null pointer");
  IVX_REQUIRE(max_knots <= 1024 * SPL_RMAX, IVX_ERR_UNSUPPORTED, "ivx_spline_fit_eval: at most %d knots per slice (got max_knots %lld)", 1024 * SPL_RMAX, (long long)max_knots);
  // slices longer than the capacity the launch is sized for come back as NaN rows
  const int nq = (int)n_q;
  const bool ladder = lu_cp != nullptr;
#define IVX_SPL(G, R)                                                                                                                        \
  return !ladder ? launch_spline_t<G, R, 0>(n_slices, knot_off, x_off, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, y, nq, xq, extrap, yq, stream) \
         : x_own ? launch_spline_t<G, R, 2>(n_slices, knot_off, x_off, x, lu_cp, lu_inv, lu_ih, slice_ladder, x_own, ladder_off, y, nq, xq, extrap, yq, stream) \
                 : launch_spline_t<G, R, 1>(n_slices, knot_off, x_off, x, lu_cp, lu_inv, lu_ih, slice_ladder, nullptr, nullptr, y, nq, xq, extrap, yq, stream)
  switch (spline_group_width(max_knots)) {
    case 32: IVX_SPL(32, 8);
    case 64: IVX_SPL(64, 8);
    case 128: IVX_SPL(128, 8);
    case 256: IVX_SPL(256, 8);     // measured on B200: 8 rows x 256 threads beats 4 rows x 512 threads (24 vs 30 ms per day)
    case 512: IVX_SPL(512, 8);
    default: IVX_SPL(1024, 8);
  }
#undef IVX_SPL
}

template <int G, int SPL_R, int MODE>
static int launch_spline_list_t(int64_t n_slices, const int32_t* need, const uint32_t* n_need, const int64_t* knot_off, const int64_t* x_off, const double* x,
                                const double* lu_cp, const double* lu_inv, const double* lu_ih, const int32_t* slice_ladder,
                                const double* x_own, const int64_t* ladder_off, const double* y, int n_q, const double* xq, int extrap,
                                double* yq, cudaStream_t stream) {
  using L = SplLayout<G, SPL_R>;
  static DeviceOnce attr_once;
  if (once_per_device(attr_once)) IVX_CUDA(cudaFuncSetAttribute(k_spline_fit_eval_list<G, SPL_R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
  const unsigned blocks = (unsigned)(sm_count() * (G <= 256 ? 4 : 1));
  k_spline_fit_eval_list<G, SPL_R, MODE><<<blocks, L::BT, L::SMEM, stream>>>(n_slices, need, n_need, IVX_SPL_ARGS);
  IVX_LAUNCH_CHECK("k_spline_fit_eval_list");
  return IVX_OK;
}

// the exact solve for the slices flagged in need[]: ladder tables where a slice may use them (MODE 1 when the slices are known
// to sit on their ladders, MODE 2 when that is verified per slice), the general scan otherwise
int ivx::spline_fit_eval_list(int64_t n_slices, const int32_t* need, const uint32_t* n_need, int64_t max_knots, const int64_t* knot_off, const int64_t* x_off,
                              const double* x, const double* lu_cp, const double* lu_inv, const double* lu_ih, const int32_t* slice_ladder,
                              const double* x_own, const int64_t* ladder_off, const double* y, int n_q, const double* xq, int extrap,
                              double* yq, cudaStream_t stream) {
#define IVX_SPLL(G, R)                                                                                                                   \
  return x_own ? launch_spline_list_t<G, R, 2>(n_slices, need, n_need, IVX_SPL_ARGS, stream) : launch_spline_list_t<G, R, 1>(n_slices, need, n_need, IVX_SPL_ARGS, stream)
  switch (spline_group_width(max_knots)) {
    case 32: IVX_SPLL(32, 8);
    case 64: IVX_SPLL(64, 8);
    case 128: IVX_SPLL(128, 8);
    case 256: IVX_SPLL(256, 8);
    case 512: IVX_SPLL(512, 8);
    default: IVX_SPLL(1024, 8);
  }
#undef IVX_SPLL
}

extern "C" int ivx_spline_fit_eval_xoff(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, int64_t max_knots,
                                        const double* x, const double* y, int64_t n_q, const double* xq, int32_t extrap,
                                        double* yq, ivx_stream_t stream_) {
  return spline_dispatch(n_slices, knot_off, x_off, max_knots, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, y, n_q, xq, extrap, yq, (cudaStream_t)stream_);
}

extern "C" int64_t ivx_spline_ladder_table_len(int64_t n_ladders, int64_t max_knots) {
  if (n_ladders < 0 || max_knots < 0 || max_knots > 1024 * SPL_RMAX) return -1;
  return n_ladders * (int64_t)spline_group_width(max_knots) * SPL_RMAX;
}

extern "C" int ivx_spline_ladder_factor(int64_t n_ladders, const int64_t* ladder_off, int64_t max_knots, const double* x,
                                        double* lu_cp, double* lu_inv, double* lu_ih, ivx_stream_t stream_) {
  IVX_REQUIRE(n_ladders >= 0 && max_knots >= 0 && max_knots <= 1024 * SPL_RMAX, IVX_ERR_INVALID_ARG, "ivx_spline_ladder_factor: bad sizes");
  if (n_ladders == 0) return IVX_OK;
  IVX_REQUIRE(ladder_off && x && lu_cp && lu_inv && lu_ih, IVX_ERR_INVALID_ARG, "ivx_spline_ladder_factor: null pointer");
  const int G = spline_group_width(max_knots);
  const size_t smem = (size_t)3 * G * SPL_RMAX * sizeof(double);
  static DeviceOnce attr_once;                       // the attribute is per device, not per process
  if (once_per_device(attr_once)) IVX_CUDA(cudaFuncSetAttribute(k_spline_ladder_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 1024 * SPL_RMAX * (int)sizeof(double)));
  k_spline_ladder_factor<<<(unsigned)n_ladders, 256, smem, (cudaStream_t)stream_>>>(n_ladders, ladder_off, x, G, lu_cp, lu_inv, lu_ih);
  IVX_LAUNCH_CHECK("k_spline_ladder_factor");
  return IVX_OK;
}

extern "C" int ivx_spline_fit_eval_ladder(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, const int32_t* slice_ladder,
                                          int64_t max_knots, const double* x, const double* lu_cp, const double* lu_inv,
                                          const double* lu_ih, const double* y, int64_t n_q, const double* xq, int32_t extrap,
                                          double* yq, ivx_stream_t stream_) {
  IVX_REQUIRE(x_off && slice_ladder && lu_cp && lu_inv && lu_ih, IVX_ERR_INVALID_ARG, "ivx_spline_fit_eval_ladder: x_off, slice_ladder and the three factor tables are required");
  return spline_dispatch(n_slices, knot_off, x_off, max_knots, x, lu_cp, lu_inv, lu_ih, slice_ladder, nullptr, nullptr, y, n_q, xq, extrap, yq, (cudaStream_t)stream_);
}

extern "C" int ivx_spline_fit_eval_ladder_verify(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, const int32_t* slice_ladder,
                                                 const int64_t* ladder_off, int64_t max_knots, const double* x, const double* x_own,
                                                 const double* lu_cp, const double* lu_inv, const double* lu_ih, const double* y,
                                                 int64_t n_q, const double* xq, int32_t extrap, double* yq, ivx_stream_t stream_) {
  IVX_REQUIRE(x_off && slice_ladder && ladder_off && x_own && lu_cp && lu_inv && lu_ih, IVX_ERR_INVALID_ARG,
              "ivx_spline_fit_eval_ladder_verify: x_off, slice_ladder, ladder_off, x_own and the three factor tables are required");
  return spline_dispatch(n_slices, knot_off, x_off, max_knots, x, lu_cp, lu_inv, lu_ih, slice_ladder, x_own, ladder_off, y, n_q, xq, extrap, yq, (cudaStream_t)stream_);
}

extern "C" int ivx_spline_fit_eval(int64_t n_slices, const int64_t* knot_off, int64_t max_knots, const double* x,
                                   const double* y, int64_t n_q, const double* xq, int32_t extrap, double* yq,
                                   void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  (void)workspace; (void)workspace_bytes;   // kept in the ABI; this version needs no scratch
  return ivx_spline_fit_eval_xoff(n_slices, knot_off, nullptr, max_knots, x, y, n_q, xq, extrap, yq, stream_);
}

static cudaError_t strided_smem_attr() {
  static DeviceOnce attr_once;                         // per device, not per process
  if (!once_per_device(attr_once)) return cudaSuccess;
  return cudaFuncSetAttribute(k_spline_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIDED_MAXK * 128 * (int)sizeof(double));
}

extern "C" int ivx_spline_strided(int64_t n_systems, int32_t n_knots, const double* x, const double* y, int32_t n_q,
                                  const double* xq, int32_t extrap, double* yq, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_systems >= 0 && n_knots >= 0 && n_q >= 0, IVX_ERR_INVALID_ARG, "ivx_spline_strided: bad sizes");
  IVX_REQUIRE(n_knots <= STRIDED_MAXK, IVX_ERR_UNSUPPORTED, "ivx_spline_strided: at most %d knots per system (got %d)", STRIDED_MAXK, n_knots);
  IVX_REQUIRE(extrap >= 0 && extrap <= 2, IVX_ERR_INVALID_ARG, "ivx_spline_strided: extrap must be 0, 1 or 2");
  if (n_systems == 0 || n_q == 0) return IVX_OK;
  IVX_REQUIRE(x && y && xq && yq, IVX_ERR_INVALID_ARG, "ivx_spline_strided: null pointer");
  IVX_CUDA(strided_smem_attr());
  k_spline_strided<<<(unsigned)((n_systems + 128 * STRIDED_SPT - 1) / (128 * STRIDED_SPT)), 128, (size_t)n_knots * 128 * sizeof(double), stream>>>(n_systems, n_knots, x, 0, 1, y, n_q, xq, extrap, yq);
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
  IVX_CUDA(strided_smem_attr());
  const int64_t gx = (n_systems + 128 * STRIDED_SPT - 1) / (128 * STRIDED_SPT);
  for (int64_t b0 = 0; b0 < n_batches; b0 += 65535) {   // gridDim.y limit
    const int64_t nb = n_batches - b0 < 65535 ? n_batches - b0 : 65535;
    dim3 grid((unsigned)gx, (unsigned)nb);
    k_spline_strided<<<grid, 128, (size_t)n_knots * 128 * sizeof(double), stream>>>(n_systems, n_knots, x + b0 * x_batch_stride, x_batch_stride, x_knot_stride,
                                               y + b0 * n_knots * n_systems, n_q, xq, extrap, yq + b0 * n_q * n_systems);
    IVX_LAUNCH_CHECK("k_spline_strided(batched)");
  }
  return IVX_OK;
}
