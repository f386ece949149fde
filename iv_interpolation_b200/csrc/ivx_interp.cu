// ivx_interp.cu — kernel A: per-contract time-linear interpolation + forward fill.
//
// Replaces reference src/interpolation/core.py:26-81 for a CSR batch of contracts.
//   plan  (one warp per contract, reads only the SOURCE rows — ~60x fewer than the output):
//         guards, minute grid anchored at min(date), left-join positions (off-grid rows vanish,
//         duplicate stamps widen the axis), and for every on-grid source row the parameters of
//         the interval that starts there: per interpolated column the bracketing valid knots
//         (x0,y0,x1,y1,slope), per fill column the carried value / row index.
//   exec  (tiles of TA_ROWS output rows): each output row finds its interval and evaluates
//         y = slope*(p-x0) + y0 with a separate multiply and add plus np.interp's NaN fallbacks —
//         bit-equal to pandas' Series.interpolate('linear') — and streams the row out.
// HBM traffic: ~81 B written per output row + the (small) interval table read through L1/L2.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr int TA_THREADS = 256;
constexpr int TA_ROWS = 1024;
constexpr int PLAN_WARPS = 4;     // 128 threads, 24 KB of records: five CTAs per SM at 100 registers

// everything exec needs about the interval that starts at one on-grid source row, in ONE record: a row's ~20 parameter
// loads then fall into two or three cache lines shared by the whole warp, at constant offsets from one address
struct __align__(16) IntervalRec {
  int64_t gts;
  int32_t grow, pad_;
  double y0[3], y1[3], slope[3];
  int32_t x0[3], x1[3];
  double fval[IVX_MAX_FILL_COLS];
  int32_t fidx[IVX_MAX_IDX_COLS];
};
static_assert(sizeof(IntervalRec) == 192, "IntervalRec layout");

struct InterpWs {
  // per compacted on-grid source row
  int32_t* gpos;          // dense: searched by exec
  IntervalRec* rec;       // assembled by the plan's backward sweep, read by exec
  // column-wise scratch of the plan's forward sweep (coalesced per column)
  int64_t* s_gts;
  int32_t* s_grow;
  double* s_y0[3];
  int32_t* s_x0[3];
  double* s_fval[IVX_MAX_FILL_COLS];
  int32_t* s_fidx[IVX_MAX_IDX_COLS];
  // per contract
  int32_t* n_on;
  int32_t* p0;
  int64_t* tile_count;
  int64_t* tile_off;  // n_sym + 1
};

InterpWs carve(void* ws, int64_t bytes, int64_t n_sym, int64_t n_src, int n_fill, int n_idx, bool* ok) {
  Carver cv(ws, bytes);
  InterpWs W;
  W.gpos = cv.take<int32_t>(n_src);
  W.rec = cv.take<IntervalRec>(n_src);
  W.s_gts = cv.take<int64_t>(n_src);
  W.s_grow = cv.take<int32_t>(n_src);
  for (int c = 0; c < 3; ++c) { W.s_y0[c] = cv.take<double>(n_src); W.s_x0[c] = cv.take<int32_t>(n_src); }
  for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) W.s_fval[k] = k < n_fill ? cv.take<double>(n_src) : nullptr;
  for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) W.s_fidx[k] = k < n_idx ? cv.take<int32_t>(n_src) : nullptr;
  W.n_on = cv.take<int32_t>(n_sym);
  W.p0 = cv.take<int32_t>(n_sym);
  W.tile_count = cv.take<int64_t>(n_sym + 1);
  W.tile_off = cv.take<int64_t>(n_sym + 2);
  *ok = cv.ok();
  return W;
}

struct PlanParams {
  int64_t n_sym;
  const int64_t* src_off;
  const int64_t* ts;
  const double* icol[3];
  const double* fcol[IVX_MAX_FILL_COLS];
  const uint8_t* ivalid[IVX_MAX_IDX_COLS];
  int32_t n_fill, n_idx;
  int64_t step, max_span, max_pts;
  int32_t min_pts;
  int64_t* out_count;
  uint8_t* sym_status;
  uint32_t* sym_flags;
  InterpWs W;
};

__device__ __forceinline__ int highest_bit_le(uint32_t m, int lane) {  // highest set bit at or below lane, -1 if none
  uint32_t mm = m & (0xffffffffu >> (31 - lane));
  return mm ? 31 - __clz(mm) : -1;
}
__device__ __forceinline__ int lowest_bit_gt(uint32_t m, int lane) {   // lowest set bit strictly above lane, -1 if none
  uint32_t mm = (lane == 31) ? 0u : (m & (0xffffffffu << (lane + 1)));
  return mm ? __ffs(mm) - 1 : -1;
}

__global__ void __launch_bounds__(PLAN_WARPS * 32) k_interp_plan(const PlanParams P) {
  __shared__ IntervalRec s_rec[PLAN_WARPS * 32];   // 24 KB: one record per lane, copied out coalesced
  const int lane = threadIdx.x & 31;
  const int64_t s = blockIdx.x * (int64_t)PLAN_WARPS + (threadIdx.x >> 5);
  if (s >= P.n_sym) return;
  const int64_t a = P.src_off[s], b = P.src_off[s + 1];
  const int64_t n = b - a;
  const InterpWs& W = P.W;

  auto finish = [&](uint8_t status, int64_t count, int32_t n_on, int32_t p0, uint32_t flags) {
    if (lane == 0) {
      P.sym_status[s] = status;
      P.out_count[s] = count;
      W.n_on[s] = n_on;
      W.p0[s] = p0;
      W.tile_count[s] = (count + TA_ROWS - 1) / TA_ROWS;
      if (P.sym_flags) P.sym_flags[s] = flags;
    }
  };

  if (n < P.min_pts) { finish(IVX_SYM_TOO_FEW, 0, 0, 0, 0); return; }        // core.py:26-28

  // ---- span (NaT ignored, like Series.min/max) ----------------------------------------------------
  int64_t tmin = INT64_MAX, tmax = INT64_MIN;
  uint32_t flags = 0;
  int64_t prev_last = IVX_NAT;
  for (int64_t base = a; base < b; base += 128) {       // four chunks of stamps in flight: the sweep is latency-bound
    int64_t tt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = base + 32 * u + lane;
      tt[u] = (r < b) ? __ldg(P.ts + r) : IVX_NAT;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (base + 32 * u >= b) break;
      const int64_t t = tt[u];
      if (t != IVX_NAT) { tmin = min(tmin, t); tmax = max(tmax, t); }
      int64_t tp = __shfl_up_sync(IVX_FULL_MASK, t, 1);
      if (lane == 0) tp = prev_last;
      if (t != IVX_NAT && tp != IVX_NAT && t < tp) flags |= IVX_FLAG_UNSORTED;
      prev_last = __shfl_sync(IVX_FULL_MASK, t, 31);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    tmin = min(tmin, __shfl_xor_sync(IVX_FULL_MASK, tmin, d));
    tmax = max(tmax, __shfl_xor_sync(IVX_FULL_MASK, tmax, d));
    flags |= __shfl_xor_sync(IVX_FULL_MASK, flags, d);
  }
  if (tmax == INT64_MIN) { finish(IVX_SYM_NO_DATES, 0, 0, 0, flags); return; }
  const int64_t span = tmax - tmin;
  if (span > P.max_span) { finish(IVX_SYM_SPAN, 0, 0, 0, flags); return; }     // core.py:36-39
  const int64_t n_grid = span / P.step + 1;                                     // core.py:42-46
  if (n_grid > P.max_pts) { finish(IVX_SYM_TOO_LONG, 0, 0, 0, flags); return; } // core.py:49-51

  // ---- forward sweep: positions and carried-forward state ------------------------------------------
  int32_t carry_on = 0, carry_dups = 0;
  int64_t carry_k = -1;
  int32_t cpos[3] = {-1, -1, -1};
  double cval[3] = {0, 0, 0};
  int32_t first_pos[3] = {INT32_MAX, INT32_MAX, INT32_MAX};
  double cfill[IVX_MAX_FILL_COLS];
  int32_t cidx[IVX_MAX_IDX_COLS];
  int32_t first_sym_pos = INT32_MAX;
#pragma unroll
  for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) cfill[k] = qnan();
#pragma unroll
  for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) cidx[k] = -1;

  for (int64_t base = a; base < b; base += 32) {
    const int64_t r = base + lane;
    const bool inr = r < b;
    // every load of the chunk is issued before its first store (one memory latency per chunk, not one per column: the
    // workspace the sweep writes could alias the columns as far as the compiler knows), and the next chunk's lines are
    // requested into L2 meanwhile — 2 400 contracts are 16 warps per SM, nothing else hides the latency
    const int64_t t = inr ? __ldg(P.ts + r) : IVX_NAT;
    double vc[3], vf[IVX_MAX_FILL_COLS];
    bool vi[IVX_MAX_IDX_COLS];
#pragma unroll
    for (int c = 0; c < 3; ++c) vc[c] = inr ? __ldg(P.icol[c] + r) : qnan();
#pragma unroll
    for (int kf = 0; kf < IVX_MAX_FILL_COLS; ++kf) vf[kf] = (kf < P.n_fill && inr) ? __ldg(P.fcol[kf] + r) : qnan();
#pragma unroll
    for (int ki = 0; ki < IVX_MAX_IDX_COLS; ++ki) vi[ki] = (ki < P.n_idx && inr && P.ivalid[ki]) ? (__ldg(P.ivalid[ki] + r) != 0) : true;
    if ((lane == 0 || lane == 16 || lane == 31) && r + 32 < b) {      // these three rows touch every line of the next chunk
      prefetch_l2(P.ts + r + 32);
#pragma unroll
      for (int c = 0; c < 3; ++c) prefetch_l2(P.icol[c] + r + 32);
#pragma unroll
      for (int kf = 0; kf < IVX_MAX_FILL_COLS; ++kf) if (kf < P.n_fill) prefetch_l2(P.fcol[kf] + r + 32);
    }
    int64_t k = -2;
    bool on = false;
    if (t != IVX_NAT) {
      int64_t off = t - tmin;
      int64_t q = off / P.step;
      on = (q * P.step == off);                                   // core.py:54-55 left join on equality
      k = q;
    }
    const uint32_t m_on = __ballot_sync(IVX_FULL_MASK, on);
    // k of the previous on-grid row
    int src = highest_bit_le(m_on & ~(1u << lane), lane);
    int64_t kprev = __shfl_sync(IVX_FULL_MASK, k, src < 0 ? 0 : src);
    if (src < 0) kprev = carry_k;
    const bool dup = on && (k == kprev);
    const uint32_t m_dup = __ballot_sync(IVX_FULL_MASK, dup);
    const uint32_t le = 0xffffffffu >> (31 - lane);
    const int32_t pos = (int32_t)k + carry_dups + __popc(m_dup & le);
    const int64_t g = a + carry_on + __popc(m_on & (le >> 1));
    if (on) { W.gpos[g] = pos; W.s_gts[g] = t; W.s_grow[g] = (int32_t)r; }

#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = vc[c];
      const bool valid = on && !isnan(v);
      if (valid && isinf(v)) flags |= IVX_FLAG_NONFINITE;
      const uint32_t m_v = __ballot_sync(IVX_FULL_MASK, valid);
      const int sl = highest_bit_le(m_v, lane);
      int32_t ppos = __shfl_sync(IVX_FULL_MASK, pos, sl < 0 ? 0 : sl);
      double pval = __shfl_sync(IVX_FULL_MASK, v, sl < 0 ? 0 : sl);
      if (sl < 0) { ppos = cpos[c]; pval = cval[c]; }
      if (on) { W.s_x0[c][g] = ppos; W.s_y0[c][g] = pval; }
      if (m_v) {
        const int last = 31 - __clz(m_v);
        if (first_pos[c] == INT32_MAX) first_pos[c] = __shfl_sync(IVX_FULL_MASK, pos, __ffs(m_v) - 1);
        cpos[c] = __shfl_sync(IVX_FULL_MASK, pos, last);
        cval[c] = __shfl_sync(IVX_FULL_MASK, v, last);
      }
    }
#pragma unroll
    for (int kf = 0; kf < IVX_MAX_FILL_COLS; ++kf) {
      if (kf < P.n_fill) {
        const double v = vf[kf];
        const bool valid = on && !isnan(v);
        const uint32_t m_v = __ballot_sync(IVX_FULL_MASK, valid);
        const int sl = highest_bit_le(m_v, lane);
        double pval = __shfl_sync(IVX_FULL_MASK, v, sl < 0 ? 0 : sl);
        if (sl < 0) pval = cfill[kf];
        if (on) W.s_fval[kf][g] = pval;
        if (m_v) cfill[kf] = __shfl_sync(IVX_FULL_MASK, v, 31 - __clz(m_v));
      }
    }
#pragma unroll
    for (int ki = 0; ki < IVX_MAX_IDX_COLS; ++ki) {
      if (ki < P.n_idx) {
        const bool valid = on && vi[ki];
        const uint32_t m_v = __ballot_sync(IVX_FULL_MASK, valid);
        const int sl = highest_bit_le(m_v, lane);
        int32_t pidx = __shfl_sync(IVX_FULL_MASK, (int32_t)r, sl < 0 ? 0 : sl);
        if (sl < 0) pidx = cidx[ki];
        if (on) W.s_fidx[ki][g] = pidx;
        if (m_v) {
          if (ki == 0 && first_sym_pos == INT32_MAX) first_sym_pos = __shfl_sync(IVX_FULL_MASK, pos, __ffs(m_v) - 1);
          cidx[ki] = __shfl_sync(IVX_FULL_MASK, (int32_t)r, 31 - __clz(m_v));
        }
      }
    }
    if (m_on) carry_k = __shfl_sync(IVX_FULL_MASK, k, 31 - __clz(m_on));
    carry_on += __popc(m_on);
    carry_dups += __popc(m_dup);
  }
  const int32_t n_on = carry_on;
  const int64_t n_pos = n_grid + carry_dups;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) flags |= __shfl_xor_sync(IVX_FULL_MASK, flags, d);

  // ---- backward sweep: next valid knot of each interpolated column, slopes ----------------------------
  // the forward sweep wrote gpos / s_* to global memory lane by lane; the backward sweep reads them back with another
  // lane-to-row mapping: order the two (shuffles and ballots give no memory ordering between lanes)
  __syncwarp();
  int32_t npos[3] = {-1, -1, -1};
  double nval[3] = {0, 0, 0};
  for (int64_t hi = a + n_on; hi > a; hi -= 32) {
    const int64_t g = hi - 32 + lane;              // chunk [hi-32, hi), may start before a
    const bool ing = g >= a;
    // the chunk's loads together, ahead of the shuffles (one latency per chunk)
    const int32_t pos = ing ? W.gpos[g] : -1;
    const int64_t gts = ing ? W.s_gts[g] : 0;
    const int32_t grow = ing ? W.s_grow[g] : -1;
    int32_t x0s[3];
    double y0s[3], fv[IVX_MAX_FILL_COLS];
    int32_t fi[IVX_MAX_IDX_COLS];
#pragma unroll
    for (int c = 0; c < 3; ++c) { x0s[c] = ing ? W.s_x0[c][g] : -1; y0s[c] = ing ? W.s_y0[c][g] : 0.0; }
#pragma unroll
    for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) fv[k] = (ing && k < P.n_fill) ? W.s_fval[k][g] : 0.0;
#pragma unroll
    for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) fi[k] = (ing && k < P.n_idx) ? W.s_fidx[k][g] : -1;
    IntervalRec& R = s_rec[threadIdx.x];             // this lane's record, assembled in shared memory
    R.gts = gts; R.grow = grow; R.pad_ = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int32_t x0 = x0s[c];
      const double y0 = y0s[c];
      const bool valid = ing && (x0 == pos);       // the row's own value is a knot
      const uint32_t m_v = __ballot_sync(IVX_FULL_MASK, valid);
      const int sl = lowest_bit_gt(m_v, lane);
      int32_t x1 = __shfl_sync(IVX_FULL_MASK, pos, sl < 0 ? 0 : sl);
      double y1 = __shfl_sync(IVX_FULL_MASK, y0, sl < 0 ? 0 : sl);
      if (sl < 0) { x1 = npos[c]; y1 = nval[c]; }
      double slope = 0.0;
      if (x0 >= 0 && x1 >= 0) slope = __ddiv_rn(__dsub_rn(y1, y0), (double)(x1 - x0));   // np.interp slope
      R.x0[c] = x0; R.y0[c] = y0; R.x1[c] = x1; R.y1[c] = y1; R.slope[c] = slope;
      if (m_v) {
        const int first = __ffs(m_v) - 1;
        npos[c] = __shfl_sync(IVX_FULL_MASK, pos, first);
        nval[c] = __shfl_sync(IVX_FULL_MASK, y0, first);
      }
    }
#pragma unroll
    for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) R.fval[k] = fv[k];
#pragma unroll
    for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) R.fidx[k] = fi[k];
    __syncwarp();
    // the chunk's records [hi-32, hi) are contiguous in global memory: the warp copies them out 16 bytes per lane at a time
    {
      const int64_t g_lo = max(hi - 32, a);
      const int first_lane = (int)(g_lo - (hi - 32));
      const uint4* src = reinterpret_cast<const uint4*>(&s_rec[(threadIdx.x & ~31) + first_lane]);
      uint4* dst = reinterpret_cast<uint4*>(W.rec + g_lo);
      const int n16 = (int)(hi - g_lo) * (int)(sizeof(IntervalRec) / 16);
      for (int i = lane; i < n16; i += 32) dst[i] = src[i];
    }
    __syncwarp();
  }

  // ---- dropna (core.py:74): rows before the first valid value of any key column go ------------------------
  int32_t p0 = 0;
  bool empty = false;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (first_pos[c] == INT32_MAX) empty = true; else p0 = max(p0, first_pos[c]);
  }
  if (P.n_idx > 0) {
    if (first_sym_pos == INT32_MAX) empty = true; else p0 = max(p0, first_sym_pos);
  }
  if (empty) { finish(IVX_SYM_EMPTY_OUT, 0, n_on, 0, flags); return; }             // core.py:76-78
  finish(IVX_SYM_OK, n_pos - p0, n_on, p0, flags);
}

// single-block exclusive scan (n is a symbol count: at most a few million)
__global__ void __launch_bounds__(1024) k_scan_i64(int64_t n, const int64_t* __restrict__ in, int64_t* __restrict__ out) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < n ? in[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t u = __shfl_up_sync(IVX_FULL_MASK, incl, d);
      if (lane >= d) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane];
      int64_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int64_t u = __shfl_up_sync(IVX_FULL_MASK, wi, d);
        if (lane >= d) wi += u;
      }
      s_warp[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    const int64_t carry = s_carry;
    if (i < n) out[i] = carry + s_warp[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_carry;
}

struct ExecParams {
  int64_t n_sym;
  const int64_t* src_off;
  const int64_t* out_off;
  int64_t step;
  int32_t n_fill, n_idx;
  int64_t* o_ts;
  double* o_interp[3];
  double* o_fill[IVX_MAX_FILL_COLS];
  int32_t* o_idx[IVX_MAX_IDX_COLS];
  int32_t* o_src_row;
  int32_t* o_pos;
  uint8_t* o_keep;
  uint8_t* o_isint;
  InterpWs W;
};

__global__ void __launch_bounds__(TA_THREADS) k_interp_exec(const ExecParams P) {
  __shared__ int64_t s_sym;
  __shared__ int32_t s_g0;
  __shared__ uint32_t s_bits[TA_ROWS / 32];    // bit r set: a source knot sits on row r of the tile
  __shared__ int32_t s_pref[TA_ROWS / 32];     // knots before each 32-row group
  const InterpWs& W = P.W;
  const int64_t tile = blockIdx.x;
  if (tile >= W.tile_off[P.n_sym]) return;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    s_bits[threadIdx.x] = 0;
    const int64_t lo = warp_upper_le<int64_t>(W.tile_off, 0, P.n_sym, tile, lane);   // largest s with tile_off[s] <= tile
    if (lane == 0) s_sym = lo;
  }
  __syncthreads();
  const int64_t s = s_sym;
  const int64_t a = P.src_off[s];
  const int32_t n_on = W.n_on[s];
  const int32_t p0 = W.p0[s];
  const int64_t obase = P.out_off[s];
  const int64_t count = P.out_off[s + 1] - obase;
  const int64_t q0 = (tile - W.tile_off[s]) * TA_ROWS;
  const int32_t* __restrict__ gpos = W.gpos + a;
  const int32_t P0 = p0 + (int32_t)q0;         // position of the tile's first row on the contract's axis

  // interval of the first row (gpos[0] == 0 <= P0), then one bit per knot inside the tile: a row's interval is
  // g0 + (number of knots at or before it), i.e. a prefix popcount instead of a binary search per row
  if (threadIdx.x < 32) {
    const int32_t g0 = (int32_t)warp_upper_le<int32_t>(gpos, 0, n_on, P0, lane);
    if (lane == 0) s_g0 = g0;
  }
  __syncthreads();
  const int32_t g0 = s_g0;
  for (int32_t k = g0 + 1 + (int32_t)threadIdx.x; k < n_on; k += TA_THREADS) {
    const int32_t d = __ldg(gpos + k) - P0;
    if (d >= TA_ROWS) break;
    atomicOr(&s_bits[d >> 5], 1u << (d & 31));
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int c = __popc(s_bits[threadIdx.x]);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(IVX_FULL_MASK, incl, d);
      if (lane >= d) incl += u;
    }
    s_pref[threadIdx.x] = incl - c;
  }
  __syncthreads();

#pragma unroll 1
  for (int it = 0; it < TA_ROWS / TA_THREADS; ++it) {
    const int r = it * TA_THREADS + threadIdx.x;
    const int64_t q = q0 + r;
    const bool act = q < count;
    const int32_t p = P0 + r;
    const int32_t l2 = g0 + s_pref[r >> 5] + __popc(s_bits[r >> 5] & (0xffffffffu >> (31 - (r & 31))));
    if (!act) continue;
    const int64_t g = a + l2;
    const int32_t gp = __ldg(W.gpos + g);
    const IntervalRec* __restrict__ R = W.rec + g;
    const int64_t o = obase + q;
    st_stream(P.o_ts + o, __ldg(&R->gts) + (int64_t)(p - gp) * P.step);
    bool keep = true;
    // every parameter of the interval is requested up front (independent loads, one latency round) — no load waits on a
    // branch taken on an earlier one
    int32_t x0v[3], x1v[3];
    double y0v[3], slv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x0v[c] = __ldg(&R->x0[c]); x1v[c] = __ldg(&R->x1[c]);
      y0v[c] = __ldg(&R->y0[c]); slv[c] = __ldg(&R->slope[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int32_t x0 = x0v[c], x1 = x1v[c];
      const double y0 = y0v[c];
      double v;
      if (x0 < 0) {
        v = qnan();                                              // leading NaNs are preserved
      } else if (x1 < 0 || p == x0) {
        v = y0;                                                  // a knot itself / flat tail
      } else {
        const double sl = slv[c];
        v = __dadd_rn(__dmul_rn(sl, (double)(p - x0)), y0);      // np.interp: slope*(x-x0)+y0, no FMA
        if (isnan(v)) {
          const double y1 = __ldg(&R->y1[c]);
          v = __dadd_rn(__dmul_rn(sl, (double)(p - x1)), y1);
          if (isnan(v) && y0 == y1) v = y0;
        }
      }
      keep = keep && !isnan(v);
      st_stream(P.o_interp[c] + o, v);
    }
#pragma unroll
    for (int k = 0; k < IVX_MAX_FILL_COLS; ++k)
      if (k < P.n_fill && P.o_fill[k]) st_stream(P.o_fill[k] + o, __ldg(&R->fval[k]));
#pragma unroll
    for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) {
      if (k < P.n_idx) {
        const int32_t ix = __ldg(&R->fidx[k]);
        if (k == 0) keep = keep && ix >= 0;
        if (P.o_idx[k]) st_stream(P.o_idx[k] + o, ix);
      }
    }
    if (P.o_src_row) st_stream(P.o_src_row + o, p == gp ? __ldg(&R->grow) : -1);
    if (P.o_pos) st_stream(P.o_pos + o, p);
    if (P.o_keep) P.o_keep[o] = keep ? 1 : 0;
    if (P.o_isint) P.o_isint[o] = 0;                             // core.py:71, always False
  }
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_interp_workspace_bytes(int64_t n_sym, int64_t n_src, int32_t n_fill, int32_t n_idx) {
  if (n_sym < 0 || n_src < 0 || n_fill < 0 || n_idx < 0) return -1;
  int64_t per_row = 4 + (int64_t)sizeof(IntervalRec) + 8 + 4 + 3 * (8 + 4) + 8 * (int64_t)n_fill + 4 * (int64_t)n_idx;
  int64_t arrays = 2 + 2 + 6 + n_fill + n_idx + 4;
  return per_row * n_src + (4 + 4 + 8 + 8) * (n_sym + 2) + 256 * (arrays + 2);
}

extern "C" int ivx_exclusive_scan_i64(int64_t n, const int64_t* in, int64_t* out, ivx_stream_t stream_) {
  IVX_REQUIRE(n >= 0 && out && (n == 0 || in), IVX_ERR_INVALID_ARG, "ivx_exclusive_scan_i64: bad arguments");
  k_scan_i64<<<1, 1024, 0, (cudaStream_t)stream_>>>(n, in, out);
  IVX_LAUNCH_CHECK("k_scan_i64");
  return IVX_OK;
}

extern "C" int ivx_interp_plan(int64_t n_sym, const int64_t* src_off, int64_t n_src, const int64_t* src_ts,
                               const double* const* interp_cols, int32_t n_fill, const double* const* fill_cols,
                               int32_t n_idx, const uint8_t* const* idx_valid, int64_t step_ticks,
                               int64_t max_span_ticks, int64_t max_grid_points, int32_t min_points,
                               int64_t* out_count, int64_t* out_off, uint8_t* sym_status, uint32_t* sym_flags,
                               void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_src >= 0, IVX_ERR_INVALID_ARG, "ivx_interp_plan: negative size");
  IVX_REQUIRE(n_src < ((int64_t)1 << 31), IVX_ERR_TOO_LARGE, "ivx_interp_plan: n_src %lld exceeds 2^31 per call; split the batch", (long long)n_src);
  IVX_REQUIRE(n_fill >= 0 && n_fill <= IVX_MAX_FILL_COLS && n_idx >= 0 && n_idx <= IVX_MAX_IDX_COLS, IVX_ERR_INVALID_ARG, "ivx_interp_plan: n_fill/n_idx out of range");
  IVX_REQUIRE(step_ticks >= 1 && max_grid_points >= 1 && max_grid_points < ((int64_t)1 << 30), IVX_ERR_INVALID_ARG, "ivx_interp_plan: bad step/max_grid_points");
  IVX_REQUIRE(src_off && out_count && out_off && sym_status && workspace && interp_cols, IVX_ERR_INVALID_ARG, "ivx_interp_plan: null pointer");
  IVX_REQUIRE(n_src == 0 || (src_ts && interp_cols[0] && interp_cols[1] && interp_cols[2]), IVX_ERR_INVALID_ARG, "ivx_interp_plan: null input column");
  IVX_REQUIRE(workspace_bytes >= ivx_interp_workspace_bytes(n_sym, n_src, n_fill, n_idx), IVX_ERR_INVALID_ARG, "ivx_interp_plan: workspace too small");
  bool ok;
  PlanParams P;
  P.W = carve(workspace, workspace_bytes, n_sym, n_src, n_fill, n_idx, &ok);
  IVX_REQUIRE(ok, IVX_ERR_INVALID_ARG, "ivx_interp_plan: workspace carve overflow");
  P.n_sym = n_sym; P.src_off = src_off; P.ts = src_ts;
  for (int c = 0; c < 3; ++c) P.icol[c] = interp_cols[c];
  for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) P.fcol[k] = k < n_fill ? fill_cols[k] : nullptr;
  for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) P.ivalid[k] = k < n_idx ? idx_valid[k] : nullptr;
  for (int k = 0; k < n_fill; ++k) IVX_REQUIRE(n_src == 0 || P.fcol[k], IVX_ERR_INVALID_ARG, "ivx_interp_plan: null fill column %d", k);
  P.n_fill = n_fill; P.n_idx = n_idx;
  P.step = step_ticks; P.max_span = max_span_ticks; P.max_pts = max_grid_points; P.min_pts = min_points;
  P.out_count = out_count; P.sym_status = sym_status; P.sym_flags = sym_flags;
  if (n_sym > 0) {
    k_interp_plan<<<(unsigned)((n_sym + PLAN_WARPS - 1) / PLAN_WARPS), PLAN_WARPS * 32, 0, stream>>>(P);
    IVX_LAUNCH_CHECK("k_interp_plan");
  }
  k_scan_i64<<<1, 1024, 0, stream>>>(n_sym, out_count, out_off);
  IVX_LAUNCH_CHECK("k_scan_i64(out_off)");
  k_scan_i64<<<1, 1024, 0, stream>>>(n_sym, P.W.tile_count, P.W.tile_off);
  IVX_LAUNCH_CHECK("k_scan_i64(tile_off)");
  return IVX_OK;
}

extern "C" int ivx_interp_exec(int64_t n_sym, const int64_t* src_off, int64_t n_src, const int64_t* out_off,
                               int64_t n_out, int64_t step_ticks, int32_t n_fill, int32_t n_idx, int64_t* o_ts,
                               double* const* o_interp, double* const* o_fill, int32_t* const* o_idx,
                               int32_t* o_src_row, int32_t* o_pos, uint8_t* o_keep, uint8_t* o_is_interpolated,
                               const void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_src >= 0 && n_out >= 0, IVX_ERR_INVALID_ARG, "ivx_interp_exec: negative size");
  IVX_REQUIRE(n_fill >= 0 && n_fill <= IVX_MAX_FILL_COLS && n_idx >= 0 && n_idx <= IVX_MAX_IDX_COLS, IVX_ERR_INVALID_ARG, "ivx_interp_exec: n_fill/n_idx out of range");
  if (n_out == 0 || n_sym == 0) return IVX_OK;
  IVX_REQUIRE(src_off && out_off && workspace && o_ts && o_interp && o_interp[0] && o_interp[1] && o_interp[2], IVX_ERR_INVALID_ARG, "ivx_interp_exec: null pointer");
  IVX_REQUIRE(workspace_bytes >= ivx_interp_workspace_bytes(n_sym, n_src, n_fill, n_idx), IVX_ERR_INVALID_ARG, "ivx_interp_exec: workspace too small");
  bool ok;
  ExecParams P;
  P.W = carve(const_cast<void*>(workspace), workspace_bytes, n_sym, n_src, n_fill, n_idx, &ok);
  P.n_sym = n_sym; P.src_off = src_off; P.out_off = out_off; P.step = step_ticks;
  P.n_fill = n_fill; P.n_idx = n_idx; P.o_ts = o_ts;
  for (int c = 0; c < 3; ++c) P.o_interp[c] = o_interp[c];
  for (int k = 0; k < IVX_MAX_FILL_COLS; ++k) P.o_fill[k] = (k < n_fill && o_fill) ? o_fill[k] : nullptr;
  for (int k = 0; k < IVX_MAX_IDX_COLS; ++k) P.o_idx[k] = (k < n_idx && o_idx) ? o_idx[k] : nullptr;
  P.o_src_row = o_src_row; P.o_pos = o_pos; P.o_keep = o_keep; P.o_isint = o_is_interpolated;
  // upper bound on the tile count known without a device read: ceil per contract <= n_out/TA_ROWS + n_sym
  const int64_t max_tiles = n_out / TA_ROWS + n_sym + 1;
  IVX_REQUIRE(max_tiles < ((int64_t)1 << 31), IVX_ERR_TOO_LARGE, "ivx_interp_exec: too many tiles; split the batch");
  k_interp_exec<<<(unsigned)max_tiles, TA_THREADS, 0, stream>>>(P);
  IVX_LAUNCH_CHECK("k_interp_exec");
  return IVX_OK;
}
