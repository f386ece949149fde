// ivx_bridge.cu — the data bridge: interpolated minute rows -> synthetic 1-minute OHLCV candles.
//
// Replaces, for a CSR batch of symbols, reference src/data_bridge/ohlcv_converter.py:138-400
// (_generate_ohlcv_from_interpolated and its four candle strategies, _process_volume, _validate_ohlcv_quality) and
// complete_pipeline.py:473-510 (_generate_ohlcv_candles).  The reference draws from numpy's global generator row by
// row; here the random input arrives as TAPES in the reference's call order — uniforms u (k per emitted row, k fixed
// per strategy), standard exponentials e (one per emitted row whose volume is missing), standard normals z (one per
// emitted row, trend_following only) — so with tapes drawn from an identically seeded RandomState every candle is
// bit-equal to the reference's (tests/golden/bridge_cases.npz).
//
// This file is compiled with --fmad=false: numpy evaluates every expression with separate multiplies and adds, and a
// fused multiply-add here would change last bits.  round(x, 4) on an np.float64 is rint(x * 1e4) / 1e4.
//
// ivx_bridge_plan (k_bridge_count_seg, two scans, k_bridge_sym_off): the candles and exponential draws that precede every
//   2048-row segment and every symbol.  The global candle index of a row is the number of emitting rows before it —
//   symbol boundaries do not enter — so these prefixes place every segment and every symbol in the outputs and the tapes.
// k_bridge_ohlcv_seg: one warp per SEGMENT (k_bridge_ohlcv: one warp per symbol), 32 rows per step.  Loads are coalesced
//   (a lane's k uniforms are contiguous, the warp's 32 k are too); emitted rows are compacted with a ballot.  What cannot
//   be row-parallel is the dependence of a candle on the ROUNDED close of the previous emitted candle (spread_simulation:
//   trend bias; trend_following: the last five closes).  For spread_simulation that chain contracts by 0.15 per emitted
//   row, which is used twice, both times with an exact check instead of an error bound: inside a chunk the 32 rows are
//   walked in parallel from a 12-rank warm-up (bridge_piece), and a segment that starts inside a symbol walks the 256
//   rows before it (k_bridge_ohlcv_seg + k_bridge_check); a failed check falls back to the sequential walk.  Opens,
//   highs, lows, volumes and the quality checks are formed in parallel once every lane knows its predecessor's close.
// HBM traffic: 16 B per input row + 8 k B of tape + 56 B per emitted candle.
#include "ivx_common.cuh"

namespace ivx {
namespace {

enum : int { ST_SPREAD = 0, ST_MIDPOINT = 1, ST_TREND = 2, ST_SIMPLE = 3, ST_PIPELINE = 4 };
constexpr uint32_t Q_HIGH_LT_LOW = 1, Q_HIGH_LT_OC = 2, Q_LOW_GT_OC = 4, Q_SPREAD = 8, Q_NONPOSITIVE = 16;

__device__ __forceinline__ int u_per_row(int strategy) {
  return strategy == ST_SPREAD ? 5 : strategy == ST_MIDPOINT ? 2 : strategy == ST_PIPELINE ? 4 : strategy == ST_SIMPLE ? 1 : 0;
}
__device__ __forceinline__ bool row_ok(double b) { return !(isnan(b) || b <= 0.0); }            // ohlcv_converter.py:152
__device__ __forceinline__ bool vol_missing(double v) { return isnan(v) || v <= 0.0; }          // :373
// round(x, d) of a PYTHON float: the reference's rows come from df.iterrows() on a frame that holds strings and stamps next
// to the prices (ohlcv_converter.py:150, complete_pipeline.py:476), so prices, draws and volumes are Python floats and
// round() is the correctly rounded decimal rounding of the EXACT binary value, half to even — not numpy's
// rint(x * 10^d) / 10^d, whose rounded product invents ties (100.35015 -> 100.3502 under numpy, 100.3501 in Python).
//   n = the integer nearest to the exact x * 10^d:  p = RN(x s), e = x s - p exactly (one fma), a = rint(p), f = p - a
//   exactly; the exact value a + f + e lies beyond the half-way point on f's side iff e points the same way and
//   |e| > 1/2 - |f| (both exact) — then rint's answer moves by one; a true tie (e = 0, |f| = 1/2) keeps rint's even choice.
//   result = RN(n / 10^d), which is what Python's dtoa/strtod round trip yields.  The division is by a constant: with
//   rc = RN(1/s), q0 = RN(n rc), r = n - s q0 (exact in one fma), q = RN(q0 + r rc) is the correctly rounded quotient
//   (Markstein) — three dependent operations instead of a general division; this sits in the serial chain of the carried
//   close.  Compared bit for bit with Python's round over millions of values in tests/test_gpu_bridge.py.
//   Large |x| (ulp(x) above one unit of the rounded place: 2^39 for 4 decimals, 2^33 for 6; also inf and NaN) round to
//   themselves; below that x * 10^d stays under 2^53 and every integer involved is exact.
template <int D>
__device__ __forceinline__ double round_np(double x) {
  constexpr double s = D == 4 ? 1e4 : 1e6;
  constexpr double rc = D == 4 ? 1e-4 : 1e-6;            // RN(1/s)
  constexpr double BIG = D == 4 ? 549755813888.0 : 8589934592.0;      // 2^39, 2^33
  if (!(fabs(x) < BIG)) return x;
  const double p = x * s;
  double a = rint(p);
  const double f = p - a;                                // exact
  // the residual can only matter within |e| <= ulp(p)/2 of the half-way point: a rare branch (the common path is rint alone)
  if (0.5 - fabs(f) <= fabs(p) * 2.220446049250313e-16) {
    const double e = fma(x, s, -p);                      // exact
    if (fabs(p) < 4503599627370496.0) {                  // below 2^52
      const double g = 0.5 - fabs(f);
      if ((f > 0.0 ? e : -e) > g) a += f > 0.0 ? 1.0 : -1.0;
    } else if (fabs(e) == 0.5) {                         // p is an integer and the exact product sits half way to the next one: even wins
      const double half = 0.5 * p;
      if (half != rint(half)) a = p + (e + e);
    }
  }
  const double q0 = a * rc;
  const double r = fma(-s, q0, a);
  return fma(r, rc, q0);
}
__device__ __forceinline__ double uni(double lo, double hi, double u) { return lo + (hi - lo) * u; }   // RandomState.uniform

struct BridgeParams {
  int64_t n_sym;
  const int64_t* in_off;
  const double *base, *volume;
  int strategy;
  double base_spread_percent, volatility_factor;
  const double *u, *e, *z;
  const int64_t *out_off, *e_off;
  int64_t* o_row;                                           // may be null in the generated-tape form (o_ts stands for it)
  double *o_open, *o_high, *o_low, *o_close, *o_volume, *o_source;
  uint32_t* sym_quality;
  // GEN form (ivx_bridge_ohlcv_philox): no tapes in memory — element i of the uniform / exponential tape is Philox(seed, stream, i),
  // the very numbers ivx_random_tape would have written (40 B written and 40 B read per candle less), and the candle's stamp
  // is copied from the row's (no gather pass behind the bridge)
  uint64_t seed;
  uint32_t stream_u, stream_e;
  const int64_t* ts;
  int64_t* o_ts;
};

// the K uniforms of candle o: elements o K .. o K + K - 1 of the tape, i.e. (K + 1) / 2 Philox blocks (K odd: the run starts on
// the second double of its first block for odd o)
template <int K>
__device__ __forceinline__ void gen_uniforms(const BridgeParams& P, int64_t o, double (&uu)[K > 0 ? K : 1]) {
  constexpr int NB = (K + 1) / 2;
  const uint64_t e0 = (uint64_t)o * (uint64_t)K;
  double v[2 * NB + 2];
#pragma unroll
  for (int j = 0; j < NB; ++j) philox_pair((e0 >> 1) + j, P.stream_u, P.seed, v[2 * j], v[2 * j + 1]);
  const bool odd = (K & 1) && (e0 & 1);
#pragma unroll
  for (int k = 0; k < K; ++k) uu[k] = odd ? v[k + 1] : v[k];
}

// closes of the candles emitted so far in the current symbol (rounded, newest first) — warp-uniform
struct BridgeState {
  double pc1, pc2, pc3, pc4, pc5;
  int n_prev;
};
__device__ __forceinline__ BridgeState fresh_state() { return BridgeState{0.0, 0.0, 0.0, 0.0, 0.0, 0}; }

constexpr int CHAIN_W = 12;                                // warm-up ranks of the parallel close walk: 0.15^12 = 1.3e-10
struct BridgeSmem {
  double b[4][32], c[4][32], k[4][32], f[4][32];           // per warp: base, close offset and grid step constant by rank; final closes by rank
};

// Rows [a, b) of ONE symbol, in chunks of 32, continuing from `st`.  o_base / e_base: candles and exponential draws before
// row a.  WRITE = false walks the close chain only (the warm-up of a segment that starts inside a symbol).
template <int STRATEGY, bool WRITE, bool GEN = false>
__device__ __forceinline__ void bridge_piece(const BridgeParams& P, const int64_t a, const int64_t b, int64_t& o_base, int64_t& e_base,
                                             BridgeState& st, uint32_t& qflags, BridgeSmem& sm) {
  const int lane = lane_id();
  const uint32_t lt = (1u << lane) - 1u;
  constexpr int K = STRATEGY == ST_SPREAD ? 5 : STRATEGY == ST_MIDPOINT ? 2 : STRATEGY == ST_PIPELINE ? 4 : STRATEGY == ST_SIMPLE ? 1 : 0;
  double (*s_b)[32] = sm.b; double (*s_c)[32] = sm.c; double (*s_k)[32] = sm.k; double (*s_f)[32] = sm.f;
  double &pc1 = st.pc1, &pc2 = st.pc2, &pc3 = st.pc3, &pc4 = st.pc4, &pc5 = st.pc5;
  int& n_prev = st.n_prev;
  // base/volume of the next chunk are requested before this chunk's (latency-bound) close walk
  double bp_next = (a + lane < b) ? ld_stream(P.base + a + lane) : qnan();
  double vin_next = (a + lane < b) ? ld_stream(P.volume + a + lane) : 1.0;
  for (int64_t r0 = a; r0 < b; r0 += 32) {
    const int64_t r = r0 + lane;
    const bool in = r < b;
    const double bp = bp_next, vin = vin_next;
    {
      const int64_t rn = r + 32;
      bp_next = rn < b ? ld_stream(P.base + rn) : qnan();
      vin_next = rn < b ? ld_stream(P.volume + rn) : 1.0;
    }
    const bool ok = in && row_ok(bp);
    const bool nv = ok && vol_missing(vin);
    const uint32_t okm = __ballot_sync(IVX_FULL_MASK, ok), nvm = __ballot_sync(IVX_FULL_MASK, nv);
    const int64_t o = o_base + __popc(okm & lt);
    double uu[K > 0 ? K : 1];
    double ev = 0.0;
    int64_t tsv = 0;
    if (GEN) {
#pragma unroll
      for (int k = 0; k < K; ++k) uu[k] = 0.0;
      if (WRITE && ok && P.o_ts) tsv = ld_stream(P.ts + r);
      if (K > 0 && ok) gen_uniforms<K>(P, o, uu);
      if (WRITE && nv) {
        const uint64_t ei = (uint64_t)(e_base + __popc(nvm & lt));
        double ea, eb;
        philox_pair(ei >> 1, P.stream_e, P.seed, ea, eb);
        ev = exp_variate((ei & 1) ? eb : ea);
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) uu[k] = (ok && (WRITE || k == 0 || k == 2)) ? ld_stream(P.u + o * K + k) : 0.0;
      ev = (WRITE && nv) ? ld_stream(P.e + e_base + __popc(nvm & lt)) : 0.0;
    }
    double op = 0.0, hp = 0.0, lp = 0.0, cp = 0.0;         // unrounded open/high/low, ROUNDED close
    if (STRATEGY == ST_SPREAD) {
      const double vm = uni(0.5, P.volatility_factor, uu[0]);                  // :215
      const double spread = bp * P.base_spread_percent * vm;                   // :216
      const double third = spread / 3.0;
      const double oo = uni(-third, third, uu[1]), co = uni(-third, third, uu[2]);
      // ---- the close chain: close[j] = round4(base[j] + (co[j] + ((base[j] - close[j-1]) * 0.3) * 0.5)) over the emitted rows.
      // Rounding makes it non-affine (no scan), but it CONTRACTS: a change of close[j-1] moves close[j] by 0.15 of it, and
      // two walks that ever round to the same value stay identical.  So every emitted row of the chunk is walked in
      // parallel: the rows are compacted by rank into shared memory; lane j starts CHAIN_W ranks back from a guess (ranks
      // 0..CHAIN_W start from the true carried close) and walks up to its own rank — 13 dependent steps for the whole chunk
      // instead of 32.  The result is then PROVEN, not assumed: lane j's value for rank j-1 must equal lane j-1's own final
      // value bit for bit; by induction from the exact lanes every close is then the sequential one.  A chunk that fails
      // the check (expected ~1e-3 of chunks: a 1e-9 residue straddling a rounding boundary) is redone sequentially.
      double prev = 0.0;                                   // my predecessor's rounded close
      bool has_prev = false;
      const int wi = threadIdx.x >> 5;
      const int rank = __popc(okm & lt), n = __popc(okm);
      __syncwarp();                                        // the previous chunk's readers are done with the staging rows
      if (ok) {
        s_b[wi][rank] = bp; s_c[wi][rank] = co;
        // the step in units of 1e-4: close[j] * 1e4 ~ rint(k[j] - 0.15 * (close[j-1] * 1e4)); no trend bias on a symbol's first candle
        s_k[wi][rank] = ((rank == 0 && n_prev == 0) ? bp + co : bp * 1.15 + co) * 1e4;
      }
      __syncwarp();
      // Only the LAST step of a lane has to be the reference's arithmetic: the check below proves that the predecessor's
      // close a lane arrived with IS its neighbour's final close, bit for bit, and the last step then makes the lane's own
      // close the sequential one (induction from lane 0, which starts from the carried close).  The steps before it only
      // have to land on the same 1e-4 grid point, so they run on the grid: one fused multiply-add and one rint per rank
      // instead of the five separately rounded operations and the decimal rounding of the exact step (they differ from
      // the exact walk by ~1e-10 of a grid step; a lane that does land elsewhere fails the check and the chunk is redone).
      double a4 = pc1 * 1e4;                               // the predecessor's close in units of 1e-4
      double p = pc1, p_before = pc1;
      if (lane < n) {
        int i = lane - CHAIN_W;
        if (lane > CHAIN_W) a4 = s_b[wi][i - 1] * 1e4;     // the guess: no trend bias at the row before the walk
#pragma unroll
        for (int t = 0; t < CHAIN_W; ++t, ++i)
          if (i >= 0) a4 = rint(fma(-0.15, a4, s_k[wi][i]));
        if (lane > 0) {                                    // a4 / 1e4, correctly rounded (see round_np): the close a4 stands for
          const double q0 = a4 * 1e-4;
          p_before = fma(fma(-1e4, q0, a4), 1e-4, q0);
        }
        const double tb_ = (lane == 0 && n_prev == 0) ? 0.0 : (s_b[wi][lane] - p_before) * 0.3;      // :219-222
        p = round_np<4>(s_b[wi][lane] + (s_c[wi][lane] + tb_ * 0.5));                                  // :228-229, :256
      }
      const double up = __shfl_up_sync(IVX_FULL_MASK, p, 1);
      const bool bad = (lane < n && lane > 0 && !(p_before == up)) || (ok && !(fabs(bp) < 1e290));
      if (!__any_sync(IVX_FULL_MASK, bad)) {
        if (lane < n) s_f[wi][lane] = p;
        __syncwarp();
        if (ok) {
          cp = s_f[wi][rank];
          prev = rank > 0 ? s_f[wi][rank - 1] : pc1;
          has_prev = rank > 0 || n_prev > 0;
        }
        if (n > 0) { pc1 = s_f[wi][n - 1]; n_prev = 1; }
      } else {                                             // sequential walk, lane after lane
        uint32_t mm = okm;
        while (mm) {
          const int src = __ffs(mm) - 1;
          mm &= mm - 1;
          const double tbc = n_prev > 0 ? (bp - pc1) * 0.3 : 0.0;
          const double cand = round_np<4>(bp + (co + tbc * 0.5));
          if (lane == src) { prev = pc1; has_prev = n_prev > 0; cp = cand; }
          pc1 = __shfl_sync(IVX_FULL_MASK, cand, src);
          n_prev = 1;
        }
      }
      if (!WRITE) { o_base += __popc(okm); continue; }
      const double tb = has_prev ? (bp - prev) * 0.3 : 0.0;
      op = bp + (oo + tb * 0.2);                                                // :225-226
      const double cu = bp + (co + tb * 0.5);                                   // the unrounded close
      const double mid = (op + cu) / 2.0;                                       // :232
      const double he = uni(0.0, spread / 2.0, uu[3]), lr = uni(0.0, spread / 2.0, uu[4]);
      hp = (cu > op ? cu : op) + he;                                            // :236  max(open, close)
      lp = (cu < op ? cu : op) - lr;                                            // :237
      if (hp - lp < bp * 0.0005) { hp = mid + bp * 0.00025; lp = mid - bp * 0.00025; }   // :240-242
    } else if (STRATEGY == ST_MIDPOINT) {
      const double spread = bp * 0.001, q = spread / 4.0;                       // :268-269
      op = bp + uni(-q, q, uu[0]);
      cp = round_np<4>(bp + uni(-q, q, uu[1]));
      hp = bp + spread / 2.0; lp = bp - spread / 2.0;
    } else if (STRATEGY == ST_SIMPLE) {
      const double spread = bp * 0.001, hs = spread / 2.0;                      // :345
      op = bp;
      cp = round_np<4>(bp + uni(-hs, hs, uu[0]));
      hp = bp + hs; lp = bp - hs;
    } else if (STRATEGY == ST_PIPELINE) {                                       // complete_pipeline.py:483-492
      const double spread = bp * 0.001, third = spread / 3.0;
      op = bp + uni(-third, third, uu[0]);
      const double cu = bp + uni(-third, third, uu[1]);
      hp = (cu > op ? cu : op) + fabs(uni(0.0, spread / 2.0, uu[2]));
      lp = (cu < op ? cu : op) - fabs(uni(0.0, spread / 2.0, uu[3]));
      cp = round_np<4>(cu);
    } else {                                                                    // trend_following :293-340
      const double zv = ok ? ld_stream(P.z + o) : 0.0;
      const double noise = 0.0 + (bp * 0.001) * zv;                             // np.random.normal(0, base * 0.001)
      uint32_t mm = okm;
      while (mm) {
        const int src = __ffs(mm) - 1;
        mm &= mm - 1;
        if (lane == src) {
          const int lookback = n_prev < 5 ? n_prev : 5;                         // :296
          double trend = 0.0;
          if (lookback > 1) {
            const double oldest = lookback == 2 ? pc2 : lookback == 3 ? pc3 : lookback == 4 ? pc4 : pc5;
            trend = (pc1 - oldest) / (double)lookback;                          // :302
          }
          const double ts = trend * 0.6;
          op = bp + ts + noise;                                                 // :308
          const double cu = bp + ts * 1.2 + noise;                              // :309
          const double at = fabs(trend);
          const double mx = cu > op ? cu : op, mn = cu < op ? cu : op;
          if (trend > 0.0) { hp = mx + at * 0.5; lp = mn - at * 0.2; }          // :312-317
          else { hp = mx + at * 0.2; lp = mn - at * 0.5; }
          cp = round_np<4>(cu);
        }
        pc5 = pc4; pc4 = pc3; pc3 = pc2; pc2 = pc1;
        pc1 = __shfl_sync(IVX_FULL_MASK, cp, src);
        ++n_prev;
      }
    }
    if (ok) {
      double v = nv ? 50.0 * ev : vin;                                          // :373-376
      v = v > 0.0 ? v : 0.0;                                                    // max(0, volume) :378
      const double ro = round_np<4>(op), rh = round_np<4>(hp), rl = round_np<4>(lp);
      if (!GEN || P.o_row) st_stream(P.o_row + o, r);
      if (GEN && P.o_ts) st_stream(P.o_ts + o, tsv);
      st_stream(P.o_open + o, ro); st_stream(P.o_high + o, rh); st_stream(P.o_low + o, rl); st_stream(P.o_close + o, cp);
      st_stream(P.o_volume + o, round_np<6>(v));
      st_stream(P.o_source + o, bp);
      // _validate_ohlcv_quality :385-405
      if (rh < rl) qflags |= Q_HIGH_LT_LOW;
      if (rh < ro || rh < cp) qflags |= Q_HIGH_LT_OC;
      if (rl > ro || rl > cp) qflags |= Q_LOW_GT_OC;
      if ((rh - rl) / bp > 0.1) qflags |= Q_SPREAD;
      if (ro <= 0.0 || rh <= 0.0 || rl <= 0.0 || cp <= 0.0) qflags |= Q_NONPOSITIVE;
    }
    o_base += __popc(okm);
    e_base += __popc(nvm);
  }
}

// ---- one warp per SYMBOL: exact by construction.  Used for trend_following (a five-close memory: its contraction is too
// slow for the segment warm-up below) and, with `only_if` set, as the stream-ordered fallback of the segmented kernel ----
template <int STRATEGY, bool GEN = false>
__global__ void __launch_bounds__(128) k_bridge_ohlcv(const BridgeParams P, const int32_t* __restrict__ only_if) {
  if (only_if && *only_if == 0) return;
  __shared__ BridgeSmem sm;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= P.n_sym) return;
  int64_t o_base = P.out_off[s], e_base = P.e_off[s];
  BridgeState st = fresh_state();
  uint32_t qflags = 0;
  bridge_piece<STRATEGY, true, GEN>(P, P.in_off[s], P.in_off[s + 1], o_base, e_base, st, qflags, sm);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) qflags |= __shfl_xor_sync(IVX_FULL_MASK, qflags, d);
  if (lane_id() == 0 && P.sym_quality) P.sym_quality[s] = qflags;
}

// ---- one warp per SEGMENT of SEG consecutive rows, whatever symbols they belong to: the parallelism no longer depends
// on how many symbols there are or how long they are (2 400 contracts x 28 801 rows gave 16 warps per SM).  The global
// candle index of a row is the number of emitting rows before it — symbol boundaries do not enter — so two scans over
// per-segment counts place every segment in the outputs and in the tapes.
// A segment that starts INSIDE a symbol needs the carried close of the row before it.  It is obtained by walking the
// close chain over the WARM_MIN (or, where few of them emit, WARM) rows before the segment from a fresh state: the chain forgets its start at 0.15 per emitted
// row (see bridge_piece), so after a few dozen emitted rows the walk is ON the true trajectory — and that is CHECKED,
// not assumed: every segment records the state it entered with and the state it left with, k_bridge_check compares each
// speculative entry with its predecessor's exit bit for bit (by induction from the segments that start a symbol or
// whose warm-up reaches the symbol's start, all states are then the sequential ones), and any mismatch raises a flag on
// which the per-symbol kernel above — enqueued right behind, idle otherwise — recomputes the batch exactly.
constexpr int SEG = 2048, WARM = 256, WARM_MIN = 64, WARM_EMIT = 32;
struct SegArrays {
  int64_t n_seg;
  const int64_t *seg_out, *seg_e;       // [n_seg + 1] candles / exponential draws before each segment
  double *entry_pc, *exit_pc;           // [n_seg]
  int32_t *entry_has, *exit_has, *spec; // [n_seg]
  int32_t* flag;
};

__global__ void __launch_bounds__(256) k_bridge_count_seg(int64_t n_rows, const double* __restrict__ base, const double* __restrict__ volume,
                                                          int64_t n_seg, int64_t* __restrict__ n_out, int64_t* __restrict__ n_e) {
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_seg) return;
  const int lane = lane_id();
  const int64_t a = g * SEG, b = min(n_rows, a + SEG);
  int c_ok = 0, c_e = 0;
  for (int64_t r = a + lane; r < b; r += 32) {
    const bool ok = row_ok(ld_stream(base + r));
    c_ok += ok ? 1 : 0;
    c_e += (ok && vol_missing(ld_stream(volume + r))) ? 1 : 0;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    c_ok += __shfl_xor_sync(IVX_FULL_MASK, c_ok, d);
    c_e += __shfl_xor_sync(IVX_FULL_MASK, c_e, d);
  }
  if (lane == 0) { n_out[g] = c_ok; n_e[g] = c_e; }
}

// out_off[s] / e_off[s] = candles / exponential draws before the first row of symbol s (s = n_sym: the totals): the
// segment prefix plus a count over at most SEG rows
__global__ void __launch_bounds__(256) k_bridge_sym_off(int64_t n_sym, int64_t n_rows, const int64_t* __restrict__ in_off,
                                                        const double* __restrict__ base, const double* __restrict__ volume,
                                                        int64_t n_seg, const int64_t* __restrict__ seg_out, const int64_t* __restrict__ seg_e,
                                                        int64_t* __restrict__ out_off, int64_t* __restrict__ e_off) {
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s > n_sym) return;
  const int lane = lane_id();
  const int64_t r = s < n_sym ? in_off[s] : n_rows;
  const int64_t g = min(r / SEG, n_seg);
  int c_ok = 0, c_e = 0;
  for (int64_t q = g * SEG + lane; q < r; q += 32) {
    const bool ok = row_ok(ld_stream(base + q));
    c_ok += ok ? 1 : 0;
    c_e += (ok && vol_missing(ld_stream(volume + q))) ? 1 : 0;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    c_ok += __shfl_xor_sync(IVX_FULL_MASK, c_ok, d);
    c_e += __shfl_xor_sync(IVX_FULL_MASK, c_e, d);
  }
  if (lane == 0) { out_off[s] = seg_out[g] + c_ok; e_off[s] = seg_e[g] + c_e; }
}

#ifndef IVX_BRIDGE_GEN_CTAS
#define IVX_BRIDGE_GEN_CTAS 6                          // 80 registers: 24 warps per SM (99 registers and 20 warps unbounded: 7 % slower)
#endif
template <int STRATEGY, bool GEN = false>
__global__ void __launch_bounds__(128, GEN ? IVX_BRIDGE_GEN_CTAS : 1) k_bridge_ohlcv_seg(const BridgeParams P, const SegArrays A, int64_t n_rows) {
  __shared__ BridgeSmem sm;
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= A.n_seg) return;
  const int lane = lane_id();
  const int64_t r0 = g * SEG, r1 = min(n_rows, r0 + SEG);
  int64_t o_base = A.seg_out[g], e_base = A.seg_e[g];
  // the symbol that owns row r0: the last s with in_off[s] <= r0 (empty symbols before it share that offset)
  int64_t sym = warp_upper_le<int64_t>(P.in_off, 0, P.n_sym, r0, lane);
  int64_t sym_start = __ldg(P.in_off + sym), sym_end = __ldg(P.in_off + sym + 1);
  BridgeState st = fresh_state();
  uint32_t dummy_flags = 0;
  int32_t spec = 0;
  if (STRATEGY == ST_SPREAD && r0 > sym_start) {
    // WARM_MIN rows back when those hold WARM_EMIT emitting rows (the chain forgets its start at 0.15 per EMITTED row:
    // 0.15^32 = 4e-27 of the first guess's error), WARM rows back on a stretch full of rows that emit nothing
    int64_t ws = max(sym_start, r0 - WARM_MIN);
    int cnt = 0;                                       // emitting rows in [ws, r0): where the warm-up's uniforms start
    for (int64_t r = ws + lane; r < r0; r += 32) cnt += row_ok(ld_stream(P.base + r)) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(IVX_FULL_MASK, cnt, d);
    if (cnt < WARM_EMIT && ws > sym_start) {
      ws = max(sym_start, r0 - WARM);
      cnt = 0;
      for (int64_t r = ws + lane; r < r0; r += 32) cnt += row_ok(ld_stream(P.base + r)) ? 1 : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(IVX_FULL_MASK, cnt, d);
    }
    spec = ws > sym_start ? 1 : 0;
    int64_t o_w = o_base - cnt, e_w = 0;
    bridge_piece<STRATEGY, false, GEN>(P, ws, r0, o_w, e_w, st, dummy_flags, sm);
  }
  if (lane == 0) { A.entry_pc[g] = st.pc1; A.entry_has[g] = st.n_prev > 0; A.spec[g] = spec; }
  // the segment's rows, symbol piece after symbol piece
  int64_t a = r0;
  while (a < r1) {
    const int64_t b = min(r1, sym_end);
    if (b > a) {
      uint32_t qflags = 0;
      bridge_piece<STRATEGY, true, GEN>(P, a, b, o_base, e_base, st, qflags, sm);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) qflags |= __shfl_xor_sync(IVX_FULL_MASK, qflags, d);
      if (lane == 0 && P.sym_quality && qflags) atomicOr(P.sym_quality + sym, qflags);
      a = b;
    }
    if (a < r1) {                                      // next symbol: fresh state
      if (sym + 1 >= P.n_sym) break;                   // rows beyond in_off[n_sym] (inconsistent arguments): not ours
      ++sym;
      sym_start = sym_end; sym_end = __ldg(P.in_off + sym + 1);
      st = fresh_state();
    }
  }
  if (lane == 0) { A.exit_pc[g] = st.pc1; A.exit_has[g] = st.n_prev > 0; }
}

// _validate_ohlcv_quality (:371-402) on candles that are already there (a frame the caller hands in): flag bits OR-ed
__global__ void __launch_bounds__(256) k_bridge_quality(int64_t n, const double* __restrict__ o, const double* __restrict__ h,
                                                        const double* __restrict__ l, const double* __restrict__ c,
                                                        const double* __restrict__ src, uint32_t* __restrict__ flags) {
  uint32_t f = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ro = o[i], rh = h[i], rl = l[i], cp = c[i], bp = src[i];
    if (rh < rl) f |= Q_HIGH_LT_LOW;
    if (rh < ro || rh < cp) f |= Q_HIGH_LT_OC;
    if (rl > ro || rl > cp) f |= Q_LOW_GT_OC;
    if ((rh - rl) / bp > 0.1) f |= Q_SPREAD;
    if (ro <= 0.0 || rh <= 0.0 || rl <= 0.0 || cp <= 0.0) f |= Q_NONPOSITIVE;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) f |= __shfl_xor_sync(IVX_FULL_MASK, f, d);
  if (lane_id() == 0 && f) atomicOr(flags, f);
}

__global__ void __launch_bounds__(256) k_bridge_check(const SegArrays A, int32_t force) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (force && g == 0) atomicOr(A.flag, 1);
  if (g < 1 || g >= A.n_seg || !A.spec[g]) return;
  const bool same = (__double_as_longlong(A.entry_pc[g]) == __double_as_longlong(A.exit_pc[g - 1])) && (A.entry_has[g] == A.exit_has[g - 1]);
  if (!same) atomicOr(A.flag, 1);
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_bridge_n_segments(int64_t n_rows) { return n_rows < 0 ? -1 : (n_rows + SEG - 1) / SEG; }

extern "C" int64_t ivx_bridge_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return -1;
  const int64_t g = (n_rows + SEG - 1) / SEG;
  return 8 * 2 * g + 4 * (3 * g + 1) + 8 * 256;            // the larger of the plan's (two count arrays) and the candle call's (entry/exit states, spec, flag) needs
}

extern "C" int ivx_bridge_plan(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                               int64_t* seg_out, int64_t* seg_e, int64_t* out_off, int64_t* e_off, void* workspace,
                               int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_rows >= 0, IVX_ERR_INVALID_ARG, "ivx_bridge_plan: negative size");
  IVX_REQUIRE(in_off && seg_out && seg_e && out_off && e_off && workspace && (n_rows == 0 || (base && volume)), IVX_ERR_INVALID_ARG,
              "ivx_bridge_plan: null pointer");
  IVX_REQUIRE(workspace_bytes >= ivx_bridge_workspace_bytes(n_rows), IVX_ERR_INVALID_ARG, "ivx_bridge_plan: workspace too small");
  const int64_t g = (n_rows + SEG - 1) / SEG;
  Carver cv(workspace, workspace_bytes);
  int64_t* n_out_seg = cv.take<int64_t>(g);
  int64_t* n_e_seg = cv.take<int64_t>(g);
  if (g > 0) {
    k_bridge_count_seg<<<(unsigned)((g + 7) / 8), 256, 0, stream>>>(n_rows, base, volume, g, n_out_seg, n_e_seg);
    IVX_LAUNCH_CHECK("k_bridge_count_seg");
  }
  int rc = ivx_exclusive_scan_i64(g, n_out_seg, seg_out, stream_);
  if (rc != IVX_OK) return rc;
  rc = ivx_exclusive_scan_i64(g, n_e_seg, seg_e, stream_);
  if (rc != IVX_OK) return rc;
  k_bridge_sym_off<<<(unsigned)((n_sym + 1 + 7) / 8), 256, 0, stream>>>(n_sym, n_rows, in_off, base, volume, g, seg_out, seg_e, out_off, e_off);
  IVX_LAUNCH_CHECK("k_bridge_sym_off");
  return IVX_OK;
}

// the launches behind ivx_bridge_ohlcv (tapes in memory) and ivx_bridge_ohlcv_philox (GEN: tapes generated in the kernel)
template <bool GEN>
static int bridge_launch(const BridgeParams& P, int64_t n_rows, int32_t force_fallback, const int64_t* seg_out, const int64_t* seg_e,
                         void* workspace, int64_t workspace_bytes, cudaStream_t stream, const char* who) {
  const int strategy = P.strategy;
  const int64_t n_sym = P.n_sym;
  const unsigned blocks = (unsigned)((n_sym + 3) / 4);
  if (strategy == ST_TREND || workspace == nullptr || seg_out == nullptr || seg_e == nullptr) {   // one warp per symbol
    switch (strategy) {
      case ST_SPREAD: k_bridge_ohlcv<ST_SPREAD, GEN><<<blocks, 128, 0, stream>>>(P, nullptr); break;
      case ST_MIDPOINT: k_bridge_ohlcv<ST_MIDPOINT, GEN><<<blocks, 128, 0, stream>>>(P, nullptr); break;
      case ST_TREND: k_bridge_ohlcv<ST_TREND, false><<<blocks, 128, 0, stream>>>(P, nullptr); break;
      case ST_SIMPLE: k_bridge_ohlcv<ST_SIMPLE, GEN><<<blocks, 128, 0, stream>>>(P, nullptr); break;
      default: k_bridge_ohlcv<ST_PIPELINE, GEN><<<blocks, 128, 0, stream>>>(P, nullptr); break;
    }
    IVX_LAUNCH_CHECK("k_bridge_ohlcv");
    return IVX_OK;
  }
  // one warp per segment of SEG rows
  IVX_REQUIRE(workspace_bytes >= ivx_bridge_workspace_bytes(n_rows), IVX_ERR_INVALID_ARG, "%s: workspace too small", who);
  const int64_t g = (n_rows + SEG - 1) / SEG;
  Carver cv(workspace, workspace_bytes);
  SegArrays A;
  A.n_seg = g; A.seg_out = seg_out; A.seg_e = seg_e;
  A.entry_pc = cv.take<double>(g); A.exit_pc = cv.take<double>(g);
  A.entry_has = cv.take<int32_t>(g); A.exit_has = cv.take<int32_t>(g); A.spec = cv.take<int32_t>(g);
  A.flag = cv.take<int32_t>(1);
  IVX_REQUIRE(cv.ok(), IVX_ERR_INVALID_ARG, "%s: workspace too small", who);
  IVX_CUDA(cudaMemsetAsync(A.flag, 0, sizeof(int32_t), stream));
  if (P.sym_quality) IVX_CUDA(cudaMemsetAsync(P.sym_quality, 0, (size_t)n_sym * sizeof(uint32_t), stream));
  const unsigned sblocks = (unsigned)((g + 3) / 4);
  switch (strategy) {
    case ST_SPREAD: k_bridge_ohlcv_seg<ST_SPREAD, GEN><<<sblocks, 128, 0, stream>>>(P, A, n_rows); break;
    case ST_MIDPOINT: k_bridge_ohlcv_seg<ST_MIDPOINT, GEN><<<sblocks, 128, 0, stream>>>(P, A, n_rows); break;
    case ST_SIMPLE: k_bridge_ohlcv_seg<ST_SIMPLE, GEN><<<sblocks, 128, 0, stream>>>(P, A, n_rows); break;
    default: k_bridge_ohlcv_seg<ST_PIPELINE, GEN><<<sblocks, 128, 0, stream>>>(P, A, n_rows); break;
  }
  IVX_LAUNCH_CHECK("k_bridge_ohlcv_seg");
  if (strategy == ST_SPREAD) {
    // the carried closes the segments assumed are checked against the ones their predecessors produced; on a mismatch
    // (never seen; ~1e-9 per boundary by the contraction argument) the per-symbol kernel recomputes the batch, stream-ordered
    k_bridge_check<<<(unsigned)((g + 255) / 256), 256, 0, stream>>>(A, force_fallback);
    IVX_LAUNCH_CHECK("k_bridge_check");
    k_bridge_ohlcv<ST_SPREAD, GEN><<<blocks, 128, 0, stream>>>(P, A.flag);
    IVX_LAUNCH_CHECK("k_bridge_ohlcv(fallback)");
  }
  return IVX_OK;
}

extern "C" int ivx_bridge_ohlcv(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                                int32_t strategy, double base_spread_percent, double volatility_factor, const double* u,
                                const double* e, const double* z, const int64_t* out_off, const int64_t* e_off,
                                const int64_t* seg_out, const int64_t* seg_e,
                                int64_t* o_row, double* o_open, double* o_high, double* o_low, double* o_close,
                                double* o_volume, double* o_source, uint32_t* sym_quality, void* workspace,
                                int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_rows >= 0, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv: negative size");
  // IVX_BRIDGE_RECOMPUTE (a flag or-ed into strategy): treat the segment check as failed, i.e. let the stream-ordered
  // per-symbol kernel recompute the batch after the segment kernel — a diagnostic that exercises the fallback (no global state)
  const int32_t force_fallback = (strategy & IVX_BRIDGE_RECOMPUTE) ? 1 : 0;
  strategy &= ~IVX_BRIDGE_RECOMPUTE;
  IVX_REQUIRE(strategy >= ST_SPREAD && strategy <= ST_PIPELINE, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv: unknown strategy %d", strategy);
  if (n_sym == 0 || n_rows == 0) return IVX_OK;
  IVX_REQUIRE(in_off && base && volume && out_off && e_off && o_row && o_open && o_high && o_low && o_close && o_volume && o_source,
              IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv: null pointer");
  IVX_REQUIRE(strategy == ST_TREND ? z != nullptr : u != nullptr, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv: the strategy's tape is missing");
  IVX_REQUIRE(e != nullptr, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv: the exponential tape is missing (pass any valid pointer when it is empty)");
  BridgeParams P{n_sym, in_off, base, volume, strategy, base_spread_percent, volatility_factor, u, e, z, out_off, e_off,
                 o_row, o_open, o_high, o_low, o_close, o_volume, o_source, sym_quality, 0, 0, 0, nullptr, nullptr};
  return bridge_launch<false>(P, n_rows, force_fallback, seg_out, seg_e, workspace, workspace_bytes, stream, "ivx_bridge_ohlcv");
}

extern "C" int ivx_bridge_ohlcv_philox(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                                       const int64_t* ts, int32_t strategy, double base_spread_percent, double volatility_factor,
                                       uint64_t seed, uint32_t stream_u, uint32_t stream_e, const int64_t* out_off,
                                       const int64_t* e_off, const int64_t* seg_out, const int64_t* seg_e, int64_t* o_row,
                                       int64_t* o_ts, double* o_open, double* o_high, double* o_low, double* o_close,
                                       double* o_volume, double* o_source, uint32_t* sym_quality, void* workspace,
                                       int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_rows >= 0, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv_philox: negative size");
  const int32_t force_fallback = (strategy & IVX_BRIDGE_RECOMPUTE) ? 1 : 0;
  strategy &= ~IVX_BRIDGE_RECOMPUTE;
  IVX_REQUIRE(strategy >= ST_SPREAD && strategy <= ST_PIPELINE && strategy != ST_TREND, IVX_ERR_INVALID_ARG,
              "ivx_bridge_ohlcv_philox: strategy %d (trend_following draws normals: pass a tape to ivx_bridge_ohlcv)", strategy);
  if (n_sym == 0 || n_rows == 0) return IVX_OK;
  IVX_REQUIRE(in_off && base && volume && out_off && e_off && o_open && o_high && o_low && o_close && o_volume && o_source,
              IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv_philox: null pointer");
  IVX_REQUIRE((o_ts == nullptr) || (ts != nullptr), IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv_philox: o_ts without ts");
  IVX_REQUIRE(o_row || o_ts, IVX_ERR_INVALID_ARG, "ivx_bridge_ohlcv_philox: neither o_row nor o_ts: the candles would have no key");
  BridgeParams P{n_sym, in_off, base, volume, strategy, base_spread_percent, volatility_factor, nullptr, nullptr, nullptr, out_off, e_off,
                 o_row, o_open, o_high, o_low, o_close, o_volume, o_source, sym_quality, seed, stream_u, stream_e, ts, o_ts};
  return bridge_launch<true>(P, n_rows, force_fallback, seg_out, seg_e, workspace, workspace_bytes, stream, "ivx_bridge_ohlcv_philox");
}

extern "C" int ivx_bridge_quality(int64_t n, const double* open, const double* high, const double* low, const double* close,
                                  const double* source_price, uint32_t* flags, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n >= 0 && flags, IVX_ERR_INVALID_ARG, "ivx_bridge_quality: bad arguments");
  IVX_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t), stream));
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(open && high && low && close && source_price, IVX_ERR_INVALID_ARG, "ivx_bridge_quality: null pointer");
  const int64_t blocks = (n + 255) / 256;
  k_bridge_quality<<<(unsigned)(blocks < 4 * 148 ? blocks : 4 * 148), 256, 0, stream>>>(n, open, high, low, close, source_price, flags);
  IVX_LAUNCH_CHECK("k_bridge_quality");
  return IVX_OK;
}
