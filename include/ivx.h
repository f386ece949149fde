/* ivx.h — C ABI of libivx.so, the B200 (sm_100a) implementation of IV_INTERPOLATION's hot path.
 *
 * The reference (liu-wei2021/IV_INTERPOLATION) is pure Python and has no FFI; the seam this
 * library sits behind is the pair of classes the reference's batch drivers construct
 * (reference src/interpolation/batch_processor.py:21-24,97 and
 * src/candle_reconstruction/core.py:180,229-248).  Each entry point names the reference code
 * it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative IVX_ERR_* code; nothing throws;
 *     ivx_last_error_string() describes the last failure on the calling thread;
 *   - all array arguments are DEVICE pointers owned by the caller (e.g. torch tensors'
 *     data_ptr()); the library allocates nothing and synchronises nothing: work is enqueued
 *     on `stream` (a cudaStream_t, may be NULL for the legacy default stream);
 *   - scratch memory is passed in explicitly (`workspace`), sized by the *_workspace_bytes call;
 *   - timestamps are int64 ticks in the caller's datetime64 unit; NaT is INT64_MIN and is ignored;
 *   - rows of one symbol are contiguous (CSR offsets, int64) and sorted by timestamp ascending.
 */
#ifndef IVX_H_
#define IVX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVX_VERSION 100

/* error codes */
#define IVX_OK 0
#define IVX_ERR_INVALID_ARG (-1)
#define IVX_ERR_CUDA (-2)
#define IVX_ERR_TOO_LARGE (-3)
#define IVX_ERR_UNSUPPORTED (-4)

/* per-symbol status (uint8): why the reference would have returned None */
#define IVX_SYM_OK 0
#define IVX_SYM_TOO_FEW 1    /* interpolation/core.py:26-28, candle_reconstruction/core.py:47-49,64-66 */
#define IVX_SYM_SPAN 2       /* interpolation/core.py:36-39: max-min > 30 days */
#define IVX_SYM_TOO_LONG 3   /* interpolation/core.py:49-51: > 100000 grid points */
#define IVX_SYM_EMPTY_OUT 4  /* interpolation/core.py:76-78: nothing left after dropna */
#define IVX_SYM_NO_DATES 5   /* every date is NaT: date_range raises -> core.py:83-85 */

/* per-symbol flag bits (uint32) */
#define IVX_FLAG_IN_BAD_OHLC 1u     /* candle core.py:115-124 on the input rows */
#define IVX_FLAG_IN_NEG_VOLUME 2u   /* candle core.py:127-130 */
#define IVX_FLAG_IN_NAN 4u          /* candle core.py:133-137 */
#define IVX_FLAG_OUT_BAD_OHLC 8u    /* same three checks on the emitted candles (core.py:239) */
#define IVX_FLAG_OUT_NEG_VOLUME 16u
#define IVX_FLAG_OUT_NAN 32u
#define IVX_FLAG_UNSORTED 64u       /* timestamps decrease inside the symbol: result undefined */
#define IVX_FLAG_NONFINITE 128u     /* +-inf among the interpolated columns: dropna may cut interior rows */

#define IVX_MAX_FILL_COLS 8
#define IVX_MAX_IDX_COLS 4

typedef void* ivx_stream_t; /* cudaStream_t */

int ivx_version(void);
const char* ivx_last_error_string(void);
/* name of every exported kernel family compiled into the library, comma separated */
const char* ivx_build_info(void);

/* ---------------------------------------------------------------------------------------------
 * Kernel B — OHLCV bucket reduction.
 * Replaces CandleReconstructor.reconstruct_symbol_candles (candle_reconstruction/core.py:59-89)
 * and CompleteOptimizedPipeline._aggregate_to_5min (complete_pipeline.py:685-710), for a whole
 * CSR batch of symbols in one launch, plus validate_candle_data (core.py:108-144) as flag bits.
 *
 *   bucket      = floor(ts / bucket_ticks)                      (epoch anchored, core.py:70)
 *   open/close  = first/last non-NaN in row order; high/low = NaN-skipping max/min
 *   volume      = Kahan-compensated sum in row order, NaN skipped, all-NaN -> 0.0 (pandas group_sum)
 *   a bucket is emitted iff its raw row count >= min_count       (core.py:87-89)
 *   check_len != 0: a symbol with fewer than min_count rows emits nothing and gets
 *                   IVX_SYM_TOO_FEW (core.py:64-66); empty symbols always do (core.py:47-49).
 *
 * Outputs are compacted, symbol-major, time order.  capacity = length of the o_* arrays;
 * n_rows is an upper bound on the candle count.  o_sym_row / o_group / sym_group_off /
 * sym_flags / sym_valid may be NULL.
 *   out_off[n_sym+1]    CSR offsets of the emitted candles
 *   o_sym_row[i]        global row index of the first row with a non-null symbol in the bucket
 *                       (-1 if none) — the 'symbol':'first' aggregate, core.py:79
 *   o_group[i]          rank of the bucket among ALL buckets of the batch; minus
 *                       sym_group_off[s] it is the (not reset) index pandas returns
 *   total[0]            number of candles written (device int64)
 * ------------------------------------------------------------------------------------------- */
int64_t ivx_ohlcv_workspace_bytes(int64_t n_sym, int64_t n_rows);
int ivx_ohlcv_reduce(int64_t n_sym, const int64_t* in_off, int64_t n_rows, const int64_t* ts,
                     const double* open, const double* high, const double* low, const double* close,
                     const double* volume, const uint8_t* sym_valid, int64_t bucket_ticks,
                     int32_t min_count, int32_t check_len, int64_t capacity, int64_t* out_off,
                     int64_t* o_ts, double* o_open, double* o_high, double* o_low, double* o_close,
                     double* o_volume, int32_t* o_count, int64_t* o_sym_row, int64_t* o_group,
                     int64_t* sym_group_off, uint8_t* sym_status, uint32_t* sym_flags,
                     int64_t* total, void* workspace, int64_t workspace_bytes, ivx_stream_t stream);

/* Time-major variant for the surface pipeline: x[t*n_series + s], all series share the stamp
 * ts[t]; one candle per (bucket, series).  Same aggregation rules with value=open=high=low=close
 * source (the interpolated IV) and no volume: o_*[b*n_series + s].  bucket_of_t[t] (int32,
 * non-decreasing, dense from 0) and n_buckets come from ivx_time_buckets.  Buckets with fewer
 * than min_count stamps are written as NaN with o_count = their count. */
int ivx_time_buckets(int64_t n_t, const int64_t* ts, int64_t bucket_ticks, int32_t* bucket_of_t,
                     int64_t* bucket_ts, int32_t* bucket_start, int64_t* n_buckets, ivx_stream_t stream);
int ivx_ohlc_time_major(int64_t n_t, int64_t n_series, const double* x, int64_t n_buckets,
                        const int32_t* bucket_start, int32_t min_count, double* o_open, double* o_high,
                        double* o_low, double* o_close, int32_t* o_count, ivx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel A — per-contract time-linear interpolation + forward fill.
 * Replaces IVInterpolator.interpolate_symbol (interpolation/core.py:26-81) for a CSR batch.
 *
 * Step 1, ivx_interp_plan: guards (core.py:26-28,36-39,49-51), the minute grid anchored at
 *   min(date) (core.py:42-46), the left join (core.py:54-55; off-grid rows vanish, duplicate
 *   stamps widen the position axis), per-column bracketing knots and forward-fill sources,
 *   and the dropna cut (core.py:74).  Writes sym_status, sym_flags and out_count[n_sym]
 *   (rows each contract will emit) and out_off[n_sym+1] (their exclusive scan; out_off[n_sym]
 *   is the total, a device int64 the caller reads to size the outputs).
 * Step 2, ivx_interp_exec: writes the output rows, contract-major:
 *   o_ts       int64   grid stamp
 *   o_interp   3 x f64 iv, underlying_price, time_to_maturity:  y = slope*(p-p0) + y0 evaluated
 *              as a separate multiply and add, with np.interp's NaN fallbacks (bit-equal to pandas)
 *   o_fill     n_fill x f64: last non-NaN observation carried forward, column by column
 *   o_idx      n_idx x i32: global source-row index carried forward for string columns whose
 *              non-null mask is idx_valid[k] (NULL mask = all valid); -1 = none yet
 *   o_src_row  i32: global source row sitting exactly on this grid point, else -1
 *   o_pos      i32: position on the contract's merged axis (the index pandas leaves after dropna)
 *   o_keep     u8 : 1 where dropna keeps the row (all 1 unless IVX_FLAG_NONFINITE)
 *   Any of o_fill/o_idx/o_src_row/o_pos/o_keep/o_is_interpolated may be NULL.
 *   o_is_interpolated (u8) is written as 0: core.py:71 tests `symbol` after the forward fill.
 * idx_valid[0] is the `symbol` column's mask and takes part in dropna (core.py:74).
 * ------------------------------------------------------------------------------------------- */
int64_t ivx_interp_workspace_bytes(int64_t n_sym, int64_t n_src, int32_t n_fill, int32_t n_idx);
int ivx_interp_plan(int64_t n_sym, const int64_t* src_off, int64_t n_src, const int64_t* src_ts,
                    const double* const* interp_cols /*3, host array of device ptrs*/,
                    int32_t n_fill, const double* const* fill_cols,
                    int32_t n_idx, const uint8_t* const* idx_valid,
                    int64_t step_ticks, int64_t max_span_ticks, int64_t max_grid_points,
                    int32_t min_points, int64_t* out_count, int64_t* out_off, uint8_t* sym_status,
                    uint32_t* sym_flags, void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
int ivx_interp_exec(int64_t n_sym, const int64_t* src_off, int64_t n_src, const int64_t* out_off,
                    int64_t n_out, int64_t step_ticks, int32_t n_fill, int32_t n_idx,
                    int64_t* o_ts, double* const* o_interp /*3*/, double* const* o_fill,
                    int32_t* const* o_idx, int32_t* o_src_row, int32_t* o_pos, uint8_t* o_keep,
                    uint8_t* o_is_interpolated, const void* workspace, int64_t workspace_bytes,
                    ivx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel C — Black-Scholes implied volatility, one thread per quote (north_star kernel 1; the
 * reference has no solver, its conventions come from interpolation/greeks.py:21-22: continuous
 * rate r, no dividend, T in years).  sigma[i] = NaN outside the invertible domain
 * (non-finite/non-positive inputs, or price not strictly inside the no-arbitrage bounds; S, K, T
 * and K e^{-rT} must be NORMAL doubles below 1e300 — subnormal or larger values also give NaN).
 * iters[i] = fp64 price evaluations spent on quote i (1 on the fast path); may be NULL.  ivx_bs_price is the forward pricer (used to build synthetic tapes),
 * ivx_bs_greeks the Greeks of greeks.py:24-35 (put rho keeps the reference's sign when
 * reference_put_rho_sign != 0).
 * ------------------------------------------------------------------------------------------- */
/* With a workspace (ivx_bs_implied_vol_workspace_bytes(n); 4 B per 32 quotes) the solve is two launches: a fast kernel
 * with no fallback code in it records the warps that hold a quote its straight-line path does not accept, a retry kernel
 * re-solves those warps with the bracketed loop.  workspace == NULL: one launch, fallback in line.  Same results. */
int64_t ivx_bs_implied_vol_workspace_bytes(int64_t n);
int ivx_bs_implied_vol(int64_t n, const double* price, const double* S, const double* K,
                       const double* T, const double* r, const uint8_t* is_call, double* sigma,
                       int32_t* iters, void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
/* The same solve for quotes that form slices on PRESUMED ladders (the per-quote layout of the surface pipeline): quote i = t * period + j,
 * slice = i / n_strike, and the strike of quote (t, j) is presumed to equal x_ladder[j] (period doubles: typically the first stamp's
 * strikes).  The solver has every strike in a register anyway, so it compares it with the ladder's bit for bit on the way: a slice with
 * a differing strike gets slice_flags[slice] = 1 (the array, n_t * period / n_strike int32, and *flag_count are zeroed first;
 * *flag_count = number of flagged slices).  ivx_spline_fit_eval_banded(..., x_verified = 1) then takes the flags as the verdict and
 * does not read 8 bytes per knot again.  Two-kernel form only (workspace required), no iteration counts. */
int64_t ivx_bs_implied_vol_ladder_workspace_bytes(int64_t n_t, int64_t period);
int ivx_bs_implied_vol_ladder(int64_t n_t, int64_t period, int64_t n_strike, const double* price, const double* S, const double* K,
                              const double* T, const double* r, const uint8_t* is_call, double* sigma, const double* x_ladder,
                              int32_t* slice_flags, uint32_t* flag_count, void* workspace, int64_t workspace_bytes, ivx_stream_t stream);

/* Chain layout (one underlying, one day): quote (t, e, k) sits at price[(t*n_exp + e)*n_strike + k]; the
 * underlying S_t[n_t] and rate r_t[n_t] are per stamp, the strike K_ek[n_exp*n_strike] per contract, the time to
 * expiry T_te[n_t*n_exp] per (stamp, expiry), call/put per contract or per quote: 8-9 B/quote in, 8 B out
 * instead of 41 + 8.  ln S, ln K, rT and
 * e^{-rT} are formed once per stamp / contract / (stamp, expiry) in `workspace`, which also holds the retry list
 * (the chain solve is always fast kernel + retry kernel). */
int64_t ivx_bs_chain_workspace_bytes(int64_t n_t, int64_t n_exp, int64_t n_strike);
int ivx_bs_implied_vol_chain(int64_t n_t, int64_t n_exp, int64_t n_strike, const double* price,
                             const uint8_t* is_call /* [t][e][k] if is_call_per_quote, else [e][k] */,
                             int32_t is_call_per_quote, const double* S_t, const double* K_ek,
                             const double* T_te, const double* r_t, double* sigma, int32_t* iters,
                             void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
/* The same with the price column as int32 ticks of 1e-4 — the precision the reference stores prices at (DECIMAL(15,4),
 * src/database/schema.py:30-45) — INT32_MIN = no quote: 4 B per quote in instead of 8.  ticks / 1e4 is formed correctly
 * rounded, so the result is bit-equal to ivx_bs_implied_vol_chain on prices rounded to four decimals. */
int ivx_bs_implied_vol_chain_ticks(int64_t n_t, int64_t n_exp, int64_t n_strike, const int32_t* price_ticks,
                                   const uint8_t* is_call, int32_t is_call_per_quote, const double* S_t, const double* K_ek,
                                   const double* T_te, const double* r_t, double* sigma, int32_t* iters,
                                   void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
int ivx_bs_price(int64_t n, const double* S, const double* K, const double* T, const double* r,
                 const double* sigma, const uint8_t* is_call, double* price, ivx_stream_t stream);
int ivx_bs_greeks(int64_t n, const double* S, const double* K, const double* T, const double* r,
                  const double* sigma, const uint8_t* is_call, int32_t reference_put_rho_sign,
                  double* delta, double* gamma, double* theta, double* vega, double* rho,
                  ivx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel D — batched natural cubic spline fit + evaluation (north_star kernel 2; no reference
 * code).  Slice i has knots x/y[knot_off[i]:knot_off[i+1]] (x strictly increasing; knots whose
 * x or y is not finite are skipped) and is evaluated at the n_q shared abscissae xq (ascending):
 * yq[i*n_q + j].  extrap: 0 NaN outside the knots, 1 continue the end cubic, 2 clamp flat.
 * Fewer than 2 usable knots -> NaN row; exactly 2 -> the straight line.
 * ivx_spline_strided solves many small systems laid out [knot][system] (system index
 * contiguous): used for the maturity direction of the surface.
 * ------------------------------------------------------------------------------------------- */
/* The solve keeps its whole state in shared memory and registers: no workspace is needed (the query returns a token 256 and the
 * workspace arguments below are ignored; they stay in the signatures so that callers written against round 1 keep linking). */
int64_t ivx_spline_workspace_bytes(int64_t n_slices, int64_t n_knots_total);
int ivx_spline_fit_eval(int64_t n_slices, const int64_t* knot_off, int64_t max_knots,
                        const double* x, const double* y, int64_t n_q, const double* xq,
                        int32_t extrap, double* yq, void* workspace, int64_t workspace_bytes,
                        ivx_stream_t stream);
/* Same, with the abscissae of slice i at x[x_off[i] + j] (x_off NULL = knot_off): many slices may share one
 * strike ladder.  Needs no workspace. */
int ivx_spline_fit_eval_xoff(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, int64_t max_knots,
                             const double* x, const double* y, int64_t n_q, const double* xq,
                             int32_t extrap, double* yq, ivx_stream_t stream);
/* Shared ladders: when many slices use the same abscissae (a chain's strike ladder per expiry), the x-only half of
 * the elimination is done once per ladder.  ivx_spline_ladder_factor fills three tables of
 * ivx_spline_ladder_table_len(n_ladders, max_knots) doubles for the ladders x[ladder_off[l] : ladder_off[l+1]) (stored in
 * the order the evaluation kernel reads them; use the same max_knots in both calls).  ivx_spline_fit_eval_ladder then only
 * eliminates each slice's right-hand side: slice i uses ladder slice_ladder[i], whose abscissae start at x[x_off[i]]
 * (= x[ladder_off[slice_ladder[i]]]); its knots are the ladder or a prefix of it.  Slices with unusable knots fall back
 * to the general solve inside the same launch; results agree with ivx_spline_fit_eval to rounding. */
int64_t ivx_spline_ladder_table_len(int64_t n_ladders, int64_t max_knots);
int ivx_spline_ladder_factor(int64_t n_ladders, const int64_t* ladder_off, int64_t max_knots, const double* x,
                             double* lu_cp, double* lu_inv, double* lu_ih, ivx_stream_t stream);
int ivx_spline_fit_eval_ladder(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off,
                               const int32_t* slice_ladder, int64_t max_knots, const double* x,
                               const double* lu_cp, const double* lu_inv, const double* lu_ih, const double* y,
                               int64_t n_q, const double* xq, int32_t extrap, double* yq, ivx_stream_t stream);
/* The same when the caller does NOT know whether its slices share ladders: slice i brings its own abscissae
 * x_own[knot_off[i] + j] and is only presumed to sit on ladder slice_ladder[i] (abscissae x[x_off[i] + j], length
 * ladder_off[l+1] - ladder_off[l]).  The kernel compares the two bit for bit while staging and uses the tabulated
 * factors only where every knot matches; every other slice takes the general solve inside the same launch. */
int ivx_spline_fit_eval_ladder_verify(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off,
                                      const int32_t* slice_ladder, const int64_t* ladder_off, int64_t max_knots,
                                      const double* x, const double* x_own, const double* lu_cp, const double* lu_inv,
                                      const double* lu_ih, const double* y, int64_t n_q, const double* xq,
                                      int32_t extrap, double* yq, ivx_stream_t stream);
/* Shared ladders as a BANDED operator (ivx_spline_band.cu).  For a fixed ladder and a fixed grid, fit + evaluate is a fixed
 * linear map y -> yq whose rows decay geometrically away from the grid point's bracket: ivx_spline_band_build tabulates,
 * per (ladder, grid point), the 48 knot weights that matter at 1e-12 and where they start (band_w: ivx_spline_band_len
 * doubles, band_k0: n_ladders*n_q int32, -1 = the point evaluates to NaN), and CHECKS the band against a full solve of
 * two probe vectors — ladder_ok[l] = 0 (fewer than 48 or more than 2048 knots, not increasing, band misses 2e-12)
 * sends that ladder's slices to the exact solve.  ivx_spline_fit_eval_banded evaluates every slice as n_q (<= 512) dot
 * products out of shared memory; slices with a non-finite y, another knot count than their ladder, or (x_own != NULL) abscissae
 * that differ from the ladder's in any bit are flagged in fb_list (int32[n_slices]: 1 = solved by the exact scan, 0 = by the
 * band; their number in *fb_count) and solved by the exact scan kernel right behind, in the same stream and in slice order,
 * with the ladder tables lu_* of ivx_spline_ladder_factor.  Slices that belong to no ladder group are not touched.
 * order / lad_slice_off: the slices grouped by ladder (CSR); max_slices_per_ladder sizes the grid. */
int64_t ivx_spline_band_len(int64_t n_ladders, int64_t n_q);
int64_t ivx_spline_band_frag_len(int64_t n_ladders);   /* doubles of band_frag */
int64_t ivx_spline_band_oct_len(int64_t n_ladders);    /* int32 of band_oct */
/* band_frag / band_oct (both or neither; NULL = dot-product form only): the same weights laid out as fp64 mma B fragments per group of
 * 8 grid points, for the tensor-path form of the evaluation (the operator applied to 8 slices at a time as a block-banded matrix
 * product, mma.sync m8n8k4).  ladder_ok[l] bit 0 = band usable, bit 1 = tensor-path tables usable (every 8-point group spans at
 * most 96 knots, every 64-point slab at most 512).  band_oct ends with {largest slab extent, largest step count} per ladder:
 * the largest extent over the ladders is the slab_knots_hint of ivx_spline_fit_eval_banded (0 = unknown: 512 is assumed). */
int ivx_spline_band_build(int64_t n_ladders, const int64_t* ladder_off, const double* x_ladder, int64_t n_q, const double* xq,
                          int32_t extrap, int32_t* band_k0, double* band_w, uint8_t* ladder_ok,
                          int16_t* band_perm /* n_ladders*512: which grid point each thread of the evaluation takes */,
                          double* band_frag, int32_t* band_oct,
                          double* band_fill_w /* ivx_spline_band_fill_len(knots of all ladders) doubles, or NULL */,
                          int32_t* band_fill_k0 /* one per ladder knot, or NULL */, ivx_stream_t stream);
/* band_fill_w / band_fill_k0 (both or neither): one more row per interior ladder knot m — the weights of "fit the natural spline
 * through the ladder WITHOUT knot m and evaluate it at x[m]".  The spline through a slice's quoted knots is also a spline on the
 * full ladder, so a slice with a MISSING quote (NaN) stays on the banded path: the tensor-path kernel fills the gap with that
 * 48-term row and goes on.  Rows are checked against the full solve when they are built (ladder_ok bit 2); a slice takes the
 * exact solve after all when another missing knot lies inside the row's window, the knot is within 24 of an end, or a value is
 * infinite. */
int64_t ivx_spline_band_fill_len(int64_t n_ladder_knots);
int ivx_spline_fit_eval_banded(int64_t n_slices, const int64_t* knot_off, const int64_t* x_off, const int32_t* slice_ladder,
                               const int32_t* order, const int64_t* lad_slice_off, int64_t n_ladders, int64_t max_slices_per_ladder,
                               const int64_t* ladder_off, int64_t max_knots, const double* x_ladder, const double* x_own /*or NULL*/,
                               const double* lu_cp, const double* lu_inv, const double* lu_ih, const double* y, int64_t n_q,
                               const double* xq, int32_t extrap, const int32_t* band_k0, const double* band_w,
                               const uint8_t* ladder_ok, const int16_t* band_perm, const double* band_frag, const int32_t* band_oct,
                               const double* band_fill_w, const int32_t* band_fill_k0 /* both or NULL */,
                               int32_t slab_knots_hint,
                               int32_t all_tensor_path /* 1: the caller read ladder_ok back and bit 1 is set for every ladder: the dot-product
                                                          kernel is not launched at all */,
                               int32_t x_verified /* 1: fb_list / fb_count already hold the verdict of an abscissa comparison done elsewhere
                                                     (ivx_bs_implied_vol_ladder): kept and added to; x_own is then read only by the
                                                     exact solve of the flagged slices */,
                               double* yq, int32_t* fb_list, uint32_t* fb_count, ivx_stream_t stream);
int ivx_spline_strided(int64_t n_systems, int32_t n_knots, const double* x /*n_knots*/,
                       const double* y /*[n_knots][n_systems]*/, int32_t n_q, const double* xq,
                       int32_t extrap, double* yq /*[n_q][n_systems]*/, ivx_stream_t stream);
/* n_batches independent planes (one per timestamp): batch b reads its abscissae at
 * x[b*x_batch_stride + i*x_knot_stride] (so they can be read in place from a per-quote column),
 * its knots at y[b][knot][system] and writes yq[b][query][system]. */
int ivx_spline_strided_batched(int64_t n_batches, int64_t n_systems, int32_t n_knots, const double* x,
                               int64_t x_batch_stride, int64_t x_knot_stride, const double* y,
                               int32_t n_q, const double* xq, int32_t extrap, double* yq,
                               ivx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data bridge — interpolated minute rows -> synthetic 1-minute OHLCV candles.
 * Replaces reference src/data_bridge/ohlcv_converter.py:138-400 (InterpolatedToOHLCVConverter.
 * _generate_ohlcv_from_interpolated, its four candle strategies, _process_volume :359-369 and
 * _validate_ohlcv_quality :371-402) and complete_pipeline.py:473-510 (_generate_ohlcv_candles), for a
 * CSR batch of symbols.  Rows whose base price is NaN or <= 0 emit nothing (:152).
 * The reference draws from numpy's unseeded global generator; here the random input is passed as tapes
 * in the reference's call order, symbol after symbol:
 *   u  uniforms in [0,1): k per emitted row (k = 5 spread_simulation, 2 price_as_midpoint, 1 simple
 *      spread, 4 pipeline, 0 trend_following); symbol s starts at u[k * out_off[s]];
 *   e  standard exponentials: one per emitted row whose volume is NaN or <= 0; symbol s starts at e[e_off[s]];
 *   z  standard normals (trend_following only): one per emitted row; symbol s starts at z[out_off[s]].
 * With tapes drawn from a RandomState seeded like the reference's generator the candles are bit-equal to
 * the reference's.  strategy: 0 spread_simulation (prev-close trend bias), 1 price_as_midpoint,
 * 2 trend_following, 3 simple spread (the reference's fallback for every other string), 4 pipeline.
 * ivx_bridge_plan counts, per segment of 2048 consecutive rows (ivx_bridge_n_segments(n_rows) of them, symbol
 * boundaries ignored) and per symbol, the candles and exponential draws that precede it: seg_out / seg_e
 * (n_seg + 1 entries) and out_off / e_off (n_sym + 1 entries; the last entries are the totals, which size the
 * outputs and the tapes).  o_row[j] is the global input row of candle j (timestamps and symbols are gathered
 * from it).  sym_quality[s]: bit 0 high<low, 1 high<open|close, 2 low>open|close, 3 (high-low)/source > 0.1,
 * 4 a price <= 0 — the checks of _validate_ohlcv_quality, which reports the lowest set bit.
 * ivx_bridge_ohlcv with seg_out, seg_e and a workspace of ivx_bridge_workspace_bytes(n_rows): one warp per segment,
 * whatever the symbols' lengths (a segment that starts inside a symbol derives the carried close by a 256-row warm-up
 * walk, which a check kernel compares bit for bit with what the previous segment produced; on a mismatch the
 * per-symbol kernel recomputes the batch in stream order).  With seg_out, seg_e or workspace NULL: one warp per
 * symbol.  Results are identical either way; trend_following always runs one warp per symbol.
 * ------------------------------------------------------------------------------------------- */
/* or-ed into `strategy`: run the per-symbol kernel after the segment kernel whatever the segment check says (diagnostic; same results) */
#define IVX_BRIDGE_RECOMPUTE 0x100
int64_t ivx_bridge_n_segments(int64_t n_rows);
int64_t ivx_bridge_workspace_bytes(int64_t n_rows);
int ivx_bridge_plan(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                    int64_t* seg_out /*n_seg+1*/, int64_t* seg_e /*n_seg+1*/, int64_t* out_off /*n_sym+1*/,
                    int64_t* e_off /*n_sym+1*/, void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
int ivx_bridge_ohlcv(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                     int32_t strategy, double base_spread_percent, double volatility_factor,
                     const double* u, const double* e, const double* z,
                     const int64_t* out_off /*n_sym+1*/, const int64_t* e_off /*n_sym+1*/,
                     const int64_t* seg_out /*n_seg+1 or NULL*/, const int64_t* seg_e /*n_seg+1 or NULL*/,
                     int64_t* o_row, double* o_open, double* o_high, double* o_low, double* o_close,
                     double* o_volume, double* o_source, uint32_t* sym_quality /*n_sym or NULL*/,
                     void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
/* The same candles with the tapes GENERATED inside the kernel — the reference's normal, unseeded mode (its draws are part of the
 * row loop, ohlcv_converter.py:209-263,359-369), where no tape has to exist in memory: element i of the uniform tape is
 * Philox4x32-10(seed, stream_u, i) and element i of the exponential tape -log1p(-Philox(seed, stream_e, i)), i.e. exactly the
 * numbers ivx_random_tape(n, seed, stream_u / stream_e, kind) writes — 40 B written and 40 B read per candle less than the
 * two-step form, same candles bit for bit.  ts / o_ts (both or neither): the candle's stamp is copied from its row's, so no
 * gather pass has to follow; o_row may then be NULL.  trend_following (normals) is not available here. */
int ivx_bridge_ohlcv_philox(int64_t n_sym, int64_t n_rows, const int64_t* in_off, const double* base, const double* volume,
                            const int64_t* ts /*n_rows or NULL*/, int32_t strategy, double base_spread_percent,
                            double volatility_factor, uint64_t seed, uint32_t stream_u, uint32_t stream_e,
                            const int64_t* out_off /*n_sym+1*/, const int64_t* e_off /*n_sym+1*/,
                            const int64_t* seg_out /*n_seg+1 or NULL*/, const int64_t* seg_e /*n_seg+1 or NULL*/,
                            int64_t* o_row /*or NULL*/, int64_t* o_ts /*or NULL*/, double* o_open, double* o_high, double* o_low,
                            double* o_close, double* o_volume, double* o_source, uint32_t* sym_quality /*n_sym or NULL*/,
                            void* workspace, int64_t workspace_bytes, ivx_stream_t stream);
/* the same quality bits for candles that already exist (ohlcv_converter.py:371-402 on a frame): *flags = OR over n rows */
int ivx_bridge_quality(int64_t n, const double* open, const double* high, const double* low, const double* close,
                       const double* source_price, uint32_t* flags /*1*/, ivx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Utilities: exclusive scan of int64, stream-copy and fp64-FMA microbenchmarks that give the
 * roofline denominators measured in the same run (bench.py).
 * ------------------------------------------------------------------------------------------- */
int ivx_exclusive_scan_i64(int64_t n, const int64_t* in, int64_t* out /*n+1*/, ivx_stream_t stream);
/* out[i] = src[idx[i]] (NaT / NaN where idx[i] is outside [0, n_src)): the timestamps and source prices of the bridge's
 * candles from o_row, so that the candles can feed ivx_ohlcv_reduce without leaving the device
 * (the reference does this through a DataFrame: ohlcv_converter.py:168-184, complete_pipeline.py:641-652) */
int ivx_gather_i64(int64_t n, const int64_t* idx, int64_t n_src, const int64_t* src, int64_t* out, ivx_stream_t stream);
int ivx_gather_f64(int64_t n, const int64_t* idx, int64_t n_src, const double* src, double* out, ivx_stream_t stream);
/* Random tapes for ivx_bridge_ohlcv produced on the device: element i is a function of (seed, stream_id, i) alone
 * (Philox4x32-10 counter generator; 53-bit uniforms built like numpy's random_sample).  kind 0: uniform [0,1);
 * kind 1: standard exponential.  This is the reference's normal mode — np.random unseeded (ohlcv_converter.py:213-263,
 * 359-369) — without the host drawing 5 numbers per row; parity runs pass host-drawn numpy tapes instead. */
int ivx_random_tape(int64_t n, uint64_t seed, uint32_t stream_id, int32_t kind, double* out, ivx_stream_t stream);
/* n <= 16 int64 values (output sizes of a plan) from device memory to out_host, a plain host array, without the copy engine: a
 * one-warp kernel stores them into mapped pinned host memory (128 bytes per device, allocated on first use) and the call waits
 * on an event behind it.  An 8-byte cudaMemcpy would queue behind every bulk device-to-host copy in flight on other streams.
 * Blocks until the stream has reached this point, like the cudaMemcpy it replaces. */
int ivx_read_back_i64(int32_t n, const int64_t* src, int64_t* out_host, ivx_stream_t stream);
int ivx_probe_copy(int64_t n_f64, const double* src, double* dst, ivx_stream_t stream);
int ivx_probe_fp64_fma(int64_t n_threads, int32_t iters, double* sink, ivx_stream_t stream);
/* kernel A's store pattern alone (11 fp64 + one int32 + one uint8 stream of n_rows, arrays stride_f64 doubles apart):
 * the write-side ceiling kernel A is measured against */
int ivx_probe_write_streams(int64_t n_rows, double* base, int64_t stride_f64, ivx_stream_t stream);
/* host-side evaluation of the multiply-high floor division the bucket kernels use (tests) */
int64_t ivx_host_floor_div(int64_t t, int64_t d);

#ifdef __cplusplus
}
#endif
#endif /* IVX_H_ */
