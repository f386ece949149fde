// ivx_ohlcv.cu — kernel B: segmented OHLCV reduction over epoch-anchored time buckets.
//
// Replaces reference src/candle_reconstruction/core.py:59-89 (sort excluded: rows arrive sorted)
// and complete_pipeline.py:685-710 for a CSR batch of symbols in ONE pass over HBM:
//   * a tile of TB_ROWS rows (6 columns, 23 KB) is staged into shared memory with 1-D bulk async
//     copies (TMA engine) completing on an mbarrier;
//   * bucket ids come from a multiply-high floor division (no 64-bit divide per row);
//   * bucket heads are found with warp ballots, one thread then walks each bucket IN ROW ORDER —
//     first/last non-NaN, NaN-skipping max/min and the Kahan-compensated volume sum that pandas'
//     group_sum performs, which is order dependent and therefore not tree-reducible bit-exactly;
//   * emitted candles are compacted across tiles with a decoupled look-back (single pass).
// HBM traffic: 48 B per input row read + 52 B per emitted candle written.
#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr int TB_THREADS = 160;   // warp 0 runs the look-back while 4 warps walk ~96 five-row buckets of a tile in one pass
constexpr int TB_ROWS = 480;
constexpr int TB_PER = (TB_ROWS + TB_THREADS - 1) / TB_THREADS;   // rows per thread in the row-parallel phases
constexpr int TB_GROUPS = TB_ROWS / 32;
constexpr int64_t BID_NAT = INT64_MIN;  // bucket id given to rows without a timestamp

constexpr uint64_t ST_INVALID = 0, ST_AGG = 1, ST_INCL = 2;
__device__ __forceinline__ uint64_t st_pack(uint64_t st, uint32_t kept, uint32_t heads) {
  return (st << 62) | ((uint64_t)kept << 31) | (uint64_t)heads;
}
__device__ __forceinline__ uint32_t st_kept(uint64_t w) { return (uint32_t)((w >> 31) & 0x7fffffffu); }
__device__ __forceinline__ uint32_t st_heads(uint64_t w) { return (uint32_t)(w & 0x7fffffffu); }
__device__ __forceinline__ uint64_t st_status(uint64_t w) { return w >> 62; }

struct OhlcvParams {
  int64_t n_sym, n_rows, capacity;
  const int64_t* in_off;
  const int64_t* ts;
  const double* col[5];  // open high low close volume
  const uint8_t* sym_valid;
  FloorDiv fd;
  int64_t bucket_ticks;
  int32_t min_count;
  int64_t *o_ts;
  double* o_col[5];
  int32_t* o_count;
  int64_t *o_sym_row, *o_group;
  int64_t *raw_out_off, *raw_group_off;  // workspace, prefilled with -1
  uint32_t* sym_flags;
  int64_t* total;
  unsigned long long* tile_state;  // workspace, zeroed
  unsigned int* tile_counter;      // workspace, zeroed
  const uint32_t* tile_symhead;    // workspace [n_tiles][16]: bit i of the tile = a symbol starts at its row i (k_ohlcv_tile_plan)
  const int64_t* tile_sym_end;     // workspace [n_tiles]: end row of the symbol owning the tile's last row
  int32_t n_tiles;
  int32_t use_bulk;
};

struct TileSmem {
  int64_t ts[TB_ROWS];       // timestamps, overwritten in place by bucket ids
  double col[5][TB_ROWS];
  uint8_t valid[TB_ROWS];
  uint32_t head_mask[32];     // one word per 32 rows; sized for a warp-wide scan, words past TB_GROUPS stay 0
  uint32_t symhead_mask[32];
  uint32_t kept_mask[32];
  int32_t head_base[33];
  int32_t kept_base[33];
  uint16_t bstart[TB_ROWS + 1];
  int32_t symrow[TB_ROWS];   // per candle (at its first row): row of the first non-null symbol RELATIVE to the tile's first row
                             // (it may lie beyond the tile, in the last bucket's extension), INT32_MIN = none.  32 bits keep the
                             // struct under 28 KB: eight CTAs per SM instead of seven
  int32_t last_count_ext;    // rows of the tile's last bucket that live beyond the tile
  int64_t tile_sym_end;      // end row of the symbol owning the tile's last row
  uint32_t excl_kept, excl_heads;
  int32_t tile_id;
  int32_t s_lo;
  int32_t next_bucket;       // work counter of the bucket walk
  uint64_t bar;
};

__device__ __forceinline__ int64_t bucket_of(int64_t t, const FloorDiv& fd) {
  return t == IVX_NAT ? BID_NAT : floor_div(t, fd);
}

// index of the symbol owning global row r: largest s with in_off[s] <= r and in_off[s+1] > r
__device__ int64_t symbol_of_row(const int64_t* __restrict__ in_off, int64_t n_sym, int64_t r) {
  int64_t lo = 0, hi = n_sym;  // invariant: in_off[lo] <= r < in_off[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(in_off + mid) <= r) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint32_t row_flags(double o, double h, double l, double c, double v) {
  uint32_t f = 0;
  if ((h < l) | (h < o) | (h < c) | (l > o) | (l > c)) f |= IVX_FLAG_IN_BAD_OHLC;
  if (v < 0.0) f |= IVX_FLAG_IN_NEG_VOLUME;
  if (isnan(o) | isnan(h) | isnan(l) | isnan(c)) f |= IVX_FLAG_IN_NAN;
  return f;
}

// ---- per-tile plan: where symbols start inside every tile, and where the symbol of the tile's last row ends ---------------------
// The tiles used to find this out themselves — two 32-ary searches over in_off and a dependent load round, ~8 serialised L2 round
// trips at the head of every tile's critical path (ncu: 17 % of the kernel's stall samples sat on the first use of that load).
// One thread per symbol / per tile does it for the whole call in a few microseconds; a tile then reads 16 words and one int64.
__global__ void __launch_bounds__(256) k_ohlcv_tile_plan(int64_t n_sym, int64_t n_rows, int64_t n_tiles, const int64_t* __restrict__ in_off,
                                                         uint32_t* __restrict__ tile_symhead, int64_t* __restrict__ tile_sym_end) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_sym) {                                       // symbol i starts at row st (empty symbols mark the next one's start)
    const int64_t st = __ldg(in_off + i);
    if (st < n_rows) atomicOr(tile_symhead + (st / TB_ROWS) * 16 + ((st % TB_ROWS) >> 5), 1u << ((st % TB_ROWS) & 31));
  }
  if (i < n_tiles) {
    const int64_t last = min((i + 1) * (int64_t)TB_ROWS, n_rows) - 1;
    int64_t lo = 0, hi = n_sym;                          // in_off[lo] <= last < in_off[hi]: the LAST such lo (empty symbols before it)
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(in_off + mid) <= last) lo = mid; else hi = mid;
    }
    tile_sym_end[i] = __ldg(in_off + lo + 1);
  }
}

// running aggregate of one bucket, rows fed in order
struct Candle {
  double open, high, low, close, sum, comp;
  bool has_open;
  int64_t sym_row;
  uint32_t flags;
  __device__ __forceinline__ void init() {
    open = qnan(); high = qnan(); low = qnan(); close = qnan();
    sum = 0.0; comp = 0.0; has_open = false; sym_row = -1; flags = 0;
  }
  __device__ __forceinline__ void feed(double o, double h, double l, double c, double v, bool sv, int64_t grow) {
    if (!has_open && !isnan(o)) { open = o; has_open = true; }      // 'first' non-null
    if (!isnan(c)) close = c;                                       // 'last' non-null
    if (!isnan(h)) high = isnan(high) ? h : fmax(high, h);          // NaN-skipping max
    if (!isnan(l)) low = isnan(low) ? l : fmin(low, l);             // NaN-skipping min
    if (!isnan(v)) {                                                // pandas group_sum (Kahan)
      double y = __dsub_rn(v, comp);
      double t = __dadd_rn(sum, y);
      double cc = __dsub_rn(__dsub_rn(t, sum), y);
      comp = isnan(cc) ? 0.0 : cc;                                  // inf keeps the sum at inf
      sum = t;
    }
    if (sym_row < 0 && sv) sym_row = grow;
  }
};

__global__ void __launch_bounds__(TB_THREADS, 8) k_ohlcv_reduce(const OhlcvParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // ---- tile id in launch order (the look-back needs "earlier id => already running") ----------
  if (tid == 0) {
    S.tile_id = (int32_t)atomicAdd(P.tile_counter, 1u);
    mbar_init(&S.bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) { S.kept_mask[tid] = 0; S.head_mask[tid] = 0; }
  if (tid == 0) S.next_bucket = 0;
  __syncthreads();
  const int tile = S.tile_id;
  const int64_t r0 = (int64_t)tile * TB_ROWS;
  const int rows = (int)min((int64_t)TB_ROWS, P.n_rows - r0);

  // ---- stage the tile --------------------------------------------------------------------------
  const int bulk_rows = P.use_bulk ? (rows & ~1) : 0;  // 16-byte granules
  if (tid == 0 && bulk_rows > 0) {
    const uint32_t bytes = (uint32_t)bulk_rows * 8u;
    mbar_expect_tx(&S.bar, bytes * 6u);
    bulk_g2s(S.ts, P.ts + r0, bytes, &S.bar);
#pragma unroll
    for (int k = 0; k < 5; ++k) bulk_g2s(S.col[k], P.col[k] + r0, bytes, &S.bar);
  }
  for (int i = bulk_rows + tid; i < rows; i += TB_THREADS) {
    S.ts[i] = ld_stream(P.ts + r0 + i);
#pragma unroll
    for (int k = 0; k < 5; ++k) S.col[k][i] = ld_stream(P.col[k] + r0 + i);
  }
  if (P.sym_valid) {
    for (int i = tid; i < rows; i += TB_THREADS) S.valid[i] = P.sym_valid[r0 + i];
  }
  // symbol starts inside the tile and the end of its last row's symbol: the plan's words, one load round under the copy
  if (warp == 1) {
    S.symhead_mask[lane] = lane < 16 ? __ldg(P.tile_symhead + (int64_t)tile * 16 + lane) : 0u;
    if (lane == 0) S.tile_sym_end = __ldg(P.tile_sym_end + tile);
  }
  if (bulk_rows > 0) mbar_wait(&S.bar, 0);
  __syncthreads();

  // ---- bucket ids (in place) and sortedness ------------------------------------------------------
  // the per-row input checks of validate_candle_data (core.py:115-137) ride along here, row-parallel
  int64_t bid[TB_PER];
  uint32_t unsorted_rows = 0;
  uint32_t rflags[TB_PER];
#pragma unroll
  for (int k = 0; k < TB_PER; ++k) {
    int i = tid + k * TB_THREADS;
    rflags[k] = 0;
    if (i < rows) {
      int64_t t = S.ts[i];
      bid[k] = bucket_of(t, P.fd);
      int64_t tp = (i > 0) ? S.ts[i - 1] : (r0 > 0 ? __ldg(P.ts + r0 - 1) : IVX_NAT);
      bool symstart = (S.symhead_mask[i >> 5] >> (i & 31)) & 1u;
      if (!symstart && t != IVX_NAT && tp != IVX_NAT && t < tp) unsorted_rows |= 1u << k;
      if (P.sym_flags) rflags[k] = row_flags(S.col[0][i], S.col[1][i], S.col[2][i], S.col[3][i], S.col[4][i]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < TB_PER; ++k) {
    int i = tid + k * TB_THREADS;
    if (i < rows) S.ts[i] = bid[k];
  }
  if (P.sym_flags) {
#pragma unroll
    for (int k = 0; k < TB_PER; ++k) {
      const uint32_t f = rflags[k] | ((unsorted_rows >> k) & 1u ? IVX_FLAG_UNSORTED : 0u);
      if (f) atomicOr(P.sym_flags + symbol_of_row(P.in_off, P.n_sym, r0 + tid + k * TB_THREADS), f);
    }
  }
  __syncthreads();

  // ---- heads: a row starts a bucket if its id differs from the previous row's or a symbol starts -
  const int64_t prev_bid = (r0 > 0) ? bucket_of(__ldg(P.ts + r0 - 1), P.fd) : 0;
  for (int g = warp; g < TB_GROUPS; g += TB_THREADS / 32) {
    int i = g * 32 + lane;
    bool head = false;
    if (i < rows) {
      int64_t b = S.ts[i];
      int64_t bp = (i > 0) ? S.ts[i - 1] : prev_bid;
      bool symstart = (S.symhead_mask[g] >> lane) & 1u;
      head = symstart || (i == 0 && r0 == 0) || (b != bp);
    }
    uint32_t m = __ballot_sync(IVX_FULL_MASK, head);
    if (lane == 0) S.head_mask[g] = m;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of per-group head counts
    int c = __popc(S.head_mask[lane]);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(IVX_FULL_MASK, incl, d);
      if (lane >= d) incl += v;
    }
    S.head_base[lane] = incl - c;
    if (lane == 31) S.head_base[TB_GROUPS] = incl;
  }
  __syncthreads();
  const int n_heads = S.head_base[TB_GROUPS];
  for (int g = warp; g < TB_GROUPS; g += TB_THREADS / 32) {
    uint32_t m = S.head_mask[g];
    if ((m >> lane) & 1u) S.bstart[S.head_base[g] + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(g * 32 + lane);
  }
  if (warp == 0) {
    // rows of the last bucket that lie beyond this tile (ts only), 32 at a time
    int ext = 0;
    if (n_heads > 0) {
      const int64_t b = S.ts[rows - 1];
      int64_t r = r0 + rows;
      while (true) {
        const int64_t rr = r + lane;
        const bool same = rr < S.tile_sym_end && bucket_of(__ldg(P.ts + rr), P.fd) == b;
        const uint32_t m = __ballot_sync(IVX_FULL_MASK, same);
        const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;     // leading rows still in the bucket
        ext += run;
        if (run < 32) break;
        r += 32;
      }
    }
    if (lane == 0) { S.bstart[n_heads] = (uint16_t)rows; S.last_count_ext = ext; }
  }
  __syncthreads();

  // ---- which buckets are emitted: raw row count >= min_count (core.py:87-89) ---------------------
  for (int jb = 0; jb < n_heads; jb += TB_THREADS) {
    int j = jb + tid;
    bool kept = false;
    if (j < n_heads) {
      int cnt = (int)S.bstart[j + 1] - (int)S.bstart[j] + (j == n_heads - 1 ? S.last_count_ext : 0);
      kept = cnt >= P.min_count && S.ts[S.bstart[j]] != BID_NAT;
    }
    uint32_t m = __ballot_sync(IVX_FULL_MASK, kept);
    if (lane == 0) S.kept_mask[(jb >> 5) + warp] = m;
  }
  __syncthreads();
  if (warp == 0) {
    // ---- warp 0: kept ranks, then the decoupled look-back over earlier tiles — the other warps reduce meanwhile --
    int c = __popc(S.kept_mask[lane]);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(IVX_FULL_MASK, incl, d);
      if (lane >= d) incl += v;
    }
    S.kept_base[lane] = incl - c;
    if (lane == 31) S.kept_base[TB_GROUPS] = incl;
    const uint32_t agg_kept = (uint32_t)__shfl_sync(IVX_FULL_MASK, incl, 31);
    const uint32_t agg_heads = (uint32_t)n_heads;
    volatile unsigned long long* st = P.tile_state;
    uint32_t ex_k = 0, ex_h = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch((unsigned long long*)&P.tile_state[0], st_pack(ST_INCL, agg_kept, agg_heads));
    } else {
      if (lane == 0) atomicExch((unsigned long long*)&P.tile_state[tile], st_pack(ST_AGG, agg_kept, agg_heads));
      int look = tile - 1;
      while (true) {
        int idx = look - lane;
        unsigned long long w = (idx >= 0) ? st[idx] : st_pack(ST_INCL, 0, 0);
        // wait until every word of this window is published
        while (__any_sync(IVX_FULL_MASK, st_status(w) == ST_INVALID)) {
          if (st_status(w) == ST_INVALID) w = st[idx];
        }
        uint32_t incl_mask = __ballot_sync(IVX_FULL_MASK, st_status(w) == ST_INCL);
        int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
        uint32_t k = (lane <= first_incl) ? st_kept(w) : 0u;
        uint32_t h = (lane <= first_incl) ? st_heads(w) : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          k += __shfl_xor_sync(IVX_FULL_MASK, k, d);
          h += __shfl_xor_sync(IVX_FULL_MASK, h, d);
        }
        ex_k += k; ex_h += h;
        if (incl_mask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch((unsigned long long*)&P.tile_state[tile], st_pack(ST_INCL, ex_k + agg_kept, ex_h + agg_heads));
    }
    if (lane == 0) {
      S.excl_kept = ex_k; S.excl_heads = ex_h;
      if (tile == P.n_tiles - 1 && P.total) *P.total = (int64_t)ex_k + agg_kept;
    }
  } else {
    // ---- the other warps: one thread per bucket, rows in order; the candle is parked in the tile at its first row -----
    for (int j = tid - 32; j < n_heads; j += TB_THREADS - 32) {
      {
      const int b0 = S.bstart[j], b1 = S.bstart[j + 1];
      const bool kept = (S.kept_mask[j >> 5] >> (j & 31)) & 1u;
      Candle cd;
      cd.init();
      for (int i = b0; i < b1; ++i)
        cd.feed(S.col[0][i], S.col[1][i], S.col[2][i], S.col[3][i], S.col[4][i],
                P.sym_valid ? (S.valid[i] != 0) : true, r0 + i);
      if (j == n_heads - 1 && S.last_count_ext > 0) {
        for (int64_t r = r0 + rows; r < r0 + rows + S.last_count_ext; ++r)
          cd.feed(__ldg(P.col[0] + r), __ldg(P.col[1] + r), __ldg(P.col[2] + r), __ldg(P.col[3] + r),
                  __ldg(P.col[4] + r), P.sym_valid ? (P.sym_valid[r] != 0) : true, r);
      }
      uint32_t flags = cd.flags;
      if (kept) flags |= row_flags(cd.open, cd.high, cd.low, cd.close, cd.sum) << 3;  // IVX_FLAG_OUT_*
      S.col[0][b0] = cd.open; S.col[1][b0] = cd.high; S.col[2][b0] = cd.low; S.col[3][b0] = cd.close; S.col[4][b0] = cd.sum;
      S.symrow[b0] = cd.sym_row < 0 ? INT32_MIN : (int32_t)(cd.sym_row - r0);
      if (flags && P.sym_flags) atomicOr(P.sym_flags + symbol_of_row(P.in_off, P.n_sym, r0 + b0), flags);
      }
    }
  }
  __syncthreads();
  const int64_t base_kept = S.excl_kept;
  const int64_t base_heads = S.excl_heads;

  // ---- emit: consecutive kept candles -> consecutive addresses --------------------------------------------------------
  for (int j = tid; j < n_heads; j += TB_THREADS) {
    const int b0 = S.bstart[j];
    const bool kept = (S.kept_mask[j >> 5] >> (j & 31)) & 1u;
    const int64_t krank = base_kept + S.kept_base[j >> 5] + __popc(S.kept_mask[j >> 5] & ((1u << (j & 31)) - 1u));
    if ((S.symhead_mask[b0 >> 5] >> (b0 & 31)) & 1u) {
      const int64_t sym = symbol_of_row(P.in_off, P.n_sym, r0 + b0);
      P.raw_out_off[sym] = krank;
      P.raw_group_off[sym] = base_heads + j;
    }
    if (kept && krank < P.capacity) {
      const int cnt = (int)S.bstart[j + 1] - b0 + (j == n_heads - 1 ? S.last_count_ext : 0);
      st_stream(P.o_ts + krank, S.ts[b0] * P.bucket_ticks);
      st_stream(P.o_col[0] + krank, S.col[0][b0]);
      st_stream(P.o_col[1] + krank, S.col[1][b0]);
      st_stream(P.o_col[2] + krank, S.col[2][b0]);
      st_stream(P.o_col[3] + krank, S.col[3][b0]);
      st_stream(P.o_col[4] + krank, S.col[4][b0]);
      st_stream(P.o_count + krank, cnt);
      if (P.o_sym_row) st_stream(P.o_sym_row + krank, S.symrow[b0] == INT32_MIN ? (int64_t)-1 : r0 + S.symrow[b0]);
      if (P.o_group) st_stream(P.o_group + krank, base_heads + j);
    }
  }
}

// offsets of symbols without rows, per-symbol status
__global__ void k_ohlcv_finalize(int64_t n_sym, const int64_t* __restrict__ in_off, const int64_t* __restrict__ raw_out,
                                 const int64_t* __restrict__ raw_grp, const int64_t* __restrict__ total,
                                 int32_t min_count, int32_t check_len, int64_t* out_off, int64_t* sym_group_off,
                                 uint8_t* sym_status) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s > n_sym) return;
  if (s == n_sym) { out_off[s] = *total; return; }
  int64_t q = s;
  while (q < n_sym && raw_out[q] < 0) ++q;
  out_off[s] = q < n_sym ? raw_out[q] : *total;
  if (sym_group_off) sym_group_off[s] = raw_grp[s] < 0 ? 0 : raw_grp[s];
  if (sym_status) {
    int64_t n = in_off[s + 1] - in_off[s];
    sym_status[s] = (n == 0 || (check_len && n < min_count)) ? IVX_SYM_TOO_FEW : IVX_SYM_OK;
  }
}

// ---- time-major variant (surface pipeline) -------------------------------------------------------
__global__ void k_time_buckets(int64_t n_t, const int64_t* __restrict__ ts, FloorDiv fd, int64_t bucket_ticks,
                               int32_t* bucket_of_t, int64_t* bucket_ts, int32_t* bucket_start, int64_t* n_buckets) {
  // single block: n_t is the number of stamps of one step (thousands), not the number of quotes
  __shared__ int s_carry;
  __shared__ int s_warp[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t base = 0; base < n_t; base += blockDim.x) {
    int64_t t = base + threadIdx.x;
    bool head = false;
    int64_t b = 0;
    if (t < n_t) {
      b = floor_div(ts[t], fd);
      head = (t == 0) || (floor_div(ts[t - 1], fd) != b);
    }
    uint32_t m = __ballot_sync(IVX_FULL_MASK, head);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    int incl = before + __popc(m & ((2u << lane) - 1u));  // heads up to and including me
    if (t < n_t) {
      bucket_of_t[t] = incl - 1;
      if (head) { bucket_ts[incl - 1] = b * bucket_ticks; bucket_start[incl - 1] = (int32_t)t; }
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < nw; ++w) tot += s_warp[w]; s_carry += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *n_buckets = s_carry; bucket_start[s_carry] = (int32_t)n_t; }
}

// one thread per (bucket, series): consecutive threads = consecutive series -> coalesced rows
__global__ void __launch_bounds__(256) k_ohlc_time_major(int64_t n_t, int64_t n_series, const double* __restrict__ x,
                                                         int64_t n_buckets, const int32_t* __restrict__ bucket_start,
                                                         int32_t min_count, double* o_open, double* o_high,
                                                         double* o_low, double* o_close, int32_t* o_count) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (s >= n_series || b >= n_buckets) return;
  const int t0 = bucket_start[b], t1 = bucket_start[b + 1];
  double open = qnan(), high = qnan(), low = qnan(), close = qnan();
  bool has_open = false;
  const double* p = x + (int64_t)t0 * n_series + s;
#pragma unroll 4
  for (int t = t0; t < t1; ++t, p += n_series) {
    double v = ld_stream(p);
    if (!isnan(v)) {
      if (!has_open) { open = v; has_open = true; }
      close = v;
      high = isnan(high) ? v : fmax(high, v);
      low = isnan(low) ? v : fmin(low, v);
    }
  }
  const int cnt = t1 - t0;
  const bool keep = cnt >= min_count;
  const int64_t o = b * n_series + s;
  st_stream(o_open + o, keep ? open : qnan());
  st_stream(o_high + o, keep ? high : qnan());
  st_stream(o_low + o, keep ? low : qnan());
  st_stream(o_close + o, keep ? close : qnan());
  if (o_count && s == 0) o_count[b] = cnt;
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int64_t ivx_ohlcv_workspace_bytes(int64_t n_sym, int64_t n_rows) {
  int64_t n_tiles = (n_rows + TB_ROWS - 1) / TB_ROWS;
  return 256 * 4 + align_up((n_tiles + 1) * 8, 256) + 2 * align_up((n_sym + 1) * 8, 256) + align_up(n_tiles * 64, 256) + align_up(n_tiles * 8, 256) + 512;
}

extern "C" int ivx_ohlcv_reduce(int64_t n_sym, const int64_t* in_off, int64_t n_rows, const int64_t* ts,
                                const double* open, const double* high, const double* low, const double* close,
                                const double* volume, const uint8_t* sym_valid, int64_t bucket_ticks,
                                int32_t min_count, int32_t check_len, int64_t capacity, int64_t* out_off,
                                int64_t* o_ts, double* o_open, double* o_high, double* o_low, double* o_close,
                                double* o_volume, int32_t* o_count, int64_t* o_sym_row, int64_t* o_group,
                                int64_t* sym_group_off, uint8_t* sym_status, uint32_t* sym_flags,
                                int64_t* total, void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_sym >= 0 && n_rows >= 0, IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: negative size");
  IVX_REQUIRE(bucket_ticks >= 1, IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: bucket_ticks must be >= 1 (got %lld)", (long long)bucket_ticks);
  IVX_REQUIRE(n_rows < ((int64_t)1 << 31) - TB_ROWS, IVX_ERR_TOO_LARGE, "ivx_ohlcv_reduce: n_rows %lld exceeds 2^31 per call; split the batch", (long long)n_rows);
  IVX_REQUIRE(in_off && out_off && total && workspace, IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: null offsets/total/workspace");
  IVX_REQUIRE(n_rows == 0 || (ts && open && high && low && close && volume), IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: null input column");
  IVX_REQUIRE(capacity == 0 || (o_ts && o_open && o_high && o_low && o_close && o_volume && o_count), IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: null output column");
  IVX_REQUIRE(workspace_bytes >= ivx_ohlcv_workspace_bytes(n_sym, n_rows), IVX_ERR_INVALID_ARG, "ivx_ohlcv_reduce: workspace too small");

  const int64_t n_tiles = (n_rows + TB_ROWS - 1) / TB_ROWS;
  Carver cv(workspace, workspace_bytes);
  unsigned int* counter = cv.take<unsigned int>(256);
  unsigned long long* tile_state = cv.take<unsigned long long>(n_tiles + 1);
  int64_t* raw_out = cv.take<int64_t>(n_sym + 1);
  int64_t* raw_grp = cv.take<int64_t>(n_sym + 1);
  uint32_t* tile_symhead = cv.take<uint32_t>(n_tiles * 16);
  int64_t* tile_sym_end = cv.take<int64_t>(n_tiles);

  // the carved arrays are adjacent (256-byte aligned): one memset for the zeroed pair, one for the -1 pair
  IVX_CUDA(cudaMemsetAsync(counter, 0, (size_t)((char*)(tile_state + n_tiles + 1) - (char*)counter), stream));
  IVX_CUDA(cudaMemsetAsync(raw_out, 0xFF, (size_t)((char*)(raw_grp + n_sym + 1) - (char*)raw_out), stream));
  IVX_CUDA(cudaMemsetAsync(total, 0, 8, stream));
  if (sym_flags && n_sym) IVX_CUDA(cudaMemsetAsync(sym_flags, 0, n_sym * 4, stream));

  if (n_tiles > 0) {
    IVX_CUDA(cudaMemsetAsync(tile_symhead, 0, (size_t)n_tiles * 64, stream));
    const int64_t np = n_sym > n_tiles ? n_sym : n_tiles;
    k_ohlcv_tile_plan<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(n_sym, n_rows, n_tiles, in_off, tile_symhead, tile_sym_end);
    IVX_LAUNCH_CHECK("k_ohlcv_tile_plan");
    OhlcvParams P;
    P.n_sym = n_sym; P.n_rows = n_rows; P.capacity = capacity;
    P.in_off = in_off; P.ts = ts;
    P.col[0] = open; P.col[1] = high; P.col[2] = low; P.col[3] = close; P.col[4] = volume;
    P.sym_valid = sym_valid;
    P.fd = make_floor_div(bucket_ticks);
    P.bucket_ticks = bucket_ticks;
    P.min_count = min_count;
    P.o_ts = o_ts;
    P.o_col[0] = o_open; P.o_col[1] = o_high; P.o_col[2] = o_low; P.o_col[3] = o_close; P.o_col[4] = o_volume;
    P.o_count = o_count; P.o_sym_row = o_sym_row; P.o_group = o_group;
    P.raw_out_off = raw_out; P.raw_group_off = raw_grp;
    P.sym_flags = sym_flags; P.total = total;
    P.tile_state = tile_state; P.tile_counter = counter;
    P.tile_symhead = tile_symhead; P.tile_sym_end = tile_sym_end;
    P.n_tiles = (int32_t)n_tiles;
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    P.use_bulk = al16(ts) && al16(open) && al16(high) && al16(low) && al16(close) && al16(volume);
    static DeviceOnce attr_once;                       // the attribute is per device, not per process
    if (once_per_device(attr_once)) IVX_CUDA(cudaFuncSetAttribute(k_ohlcv_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem)));
    k_ohlcv_reduce<<<(unsigned)n_tiles, TB_THREADS, sizeof(TileSmem), stream>>>(P);
    IVX_LAUNCH_CHECK("k_ohlcv_reduce");
  }
  {
    int64_t n = n_sym + 1;
    k_ohlcv_finalize<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n_sym, in_off, raw_out, raw_grp, total, min_count,
                                                                    check_len, out_off, sym_group_off, sym_status);
    IVX_LAUNCH_CHECK("k_ohlcv_finalize");
  }
  return IVX_OK;
}

extern "C" int ivx_time_buckets(int64_t n_t, const int64_t* ts, int64_t bucket_ticks, int32_t* bucket_of_t,
                                int64_t* bucket_ts, int32_t* bucket_start, int64_t* n_buckets, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_t >= 0 && n_t < ((int64_t)1 << 31), IVX_ERR_INVALID_ARG, "ivx_time_buckets: bad n_t");
  IVX_REQUIRE(bucket_ticks >= 1, IVX_ERR_INVALID_ARG, "ivx_time_buckets: bucket_ticks must be >= 1");
  IVX_REQUIRE(ts && bucket_of_t && bucket_ts && bucket_start && n_buckets, IVX_ERR_INVALID_ARG, "ivx_time_buckets: null pointer");
  k_time_buckets<<<1, 1024, 0, stream>>>(n_t, ts, make_floor_div(bucket_ticks), bucket_ticks, bucket_of_t, bucket_ts,
                                          bucket_start, n_buckets);
  IVX_LAUNCH_CHECK("k_time_buckets");
  return IVX_OK;
}

extern "C" int ivx_ohlc_time_major(int64_t n_t, int64_t n_series, const double* x, int64_t n_buckets,
                                   const int32_t* bucket_start, int32_t min_count, double* o_open, double* o_high,
                                   double* o_low, double* o_close, int32_t* o_count, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_t >= 0 && n_series >= 0 && n_buckets >= 0 && n_buckets < 65536, IVX_ERR_INVALID_ARG, "ivx_ohlc_time_major: bad sizes");
  if (n_buckets == 0 || n_series == 0) return IVX_OK;
  IVX_REQUIRE(x && bucket_start && o_open && o_high && o_low && o_close, IVX_ERR_INVALID_ARG, "ivx_ohlc_time_major: null pointer");
  dim3 grid((unsigned)((n_series + 255) / 256), (unsigned)n_buckets);
  k_ohlc_time_major<<<grid, 256, 0, stream>>>(n_t, n_series, x, n_buckets, bucket_start, min_count, o_open, o_high,
                                              o_low, o_close, o_count);
  IVX_LAUNCH_CHECK("k_ohlc_time_major");
  return IVX_OK;
}
