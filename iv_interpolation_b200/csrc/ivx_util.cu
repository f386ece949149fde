// ivx_util.cu — small device utilities around the hot path: row gathers (timestamps of the bridge's candles) and the
// counter-based random tapes the data bridge consumes when no host-drawn tape is supplied.
//
// The reference's bridge draws np.random.uniform / exponential row by row from numpy's unseeded global generator
// (reference src/data_bridge/ohlcv_converter.py:209-263,359-369).  For parity runs the host draws the tapes from
// numpy in the reference's order (iv_interpolation_b200/data_bridge/ohlcv_converter.py); for throughput runs —
// the reference's normal, unseeded mode — the tapes are produced here, on the device, by Philox4x32-10 keyed by
// (seed, element index): element i of a tape is the same number whatever the launch shape or the shard it lands in.
#include <limits>

#include "ivx_common.cuh"

namespace ivx {
namespace {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}

// 128 random bits for counter (idx, stream) under key `seed`
__device__ __forceinline__ uint4 philox4x32_10(uint64_t idx, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 53-bit uniform in [0, 1) from two words: numpy's random_sample construction, (a >> 5) * 2^26 + (b >> 6)) / 2^53
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// out[2j], out[2j+1] from one Philox block; kind 0: uniform [0,1); kind 1: standard exponential -log(1-u)
__global__ void __launch_bounds__(256) k_random_tape(int64_t n, uint64_t seed, uint32_t stream, int32_t kind, double* __restrict__ out) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = 2 * j;
  if (i >= n) return;
  const uint4 r = philox4x32_10((uint64_t)j, stream, seed);
  double a = u53(r.x, r.y), b = u53(r.z, r.w);
  if (kind == 1) { a = -log1p(-a); b = -log1p(-b); }
  if (i + 1 < n && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0)) {
    __stcs(reinterpret_cast<double2*>(out + i), make_double2(a, b));
  } else {
    st_stream(out + i, a);
    if (i + 1 < n) st_stream(out + i + 1, b);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather(int64_t n, const int64_t* __restrict__ idx, int64_t n_src, const T* __restrict__ src,
                                               T fill, T* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = __ldcs((const long long*)idx + i);
  out[i] = (k >= 0 && k < n_src) ? __ldg(src + k) : fill;
}

}  // namespace
}  // namespace ivx

using namespace ivx;

extern "C" int ivx_random_tape(int64_t n, uint64_t seed, uint32_t stream_id, int32_t kind, double* out, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0 && (kind == 0 || kind == 1), IVX_ERR_INVALID_ARG, "ivx_random_tape: n >= 0 and kind in {0 uniform, 1 exponential}");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(out != nullptr, IVX_ERR_INVALID_ARG, "ivx_random_tape: null output");
  const int64_t pairs = (n + 1) / 2;
  k_random_tape<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, seed, stream_id, kind, out);
  IVX_LAUNCH_CHECK("k_random_tape");
  return IVX_OK;
}

extern "C" int ivx_gather_i64(int64_t n, const int64_t* idx, int64_t n_src, const int64_t* src, int64_t* out, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0 && n_src >= 0, IVX_ERR_INVALID_ARG, "ivx_gather_i64: negative size");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(idx && src && out, IVX_ERR_INVALID_ARG, "ivx_gather_i64: null pointer");
  k_gather<int64_t><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, idx, n_src, src, (int64_t)IVX_NAT, out);
  IVX_LAUNCH_CHECK("k_gather<i64>");
  return IVX_OK;
}

extern "C" int ivx_gather_f64(int64_t n, const int64_t* idx, int64_t n_src, const double* src, double* out, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0 && n_src >= 0, IVX_ERR_INVALID_ARG, "ivx_gather_f64: negative size");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(idx && src && out, IVX_ERR_INVALID_ARG, "ivx_gather_f64: null pointer");
  k_gather<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, idx, n_src, src, std::numeric_limits<double>::quiet_NaN(), out);
  IVX_LAUNCH_CHECK("k_gather<f64>");
  return IVX_OK;
}
