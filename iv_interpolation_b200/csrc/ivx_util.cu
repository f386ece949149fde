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

// out[2j], out[2j+1] from one Philox block; kind 0: uniform [0,1); kind 1: standard exponential -log(1-u)
__global__ void __launch_bounds__(256) k_random_tape(int64_t n, uint64_t seed, uint32_t stream, int32_t kind, double* __restrict__ out) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = 2 * j;
  if (i >= n) return;
  const uint4 r = philox4x32_10((uint64_t)j, stream, seed);
  double a = u53(r.x, r.y), b = u53(r.z, r.w);
  if (kind == 1) { a = exp_variate(a); b = exp_variate(b); }
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

__global__ void k_store_i64(int n, const int64_t* __restrict__ src, volatile int64_t* dst) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}

}  // namespace
}  // namespace ivx

using namespace ivx;

// A few sizes from the device to the host WITHOUT the copy engine: a one-warp kernel stores them into mapped pinned host memory
// and the host waits on an event.  A cudaMemcpy of 8 bytes queues behind every bulk device-to-host copy already submitted on
// other streams (the copy queue is first in, first out): in a streaming pipeline the size read-back of batch k+1 then waits for
// the whole drain of batch k, and compute and drain stop overlapping.
#include <mutex>
extern "C" int ivx_read_back_i64(int32_t n, const int64_t* src, int64_t* out_host, ivx_stream_t stream_) {
  constexpr int CAP = 16;
  static std::mutex mu;
  static int64_t* mapped[IVX_MAX_DEVICES] = {};
  static cudaEvent_t done[IVX_MAX_DEVICES] = {};
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n >= 0 && n <= CAP, IVX_ERR_INVALID_ARG, "ivx_read_back_i64: 0 <= n <= %d", CAP);
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(src && out_host, IVX_ERR_INVALID_ARG, "ivx_read_back_i64: null pointer");
  int dev = 0;
  IVX_CUDA(cudaGetDevice(&dev));
  IVX_REQUIRE(dev >= 0 && dev < IVX_MAX_DEVICES, IVX_ERR_INVALID_ARG, "ivx_read_back_i64: device %d beyond the table", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!mapped[dev]) {                                     // 128 bytes of mapped pinned memory and one event per device, once
    IVX_CUDA(cudaHostAlloc((void**)&mapped[dev], CAP * sizeof(int64_t), cudaHostAllocMapped | cudaHostAllocPortable));
    IVX_CUDA(cudaEventCreateWithFlags(&done[dev], cudaEventDisableTiming));
  }
  int64_t* dptr = nullptr;
  IVX_CUDA(cudaHostGetDevicePointer((void**)&dptr, mapped[dev], 0));
  k_store_i64<<<1, 32, 0, stream>>>(n, src, dptr);
  IVX_LAUNCH_CHECK("k_store_i64");
  IVX_CUDA(cudaEventRecord(done[dev], stream));
  IVX_CUDA(cudaEventSynchronize(done[dev]));
  for (int i = 0; i < n; ++i) out_host[i] = mapped[dev][i];
  return IVX_OK;
}

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
