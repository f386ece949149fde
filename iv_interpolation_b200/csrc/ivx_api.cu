// ivx_api.cu — error plumbing and library identification for libivx.so.
#include <stdarg.h>
#include <string.h>

#include "ivx_common.cuh"

namespace ivx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return IVX_ERR_CUDA;
}

// per DEVICE, not per process: one process may drive several GPUs through this library
int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 ? dev : 0;
}

int sm_count() {
  static int cached[IVX_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (dev >= IVX_MAX_DEVICES) {
    int n = 0;
    return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0 ? n : 148;
  }
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;  // 148: B200
  }
  return cached[dev];
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember it per (call site, device)
bool once_per_device(DeviceOnce& o) {
  const int dev = current_device();
  if (dev >= IVX_MAX_DEVICES) return true;          // beyond the table: set the attribute every time (idempotent)
  if (o.done[dev]) return false;
  o.done[dev] = true;
  return true;
}

}  // namespace ivx

extern "C" int ivx_version(void) { return IVX_VERSION; }
extern "C" const char* ivx_last_error_string(void) { return ivx::g_err; }
extern "C" const char* ivx_build_info(void) {
  return "libivx sm_100a: k_interp_plan,k_interp_exec,k_scan_i64,k_ohlcv_tile_plan,k_ohlcv_reduce,k_ohlcv_finalize,k_time_buckets,"
         "k_ohlc_time_major,k_bs_implied_vol,k_bs_implied_vol_ladder,k_bs_implied_vol_chain,k_bs_implied_vol_retry,k_chain_terms,"
         "k_bs_price,k_bs_greeks,k_spline_fit_eval,k_spline_fit_eval_list,k_spline_ladder_factor,k_spline_strided,"
         "k_band_build,k_band_check,k_band_perm,k_band_frag,k_band_apply,k_band_apply_mma,"
         "k_bridge_count_seg,k_bridge_sym_off,k_bridge_ohlcv_seg,k_bridge_check,k_bridge_ohlcv,k_bridge_quality,k_random_tape,k_gather,k_store_i64,k_probe_copy,k_probe_fp64";
}

// host-side check of the multiply-high floor division used by the bucket kernels (tests call this)
extern "C" int64_t ivx_host_floor_div(int64_t t, int64_t d) {
  if (d < 1) return 0;
  ivx::FloorDiv f = ivx::make_floor_div(d);
  return ivx::floor_div(t, f);
}
