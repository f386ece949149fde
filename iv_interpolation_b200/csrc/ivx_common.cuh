// ivx_common.cuh — shared device/host helpers for libivx (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ivx.h"

#define IVX_NAT INT64_MIN
#define IVX_WARP 32
#define IVX_FULL_MASK 0xffffffffu

namespace ivx {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define IVX_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ivx::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define IVX_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return ivx::cuda_fail(e__, #call); \
  } while (0)

#define IVX_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ivx::cuda_fail(e__, name);    \
  } while (0)

#define IVX_MAX_DEVICES 64
int current_device();
int sm_count();          // of the CURRENT device
// `static DeviceOnce once; if (once_per_device(once)) cudaFuncSetAttribute(...)`: true the first time on each device
struct DeviceOnce { bool done[IVX_MAX_DEVICES] = {}; };
bool once_per_device(DeviceOnce& o);

// ivx_spline.cu: exact natural-spline solve of the slices flagged in need[0 .. n_slices) (see k_spline_fit_eval_list);
// n_need (device, may be NULL): how many are flagged — zero lets every CTA leave at once
int spline_fit_eval_list(int64_t n_slices, const int32_t* need, const uint32_t* n_need, int64_t max_knots, const int64_t* knot_off, const int64_t* x_off,
                         const double* x, const double* lu_cp, const double* lu_inv, const double* lu_ih, const int32_t* slice_ladder,
                         const double* x_own, const int64_t* ladder_off, const double* y, int n_q, const double* xq, int extrap,
                         double* yq, cudaStream_t stream);

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Carver {
  char* base;
  int64_t off;
  int64_t cap;
  Carver(void* p, int64_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
  template <typename T>
  T* take(int64_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * (int64_t)sizeof(T);
    return p;
  }
  bool ok() const { return off <= cap; }
};

// ---- floor division by a launch-invariant positive divisor ---------------------------------------
// Unsigned 64-bit "round-up" magic (Granlund-Montgomery / libdivide branch-free form):
//   q = (((n - hi) >> 1) + hi) >> shift,  hi = mulhi(magic, n),  valid for every uint64 n, d >= 2.
struct FloorDiv {
  uint64_t magic;
  uint32_t shift;
  int64_t d;
  uint32_t trivial;  // d == 1
};

inline FloorDiv make_floor_div(int64_t d) {
  FloorDiv f;
  f.d = d;
  f.trivial = (d == 1);
  f.magic = 0;
  f.shift = 0;
  if (d < 2) return f;
  uint32_t lg = 63 - (uint32_t)__builtin_clzll((unsigned long long)d);
  if ((d & (d - 1)) == 0) {  // power of two: the fixed ">> 1" of the evaluation takes one bit
    f.magic = 0;
    f.shift = lg - 1;
    return f;
  }
  unsigned __int128 num = (unsigned __int128)1 << (64 + lg);
  unsigned __int128 m = num / (unsigned __int128)d;
  unsigned __int128 rem = num - m * (unsigned __int128)d;
  // proposed_m is floor(2^(64+lg)/d) with its top bit (2^64) implicit
  uint64_t proposed = (uint64_t)m;  // low 64 bits; the 2^64 bit is dropped (implicit)
  uint64_t r = (uint64_t)rem;
  proposed += proposed;
  uint64_t twice = r + r;
  if (twice >= (uint64_t)d || twice < r) proposed += 1;
  f.magic = proposed + 1;
  f.shift = lg;
  return f;
}

__host__ __device__ __forceinline__ uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

__host__ __device__ __forceinline__ uint64_t udiv_magic(uint64_t n, const FloorDiv& f) {
  uint64_t hi = mulhi_u64(f.magic, n);
  return (((n - hi) >> 1) + hi) >> f.shift;
}

// floor(t / d) for signed t, d >= 1
__host__ __device__ __forceinline__ int64_t floor_div(int64_t t, const FloorDiv& f) {
  if (f.trivial) return t;
  if (t >= 0) return (int64_t)udiv_magic((uint64_t)t, f);
  return -1 - (int64_t)udiv_magic((uint64_t)(~t), f);
}

#ifdef __CUDACC__
// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming loads/stores: data that is touched once should not displace reusable lines
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int64_t ld_stream(const int64_t* p) { return __ldcs((const long long*)p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(int64_t* p, int64_t v) { __stcs((long long*)p, (long long)v); }
__device__ __forceinline__ void st_stream(int32_t* p, int32_t v) { __stcs(p, v); }

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// last index i in [lo, hi) with key[i] <= v, given key[lo] <= v: 32-ary search by one warp (one round per 5 bits)
template <typename T>
__device__ __forceinline__ int64_t warp_upper_le(const T* __restrict__ key, int64_t lo, int64_t hi, T v, int lane) {
  while (hi - lo > 1) {
    const int64_t stepw = (hi - lo + 31) >> 5;
    const int64_t idx = lo + lane * stepw;
    const bool le = idx < hi && __ldg(key + idx) <= v;
    const uint32_t m = __ballot_sync(IVX_FULL_MASK, le) | 1u;
    const int last = 31 - __clz(m);
    lo += last * stepw;
    hi = min(hi, lo + stepw);
  }
  return lo;
}

// ---- counter-based random tapes (Philox4x32-10): element i of a tape is a function of (seed, stream, i) alone, so a tape
// written by k_random_tape (ivx_util.cu) and the same elements generated inside a consumer (ivx_bridge.cu) are the same numbers
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

// 53-bit uniform in [0, 1) from two words: numpy's random_sample construction, ((a >> 5) 2^26 + (b >> 6)) / 2^53.  The two
// integers are placed in the mantissas of 2^78 and 2^52 (no int-to-double conversion: those go through the narrow fp64
// pipe) and scaled by exact fused multiply-adds: hi 2^-27 + lo 2^-53, every step exact, so the value does not depend on how
// the translation unit contracts multiplies and adds (ivx_bridge.cu is compiled with --fmad=false).
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  const double H = __hiloint2double(0x44D00000, (int)(a >> 5));          // 2^78 + (a >> 5) 2^26
  const double L = __hiloint2double(0x43300000, (int)(b >> 6));          // 2^52 + (b >> 6)
  const double s = 1.0 / 9007199254740992.0;                              // 2^-53
  return fma(H, s, -33554432.0) + fma(L, s, -0.5);                        // (H - 2^78) 2^-53 + (L - 2^52) 2^-53
}

// standard exponential from a uniform in [0, 1)
__device__ __forceinline__ double exp_variate(double u) { return -log1p(-u); }
// element i of the uniform tape (seed, stream): block i >> 1, first or second double of it
__device__ __forceinline__ void philox_pair(uint64_t block, uint32_t stream, uint64_t seed, double& a, double& b) {
  const uint4 r = philox4x32_10(block, stream, seed);
  a = u53(r.x, r.y); b = u53(r.z, r.w);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif  // __CUDACC__

}  // namespace ivx
