// ivx_bsiv.cu — kernel C: fused fp64 Black-Scholes implied-volatility solver, one thread per quote.
//
// The reference has no solver; conventions follow reference
// src/interpolation/greeks.py:21-22 (continuous rate r, no dividend, T in years):
//     d1 = (ln(S/K) + (r + sigma^2/2) T) / (sigma sqrt T),  d2 = d1 - sigma sqrt T.
//
// Method.  With Kd = K e^{-rT}, x = -|ln(S/Kd)| and s = sigma sqrt(T) the out-of-the-money price,
// normalised by sqrt(S Kd), is
//     b(x,s) = e^{x/2} Phi(x/s + s/2) - e^{-x/2} Phi(x/s - s/2)      (in-the-money quotes are first
// moved across put-call parity).  b is increasing in s with its inflection at s_c = sqrt(2|x|):
//   * beta <  b(s_c): root in the convex branch — solve ln b(s) = ln beta, evaluated through
//     erfcx so that deep-wing prices never underflow: b = E/2 [erfcx(q1) - erfcx(q2)],
//     E = exp(-(h^2+t^2)/2), h = |x|/s, t = s/2, q1,2 = (h -+ t)/sqrt2, b' = E/sqrt(2 pi);
//   * beta >= b(s_c): root in the concave branch — solve b(s) = beta directly.
// The kernel is bound by instruction issue and the fp64 pipe together (ncu: profiles/README.md), so it is built to
// execute as few instructions as possible, in three phases:
//   1. an fp32 PRE-SOLVE (FP32 pipe + raw MUFU ops, runs in the shadow of other warps' fp64 work): a TABLE guess — log2 s is
//      asymptotically linear in (log2 sqrt(2|x|), log2(-ln(beta e^{-x/2}))) in both the wing and the at-the-money regime,
//      so a 76 x 79 table with bilinear interpolation is ~1e-3 accurate (tools/gen_iv_guess_table.py) — followed by one
//      Halley step in single precision: s to fp32 precision on all but a few 1e-5 of the quotes;
//   2. an fp64 FAST PATH, straight-line: ONE Halley step on b(s) - beta with b written so that the same expression
//      holds on both sides of the inflection point (solve_quote); accepted when the step is below 1e-4 s — after such
//      a step the remaining error is ~1e-12 s;
//   3. the out-of-line FALLBACK for everything the fast path does not accept: bracketed Halley iterations (third
//      order) with bisection of the bracket whenever a step leaves it ("Newton/Brent" safeguarding), a warp vote to
//      leave the loop (polish_f64).  It takes a quote from anywhere; on realistic tapes it runs for ~0 % of the quotes.
// Inputs: S, K, T positive NORMAL numbers below 1e300 (subnormal or larger inputs return NaN), any finite r.
// Loads are structure-of-arrays, one coalesced stream per column.
#include <stdlib.h>

#include "ivx_common.cuh"

namespace ivx {
namespace {

constexpr double INV_SQRT2 = 0.7071067811865476;
constexpr double SQRT_2PI = 2.5066282746310002;
constexpr double INV_SQRT_2PI = 0.3989422804014327;
constexpr double S_MAX = 40.0;   // sigma*sqrt(T) bracket; fp64 prices cannot encode more than ~17
constexpr int MAX_ITERS = 64;
constexpr int PRE_ITERS = 1;
constexpr double ACCEPT_STEP = 1e-4;   // relative size of a Halley step after which s is converged (cubic)

__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * INV_SQRT2); }
__device__ __forceinline__ double norm_pdf(double z) { return INV_SQRT_2PI * exp(-0.5 * z * z); }

// ---- set-up helpers: cheaper than the general library routines on the ranges the solver feeds them ---------------
// Polynomial coefficients live in constant memory: as immediates every fp64 coefficient costs two UMOV issue slots
// next to the DFMA that uses it, and this kernel is bound by instruction issue (profiles/README.md).
__constant__ __align__(16) double EXP_C[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                 1.0 / 40320.0,      1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,
                                 1.0 / 6.0,          0.5,               1.0,              1.0};
// Reciprocal and reciprocal square root of a NORMAL a (1e-300 < a < 1e300) without the IEEE special-case tail of the
// compiler's division: the 20-bit hardware seed (MUFU.RCP64H / RSQ64H) and ONE third-order step (2^-20 -> 2^-60).
__device__ __forceinline__ double rcp_fast(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y, 1.0);                  // 2^-20
  return fma(y, fma(e, e, e), y);                    // y (1 + e + e^2): one cubic step, 2^-60
}
__device__ __forceinline__ double rcp_fast1(double a) {   // one Newton step: 2^-40
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return fma(y, fma(-a, y, 1.0), y);
}
__device__ __forceinline__ double rsqrt_fast(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-(a * y), y, 1.0);            // 2^-20
  return fma(y, e * fma(0.375, e, 0.5), y);          // y (1 + e/2 + 3 e^2/8): one cubic step, 2^-60
}
// positive, finite and normal: high word in [0x00100000, 0x7ff00000)
__device__ __forceinline__ bool pos_normal(double v) { return (unsigned)(__double2hiint(v) - 0x00100000) < 0x7fe00000u; }

// ln a for a finite, positive, normal a: a = m 2^e, m in [1, 2); the top six mantissa bits pick c_j = 1 + (j + 1/2)/64 from a
// table of {rc_j ~ 1/c_j, -ln rc_j}; u = m rc_j - 1 is exact to 2^-60 in one fma, |u| < 2^-7, and ln(1+u) is its series to
// u^8 (remainder 1e-20).  No division.  Absolute error ~1e-16 (the result is a log-moneyness, used to that accuracy).
__device__ const double2 LOG_TAB[64] = {
    {0.9922480620155039, 0.007782140442054963},
    {0.9770992366412213, 0.023167059281534418},
    {0.9624060150375939, 0.03831886430213666},
    {0.9481481481481482, 0.05324451451881224},
    {0.9343065693430657, 0.06795066190850778},
    {0.920863309352518, 0.08244366921107454},
    {0.9078014184397163, 0.09672962645855114},
    {0.8951048951048951, 0.11081436634029011},
    {0.8827586206896552, 0.12470347850095725},
    {0.8707482993197279, 0.1384023228591192},
    {0.8590604026845637, 0.151916042025842},
    {0.847682119205298, 0.16524957289530717},
    {0.8366013071895425, 0.17840765747281825},
    {0.8258064516129032, 0.19139485299962947},
    {0.8152866242038217, 0.20421554142869083},
    {0.8050314465408805, 0.2168739383006143},
    {0.7950310559006211, 0.2293741010648459},
    {0.7852760736196319, 0.24171993688714513},
    {0.7757575757575758, 0.25391520998096345},
    {0.7664670658682635, 0.2659635484971379},
    {0.757396449704142, 0.2778684510034563},
    {0.7485380116959064, 0.2896332925830427},
    {0.7398843930635838, 0.30126133057816185},
    {0.7314285714285714, 0.3127557100038969},
    {0.7231638418079096, 0.324119468654212},
    {0.7150837988826816, 0.3353555419211378},
    {0.7071823204419889, 0.3464667673462086},
    {0.6994535519125683, 0.3574558889218038},
    {0.6918918918918919, 0.36832556115870757},
    {0.6844919786096256, 0.3790783529349695},
    {0.6772486772486772, 0.38971675114002524},
    {0.6701570680628273, 0.40024316412701266},
    {0.6632124352331606, 0.4106599249852683},
    {0.6564102564102564, 0.42096929464412963},
    {0.649746192893401, 0.43117346481837143},
    {0.6432160804020101, 0.4412745608048752},
    {0.6368159203980099, 0.4512746441394586},
    {0.6305418719211823, 0.46117571512217015},
    {0.624390243902439, 0.470979715218791},
    {0.6183574879227053, 0.48068852934575196},
    {0.6124401913875598, 0.4903039880451939},
    {0.6066350710900474, 0.49982786955644926},
    {0.6009389671361502, 0.5092619017898079},
    {0.5953488372093023, 0.5186077642080457},
    {0.5898617511520737, 0.5278670896208424},
    {0.5844748858447488, 0.5370414658968837},
    {0.579185520361991, 0.5461324375981356},
    {0.5739910313901345, 0.5551415075405016},
    {0.5688888888888889, 0.564070138284803},
    {0.5638766519823789, 0.5729197535617854},
    {0.5589519650655022, 0.5816917396346225},
    {0.5541125541125541, 0.5903874466021763},
    {0.5493562231759657, 0.5990081896460834},
    {0.5446808510638298, 0.6075552502245418},
    {0.540084388185654, 0.616029877215514},
    {0.5355648535564853, 0.6244332880118936},
    {0.5311203319502075, 0.6327666695710378},
    {0.5267489711934157, 0.6410311794209312},
    {0.5224489795918368, 0.6492279466251097},
    {0.5182186234817814, 0.65735807270836},
    {0.5140562248995983, 0.6654226325450905},
    {0.5099601593625498, 0.6734226752121667},
    {0.5059288537549407, 0.6813592248079031},
    {0.5019607843137255, 0.689233281238809}};
// full-mantissa scalars of the fast path, in the constant bank for the same reason (one LDCU instead of two UMOVs each)
__constant__ __align__(16) double KF[8] = {6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.4426950408889634, 0.7071067811865476,
                             2.5066282746310002, 1e-4, 1e-30, 2.3e-308};
#define KF_LN2_HI KF[0]
#define KF_LN2_LO KF[1]
#define KF_LOG2E KF[2]
#define KF_INV_SQRT2 KF[3]
#define KF_SQRT_2PI KF[4]
#define KF_ACCEPT KF[5]
#define KF_1EM30 KF[6]
#define KF_MIN_RATIO KF[7]
__constant__ __align__(16) double LOG1P_C[7] = {-1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5};
__device__ __forceinline__ double log_normal(double a) {
  const int hi = __double2hiint(a);
  const int e = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a));   // [1, 2)
  const double2 tb = __ldg(&LOG_TAB[(hi >> 14) & 63]);
  const double u = fma(m, tb.x, -1.0);
  double p = LOG1P_C[1];                               // series to u^7: the u^8 term is below 2^-59
#pragma unroll
  for (int k = 2; k < 7; ++k) p = fma(p, u, LOG1P_C[k]);
  const double l = fma(u * u, p, u);
  const double ed = (double)e;
  return fma(ed, KF_LN2_HI, tb.y + fma(ed, KF_LN2_LO, l));   // ln2 split hi + lo
}
// e^{w} for |w| <= 0.35: Taylor to w^13 (truncation 2e-18 at ln2/2)
template <int FIRST = 0>                               // FIRST = 3: Taylor to w^10 (2e-13 at ln2/2)
__device__ __forceinline__ double exp_poly(double w) {
  double p = EXP_C[FIRST];
#pragma unroll
  for (int k = FIRST + 1; k < 14; ++k) p = fma(p, w, EXP_C[k]);
  return p;
}
// e^{-z} for |z| <= 1/4 (discount factors); the library exp elsewhere
__device__ __forceinline__ double exp_neg_small(double z) {
  if ((__double2hiint(z) & 0x7fffffff) < 0x3FB00000) return exp_poly<5>(-z);   // |z| < 1/16 (high-word compare): Taylor to w^8, 4e-17 (r T of weeks to a year at a few per cent)
  if (!(fabs(z) <= 0.25)) return exp(-z);
  return exp_poly(-z);
}
// e^{-z} for 0 <= z <= 700: z = n ln2 + r, |r| <= ln2/2, e^{-z} = 2^{-n} e^{-r}
template <int FIRST = 0>
__device__ __forceinline__ double exp_neg_core(double z) {
  const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52: the low word of z*log2e + MAGIC is rint(z*log2e)
  const double tn = fma(z, KF_LOG2E, MAGIC);
  const int n = __double2loint(tn);
  const double nd = tn - MAGIC;
  double r = fma(nd, -KF_LN2_HI, z);
  r = fma(nd, -KF_LN2_LO, r);
  return exp_poly<FIRST>(-r) * __hiloint2double((1023 - n) << 20, 0);
}
// e^{-q/2} / sqrt(2 pi) for 0 <= q < 1400 (the fast path's b' = E/sqrt(2 pi), E = e^{-(h^2+t^2)/2}): q/2 = n ln2 + r; the halving is
// folded into the constants (q log2e/2, 2 ln2, coefficients 1/(k! 2^k) to k = 10: 2e-13 at |r| = ln2/2) and 1/sqrt(2 pi) = 1.5958 x 2^-2
// into the power of two that is assembled from n.  One multiple of ln2, not a hi + lo pair: n <= 1010 and an ulp of 2 ln2 move
// r by 1.5e-13 at most, the size of the truncation.
__constant__ __align__(16) double EXPH_C[12] = {2.69114445546737190e-10, 5.38228891093474463e-09, 9.68812003968253967e-08, 1.55009920634920635e-06,
                                  2.17013888888888897e-05, 2.60416666666666663e-04, 2.60416666666666652e-03, 2.08333333333333322e-02,
                                  0.125,                   0.5,                     0.7213475204444817,      1.3862943611198906};
__device__ __forceinline__ double exp_half_neg_over_sqrt2pi(double q) {
  const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52
  const double tn = fma(q, EXPH_C[10], MAGIC);
  const int n = __double2loint(tn);
  const double nd = tn - MAGIC;
  const double w = fma(nd, EXPH_C[11], -q);                       // -(q - 2 n ln2) = -2r
  double p = EXPH_C[0];
#pragma unroll
  for (int k = 1; k < 10; ++k) p = fma(p, w, EXPH_C[k]);
  p = fma(p, w, 1.0);
  return p * __hiloint2double(((1021 - n) << 20) | 0x98845, 0x33d43651);   // 2^-n / sqrt(2 pi)
}
// the library exp outside that range
__device__ __forceinline__ double exp_neg(double z) {
  if (!(z >= 0.0 && z <= 700.0)) return exp(-z);
  const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52: the low word of z*log2e + MAGIC is rint(z*log2e)
  const double tn = fma(z, 1.4426950408889634, MAGIC);
  const int n = __double2loint(tn);
  const double nd = tn - MAGIC;
  double r = fma(nd, -6.93147180369123816490e-01, z);
  r = fma(nd, -1.90821492927058770002e-10, r);
  const double pw = exp_poly(-r);
  return pw * __hiloint2double((1023 - n) << 20, 0);
}

// single-precision MUFU operations without the denormal-range fix-ups the compiler wraps around __fdividef, __logf,
// __expf, sqrtf and rsqrtf (flush-to-zero is fine here: the pre-solve only has to land near the root)
__device__ __forceinline__ float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_log(float x) { return 0.69314718f * f_lg2(x); }
__device__ __forceinline__ float f_exp(float x) { return f_ex2(1.44269504f * x); }

// erfcx(x) for x >= 0 in single precision: (x + 2) erfcx(x) is smooth in t = (x-2)/(x+2) = 1 - 4/(x+2) on [-1, 1), so one
// reciprocal serves the argument and the prefactor; degree-9 fit, 4.5e-7 relative as evaluated in fp32
// (tools/gen_erfcx_fit.py) — 13 instructions.
__device__ __forceinline__ float erfcx_pos_f(float x) {
  const float rc = f_rcp(x + 2.0f);
  const float t = fmaf(-4.0f, rc, 1.0f);
  float p = 7.254005613e-05f;
  p = fmaf(p, t, -3.868958294e-04f);
  p = fmaf(p, t, -8.597664146e-04f);
  p = fmaf(p, t, 2.871722701e-03f);
  p = fmaf(p, t, 6.485639701e-03f);
  p = fmaf(p, t, -2.143603228e-02f);
  p = fmaf(p, t, -3.644295417e-02f);
  p = fmaf(p, t, 2.794629614e-01f);
  p = fmaf(p, t, -6.871606703e-01f);
  p = fmaf(p, t, 1.021582887e+00f);
  return p * rc;
}

// ---- phase 1: single precision -------------------------------------------------------------------------
#include "ivx_bsiv_guess.inc"
// log2 s by bilinear interpolation in P = log2 sqrt(-2x), V = log2(-ln(beta e^{-x/2})) (clamped to the table: x = 0 gives
// P = -inf -> first row, which IS the at-the-money limit).  NaN coordinates clamp to 0 (fmaxf returns the other operand).
__device__ __forceinline__ float guess_table(float x, float lnbeta) {
  const float P = 0.5f * f_lg2(-2.0f * x);
  const float V = f_lg2(fmaf(0.5f, x, -lnbeta));
  const float fp = fminf(fmaxf((P - GUESS_P0) * GUESS_CELLS, 0.0f), (float)(GUESS_NP - 1) - 0.001f);
  const float fv = fminf(fmaxf((V - GUESS_V0) * GUESS_CELLS, 0.0f), (float)(GUESS_NV - 1) - 0.001f);
  const int i = (int)fp, j = (int)fv;
  const float wp = fp - (float)i, wv = fv - (float)j;
  const float* t = GUESS_TAB + (i * GUESS_NV + j);
  const float t00 = __ldg(t), t01 = __ldg(t + 1), t10 = __ldg(t + GUESS_NV), t11 = __ldg(t + GUESS_NV + 1);
  const float a0 = fmaf(wv, t01 - t00, t00), a1 = fmaf(wv, t11 - t10, t10);
  return f_ex2(fmaf(wp, a1 - a0, a0));
}

// x <= 0, 0 < beta < e^{x/2}; s = the table guess; `lower`: the root sits below the inflection point sqrt(2|x|) (objective
// ln b there, b above).  Returns s after one Householder step.
template <int LOWER>                                 // 1 / 0: the branch is known at compile time (straight-line code); -1: `lower` decides
__device__ __forceinline__ float presolve_f32_t(float x, float beta, float lnbeta, float ex2, float s, bool lower) {
  if (LOWER >= 0) lower = LOWER != 0;
  const float kInvSqrt2 = 0.70710678f, kSqrt2Pi = 2.50662827f;
#pragma unroll
  for (int it = 0; it < PRE_ITERS; ++it) {
    const float inv_s = f_rcp(s);
    const float h = -x * inv_s, t = 0.5f * s;        // h >= 0
    const float hs = h * inv_s;
    const float r2 = h * hs - 0.25f * s;              // b2/b1: second over first derivative of b
    // Halley's step (cubic convergence): dn / (1 + A dn / 2), dn = -g/g1, A = g2/g1 (derivatives of the objective g).  The table
    // guess is within 1.6e-2 (1e-3 typically): one cubic step leaves s at single-precision rounding; the third-order Householder
    // step the closed-form guess (12 % at the 99th percentile) needed is no longer worth its ten instructions.
    float g, newton, A;
    if (lower) {
      const float q1 = (h - t) * kInvSqrt2;           // < 0 only if s overshot the inflection point
      float e1 = erfcx_pos_f(fabsf(q1));
      if (q1 < 0.0f) e1 = 2.0f * f_exp(q1 * q1) - e1;   // erfcx(-a) = 2 e^{a^2} - erfcx(a)
      const float D = e1 - erfcx_pos_f((h + t) * kInvSqrt2);
      g = -0.5f * (h * h + t * t) + f_log(0.5f * D) - lnbeta;
      newton = -g * D * (0.5f * kSqrt2Pi);           // -g / d1,  d1 = (ln b)' = 2 / (sqrt(2 pi) D)
      const float d1 = (2.0f / kSqrt2Pi) * f_rcp(D);
      A = r2 - d1;
    } else {
      // same identity as the fp64 polish: b = e^{x/2} - E/2 [erfcx((t-h)/sqrt2) + erfcx((t+h)/sqrt2)], E = e^{-(h^2+t^2)/2}
      const float E = f_exp(-0.5f * (h * h + t * t));
      const float b = ex2 - 0.5f * E * (erfcx_pos_f(fmaxf(t - h, 0.0f) * kInvSqrt2) + erfcx_pos_f((h + t) * kInvSqrt2));
      g = b - beta;
      newton = (-g * kSqrt2Pi) * f_rcp(E);
      A = r2;
    }
    const float fac = f_rcp(fmaf(0.5f * A, newton, 1.0f));
    // no bracket and no range guard here: whatever this step produces, the fp64 phase accepts only a converged value and
    // sends the rest (NaN included) to the bracketed loop
    s = s + ((fac > 0.3f && fac < 3.0f) ? newton * fac : newton);
  }
  return s;
}
__device__ __forceinline__ float presolve_f32(float x, float beta, float lnbeta, float ex2, float s, bool lower) {
  return presolve_f32_t<-1>(x, beta, lnbeta, ex2, s, lower);
}

// ---- phase 2: double precision ---------------------------------------------------------------------------
// erfcx(q1), erfcx(q2) for q1, q2 >= 0 with ONE division: g(t) = (q + 4) erfcx(q) is smooth in t = (q-4)/(q+4) = 1 - 8/(q+4)
// on [-1, 1) (it tends to 1/sqrt(pi), analytic at t = 1), so the reciprocal of q + 4 gives both the argument and the
// prefactor, and the two reciprocals are peeled off one reciprocal of the product (q1+4)(q2+4).  Degree-20 Chebyshev
// interpolant converted to the monomial basis at 50 digits (tools/gen_erfcx_fit.py): 2.4e-15 relative as evaluated in
// fp64, smooth in q, so the difference of two nearby values keeps that relative accuracy.  The two Horner chains
// interleave.  (The first version fitted (1 + 2q) erfcx(q) — the same degree for 8e-15, but a second reciprocal 1/(1+2q)
// per value: ten more fp64 operations per pair.)
__constant__ __align__(16) double ERFCX_C[21] = {
    -8.93653518802207727e-09,
    2.22219577646707465e-09,
    1.08269424077762936e-07,
    -7.91273327604327001e-08,
    -8.25704434017688251e-07,
    1.42894055117218721e-06,
    4.67854161930677550e-06,
    -1.88645718924317614e-05,
    -3.62746818991857262e-06,
    1.76814599407607931e-04,
    -4.55055022890642181e-04,
    -2.80959402893106306e-04,
    6.11205608281303880e-03,
    -2.63700532427967073e-02,
    7.63815148663746779e-02,
    -1.74010937249687136e-01,
    3.30851587879994469e-01,
    -5.40853831312786415e-01,
    7.73208702265210102e-01,
    -9.76548729080888078e-01,
    1.09599566100049106e+00};

// the same fit at degree 18, times sqrt(2 pi)/2: the fast path wants E/sqrt(2 pi) (= b') times [sqrt(2 pi)/2 (erfcx - erfcx)]
__constant__ __align__(16) double ERFCX_CS[19] = {      // degree 18 (4e-14 as evaluated: the fast path's b needs no more), times sqrt(2 pi)/2
    8.23961567720295452e-08,
    -8.59628711319826841e-08,
    -9.28046551104987971e-07,
    1.76454135697427688e-06,
    5.74657144711677158e-06,
    -2.36144512506849318e-05,
    -4.47002892624767972e-06,
    2.21585572078349753e-04,
    -5.70356973586231652e-04,
    -3.52123079228094710e-04,
    7.66033326373934641e-03,
    -3.30499622129079088e-02,
    9.57300315369120775e-02,
    -2.18090367493141701e-01,
    4.14660972493239299e-01,
    -6.77859753017408417e-01,
    9.69073397643521184e-01,
    -1.22392232793436539e+00,
    1.37362685636836201e+00};

template <bool SCALED = false>
__device__ __forceinline__ void erfcx_pair_pos(double q1, double q2, double& e1, double& e2) {
  const double* C = SCALED ? ERFCX_CS : ERFCX_C;
  const double a1 = q1 + 4.0, a2 = q2 + 4.0;
  const double r = rcp_fast(a1 * a2);               // q < 1e70: the product is normal
  const double r1 = r * a2, r2 = r * a1;             // 1/(q1+4), 1/(q2+4)
  const double t1 = fma(-8.0, r1, 1.0), t2 = fma(-8.0, r2, 1.0);
  constexpr int DEG = SCALED ? 18 : 20;
  double p1 = C[0], p2 = C[0];
#pragma unroll
  for (int k = 1; k <= DEG; ++k) { p1 = fma(p1, t1, C[k]); p2 = fma(p2, t2, C[k]); }   // operands straight from the constant bank
  e1 = p1 * r1;
  e2 = p2 * r2;
}

// lower: g = ln b - ln beta = ln(b / beta), one logarithm per evaluation: inv_beta = 1/beta, lnbeta_off = 0
// (for beta below 1e-290, where 1/beta is not finite, inv_beta = 1 and lnbeta_off = ln beta)
__device__ __noinline__ double polish_f64(double x, double beta, double inv_beta, double lnbeta_off, double ex2,
                                             bool lower, double s, int& iters) {
  double lo = 0.0, hi = S_MAX;
  const double tvc = ex2 - beta;                     // > 0: what the price lacks to its s -> inf limit
  bool done = false;
  int it = 0;
  while (true) {
    if (!done) {
      const double inv_s = rcp_fast(s);              // s >= 1e-45 2^-64: normal
      const double h = -x * inv_s, t = 0.5 * s;      // h >= 0
      const double r2 = h * h * inv_s - 0.25 * s;
      const double e = 0.5 * (h * h + t * t);
      double g, newton, den;
      if (lower) {
        const double q1 = (h - t) * INV_SQRT2, q2 = (h + t) * INV_SQRT2;
        double e1, e2;
        if (q1 >= 0.0 && q2 < 1e70) erfcx_pair_pos(q1, q2, e1, e2);
        else { e1 = erfcx(q1); e2 = erfcx(q2); }      // s overshot the inflection point (q1 < 0), or absurd inputs
        const double D = e1 - e2;
        const double arg = 0.5 * D * inv_beta;
        // ln(b/beta) with b/beta = arg e^{-e} = 1 + dl: after the pre-solve dl is ~1e-5, and ln(1+dl) is its series
        // (dl^6/6 < 1e-17 below 2e-3) — an exponential instead of a logarithm, and no cancellation against ln beta
        const double dl = fma(arg, exp_neg(e), -1.0);
        if (fabs(dl) < 2e-3 && lnbeta_off == 0.0)
          g = dl * fma(dl, fma(dl, fma(dl, fma(dl, 0.2, -0.25), 1.0 / 3.0), -0.5), 1.0);
        else
          g = log(arg) - e - lnbeta_off;
        newton = -g * D * (0.5 * SQRT_2PI);
        den = 1.0 + 0.5 * newton * r2 + 0.5 * g;
        if (!(den > 0.5 && den < 2.0)) den = 1.0;
      } else {
        // above the inflection point t > h:  Phi(t-h) = 1 - erfc((t-h)/sqrt2)/2, and both e^{+-x/2} e^{-q^2} collapse to
        // E = e^{-e}  =>  b = e^{x/2} - E/2 [erfcx((t-h)/sqrt2) + erfcx((t+h)/sqrt2)]: one exp, no library erfc
        const double q1 = (t - h) * INV_SQRT2, q2 = (h + t) * INV_SQRT2;
        if (q1 >= 0.0 && q2 < 1e70) {
          double e1, e2;
          erfcx_pair_pos(q1, q2, e1, e2);
          const double E = exp_neg(e);
          g = tvc - 0.5 * E * (e1 + e2);
          newton = -g * SQRT_2PI;                      // times 1/E, folded into the one division below
          den = fma(0.5 * newton, r2, E);
          if (!(den > 0.5 * E && den < 2.0 * E)) den = E;
                  } else {                                       // s undershot the inflection point
          const double b = 0.5 * (ex2 * erfc(-q1) - erfc(q2) / ex2);
          g = b - beta;
          newton = -g * SQRT_2PI * exp(e);
          den = 1.0 + 0.5 * newton * r2;
          if (!(den > 0.5 && den < 2.0)) den = 1.0;
        }
      }
      ++it;
      if (g > 0.0) hi = s; else lo = s;
      double step = newton * rcp_fast1(den);         // the step is ~1e-5 s: 2^-40 relative is ample; a flushed den gives NaN -> bisection
      double sn = s + step;
      const bool safe = (sn >= lo) && (sn <= hi) && (sn > 0.0);          // false for NaN
      if (!safe) { sn = 0.5 * (lo + hi); step = sn - s; }
      if ((safe && fabs(step) <= ACCEPT_STEP * sn) || !(hi - lo > 4.0e-16 * hi) || it >= MAX_ITERS) done = true;
      s = sn;
    }
    if (__all_sync(__activemask(), done)) break;
  }
  iters = it;
  return s;
}

// One quote, given Kd = K e^{-rT}.  `ok` says the inputs are finite and positive; every lane of the warp must call
// (the polish loop votes), lanes that are not ok solve a dummy alongside.  PRE_X: x_pre = ln(S/Kd) was formed
// from precomputed logarithms (chain layout) and the fp64 log is skipped.
// DEFER: a quote the fast path does not accept is NOT polished here — the function returns true, the caller records the
// warp, and the retry kernel (the same function with DEFER = false) re-solves the recorded warps.  The fast kernels then
// contain no call, no loop and no vote.
template <bool PRE_X, bool DEFER>
__device__ __forceinline__ bool solve_quote(bool ok, double p, double s_, double kd, double t_, bool call, double x_pre,
                                            double& sigma_out, int& iters_out) {
  // ok: S, Kd and T are positive normal numbers (no NaN below).  mx/mn = max/min(S, Kd); the in-the-money side is the
  // call when S > Kd, the put when S < Kd, and its intrinsic value is mx - mn.
  const bool sg = s_ > kd;
  const double mx = sg ? s_ : kd, mn = sg ? kd : s_;     // garbage in, garbage through: replaced by a dummy quote below
  const double diff = mx - mn;
  const bool itm = (call == sg);
  const double lower_bound = itm ? diff : 0.0;
  const double target = p - lower_bound;                            // the out-of-the-money twin (p - 0 is p; p - diff is exact near the bound)
  // the upper bound (call: p < S, put: p < Kd) is target < min(S, Kd) for the twin, i.e. beta < e^{x/2}: tested once, below
  ok = ok && (target > 0.0);                                        // false for NaN prices
  // x = -|ln(S/Kd)| = ln(min/max), e^{x/2} = sqrt(min/max) = min / sqrt(S Kd), beta = target / sqrt(S Kd): ONE reciprocal square root
  // (of the product S Kd, which must therefore be a normal double itself) serves e^{x/2} and beta, and x/2 = ln e^{x/2} needs no
  // reciprocal of its own.  hx = x/2 is what the fast path carries (h = -x/s = -hx (2/s), the doubling on the exponent).
  const double prod = mn * mx;
  if (!pos_normal(prod)) ok = false;                               // S Kd outside the normal range: an integer test
  const double rs = rsqrt_fast(prod);
  double ex2 = mn * rs;                                            // (0, 1]
  double beta = target * rs;
  double hx;
  if (PRE_X) hx = -0.5 * fabs(x_pre);
  else {                                                           // e^{x/2} = 1 gives +-1e-18 from the table: positive results become 0 (sign-bit mask)
    const double lx = log_normal(ex2);
    const int keep = __double2hiint(lx) >> 31;
    hx = __hiloint2double(__double2hiint(lx) & keep, __double2loint(lx) & keep);
  }
  if (ok && !(beta < ex2)) ok = false;                             // rounding pushed it onto the bound
  // lanes that are not ok: with the bracketed loop in line (DEFER = false) they solve a dummy alongside, since the loop votes
  // and must terminate; the straight-line fast kernel needs nothing of the kind — NaNs flow through (the table index clamps)
  if (!DEFER && !ok) { hx = -0.05268025782891314; ex2 = 0.9486832980505138; beta = 0.05; }   // ln 0.9 / 2

  const float xf = 2.0f * (float)hx;
  // ln beta to single precision from the exponent and the leading mantissa bits (beta may be far below FLT_MIN)
  const int bhi = __double2hiint(beta);
  const float lnbeta_f = fmaf((float)(((bhi >> 20) & 0x7ff) - 1023), 0.69314718f,
                              f_log((float)__hiloint2double((bhi & 0x000fffff) | 0x3ff00000, __double2loint(beta))));
  const float s_tab = guess_table(xf, lnbeta_f);
  const bool lower = s_tab * s_tab < -2.0f * xf;                           // below the inflection point s_c = sqrt(2|x|)
#ifdef IVX_EXP_NOPRE
  float s0 = s_tab;
#else
  float s0 = presolve_f32(xf, (float)beta, lnbeta_f, (float)ex2, s_tab, lower);
#endif
  // fp32 range exhausted (or s so small that an argument of the erfcx pair would overflow): the bracketed fp64 loop takes it from anywhere
  const bool s_ok = s0 >= 1e-30f && s0 < (float)S_MAX;
  double s = s_ok ? (double)s0 : 1.0;
  // ---- fast path: ONE fp64 Halley step on b(s) - beta, the same expression on both sides of the inflection point:
  //   b = [h < t] e^{x/2} + E/2 (sgn(h-t) erfcx(|h-t|/sqrt2) - erfcx((h+t)/sqrt2)),  b' = E / sqrt(2 pi),  b''/b' = r2.
  // Accepted when the step is below 1e-4 s (cubic convergence: what is left is ~1e-12 s); anything else — a pre-solve
  // that landed far away, arguments outside the ranges of the fast helpers — goes to the bracketed loop below.
  int it = 1;
  bool accept;
  {
    // Halley: step = -g / (b' - g b''/(2 b')),  b' = E/sqrt(2 pi) =: Ec,  b''/b' = h^2/s - s/4
    const double inv_s = rcp_fast(s);
    // (1/s)/2 and 2 (1/s) by an exponent decrement / increment (1/40 < 1/s <= 1e30: normal)
    const double half_inv_s = __hiloint2double(__double2hiint(inv_s) - 0x00100000, __double2loint(inv_s));
    const double two_inv_s = __hiloint2double(__double2hiint(inv_s) + 0x00100000, __double2loint(inv_s));
    const double h = -hx * two_inv_s, t = 0.5 * s;   // h = -x/s >= 0
    const double hh = h * h;
    const double r2h = fma(hh, half_inv_s, -0.125 * s);              // b''/(2 b')
    const double q = fma(t, t, hh);                                  // h^2 + t^2
    const double d = h - t;
    double e1, e2;
    erfcx_pair_pos<true>(fabs(d) * KF_INV_SQRT2, (h + t) * KF_INV_SQRT2, e1, e2);   // both times sqrt(2 pi)/2
    const double Ec = exp_half_neg_over_sqrt2pi(q);  // relative error of E = relative error of b: 3e-13 is ample
    // sign and mask by integer operations on the high words (an fp64 compare, negation or select costs two issue cycles each):
    // e1 >= 0 takes the sign of d; ex2 is kept where d < 0 (h - t = 0 is +0: not above)
    const int dhi = __double2hiint(d), above_mask = dhi >> 31;
    const double e1s = __hiloint2double(__double2hiint(e1) | (dhi & (int)0x80000000), __double2loint(e1));
    const double cst = __hiloint2double(__double2hiint(ex2) & above_mask, __double2loint(ex2) & above_mask);
    const double g = fma(Ec, e1s - e2, cst - beta);                  // b(s) - beta
    const double den = fma(-g, r2h, Ec);
    const double step = -g * rcp_fast1(den);
    const double sn = s + step;
    // Accepted when |step| < 2^-13 s (high-word compare, unsigned so that a NaN step fails).  s >= 1e-30 keeps every argument of
    // the pair finite; q < 1400 keeps E normal (q >= 0, and a NaN q makes the step NaN).  With q < 1400, |s b''/b'| <= 1400, so a
    // step that small implies a Halley denominator within [Ec/2, 3 Ec/2] (|correction| <= 700 |newton|/s and
    // |1 + correction| >= 8192 |newton|/s cannot both hold above 1.4e-4): den > 0 is all that is left to check, on the sign bit.
    accept = ((unsigned)((__double2hiint(step) & 0x7fffffff) + (13 << 20)) < (unsigned)__double2hiint(s)) && s_ok &&
             (__double2hiint(q) < 0x4095E000) && (__double2hiint(den) > 0);
#ifdef IVX_EXP_FORCE
    accept = true;
#endif
    if (accept) s = sn;
  }
  if (!DEFER && !accept) {
    double inv_beta = 1.0, lnbeta_off = 0.0;
    if (beta < 1e-290) lnbeta_off = log(beta); else inv_beta = 1.0 / beta;
    int it2 = 0;
    s = polish_f64(hx + hx, beta, inv_beta, lnbeta_off, ex2, lower, s, it2);
    it += it2;
  }
  sigma_out = ok ? s * rsqrt_fast(t_) : qnan();
  iters_out = ok ? it : 0;
  return DEFER && ok && !accept;
}

// warps with a deferred quote: one entry each (the lane-0 atomic is rare: realistic tapes defer ~0 % of the quotes)
__device__ __forceinline__ void record_deferred(bool deferred, uint32_t warp_id, uint32_t* __restrict__ list, uint32_t* __restrict__ counter) {
  if (__any_sync(IVX_FULL_MASK, deferred) && (threadIdx.x & 31) == 0) list[atomicAdd(counter, 1u)] = warp_id;
}

// One launch covers at most 2^31 quotes (the wrapper offsets the pointers): a 32-bit quote index keeps the address
// arithmetic of the six loads to one IMAD.WIDE each.  ITERS = false: the iteration counts are not wanted (the hot path).
template <bool ITERS, bool DEFER>
__global__ void __launch_bounds__(256, 6) k_bs_implied_vol(uint32_t n, const double* __restrict__ price,
                                                        const double* __restrict__ S, const double* __restrict__ K,
                                                        const double* __restrict__ T, const double* __restrict__ r,
                                                        const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                        int32_t* __restrict__ iters, uint32_t* __restrict__ list,
                                                        uint32_t* __restrict__ counter) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  const bool inr = i < n;
  const uint32_t j = inr ? i : 0u;                   // n >= 1: lanes past the end re-read quote 0 and drop the result
  const double p = ld_stream(price + j), s_ = ld_stream(S + j), k_ = ld_stream(K + j);
  const double t_ = ld_stream(T + j), r_ = ld_stream(r + j);
  const bool call = is_call[j] != 0;
  const double kd = k_ * exp_neg_small(r_ * t_);
  // S, K, T and Kd positive, finite and normal (a NaN or infinite r shows in Kd); the price is checked against its bounds
  const bool ok = pos_normal(s_) && pos_normal(t_) && pos_normal(kd);   // (Kd normal and positive: so is K, or the quote is outside the domain anyway)
  double out;
  int it;
  const bool deferred = solve_quote<false, DEFER>(ok, p, s_, kd, t_, call, 0.0, out, it);
  if (inr) {
    st_stream(sigma + i, out);
    if (ITERS) st_stream(iters + i, it);
  }
  if (DEFER) record_deferred(deferred && inr, i >> 5, list, counter);
}

// the recorded warps again, with the bracketed loop in line: one warp per entry, grid-stride
template <bool ITERS>
__global__ void __launch_bounds__(256) k_bs_implied_vol_retry(uint32_t n, const double* __restrict__ price,
                                                              const double* __restrict__ S, const double* __restrict__ K,
                                                              const double* __restrict__ T, const double* __restrict__ r,
                                                              const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                              int32_t* __restrict__ iters, const uint32_t* __restrict__ list,
                                                              const uint32_t* __restrict__ counter) {
  const uint32_t count = *counter;
  for (uint32_t e = blockIdx.x * 8u + (threadIdx.x >> 5); e < count; e += gridDim.x * 8u) {
    const uint32_t i = list[e] * 32u + (threadIdx.x & 31u);
    const bool inr = i < n;
    const uint32_t j = inr ? i : 0u;
    const double p = price[j], s_ = S[j], k_ = K[j], t_ = T[j], r_ = r[j];
    const bool call = is_call[j] != 0;
    const double kd = k_ * exp_neg_small(r_ * t_);
    const bool ok = pos_normal(s_) && pos_normal(t_) && pos_normal(kd);   // (Kd normal and positive: so is K, or the quote is outside the domain anyway)
    double out;
    int it;
    solve_quote<false, false>(ok, p, s_, kd, t_, call, 0.0, out, it);
    if (inr) {
      sigma[i] = out;
      if (ITERS) iters[i] = it;
    }
  }
}

// ---- quotes that form slices on presumed ladders (the per-quote layout of the surface pipeline) -------------------------------
// Quote i = t * period + j; slice = i / n_strike; the strike of quote (t, j) is PRESUMED to equal x_ladder[j] (the first stamp's
// strikes).  The strike is in a register here anyway, so it is compared with the ladder's bit for bit on the way: a slice with a
// strike that differs is flagged (slice_flags[slice] = 1, *flag_count counts slices) and the strike spline, told so
// (x_verified), does not read 8 bytes per knot again to find out.  grid: x over the quotes of one stamp, y over stamps.
__global__ void __launch_bounds__(256, 6) k_bs_implied_vol_ladder(uint32_t period, uint32_t n_strike, const double* __restrict__ price,
                                                               const double* __restrict__ S, const double* __restrict__ K,
                                                               const double* __restrict__ T, const double* __restrict__ r,
                                                               const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                               const double* __restrict__ x_ladder, int32_t* __restrict__ slice_flags,
                                                               uint32_t* __restrict__ flag_count, uint32_t* __restrict__ list,
                                                               uint32_t* __restrict__ counter) {
  const uint32_t j = blockIdx.x * 256u + threadIdx.x;
  const bool inr = j < period;
  const uint32_t jj = inr ? j : 0u;
  const uint32_t i = blockIdx.y * period + jj;       // < 2^31: the wrapper chunks the stamps
  const double p = ld_stream(price + i), s_ = ld_stream(S + i), k_ = ld_stream(K + i);
  const double t_ = ld_stream(T + i), r_ = ld_stream(r + i);
  const bool call = is_call[i] != 0;
  const double lad = __ldcg(x_ladder + jj);         // L2 only: the 480 KB of ladders must not push the guess table out of L1
  const double kd = k_ * exp_neg_small(r_ * t_);
  const bool ok = pos_normal(s_) && pos_normal(t_) && pos_normal(kd);   // (Kd normal and positive: so is K, or the quote is outside the domain anyway)
  double out;
  int it;
  const bool deferred = solve_quote<false, true>(ok, p, s_, kd, t_, call, 0.0, out, it);
  if (inr) {
    st_stream(sigma + i, out);
    if (((__double2hiint(k_) ^ __double2hiint(lad)) | (__double2loint(k_) ^ __double2loint(lad))) != 0) {   // rare
      // one atomic per (warp, slice) at most, none once the slice is flagged: a moved ladder makes all 2000 quotes of a slice differ
      const uint32_t slice = i / n_strike;
      const uint32_t peers = __match_any_sync(__activemask(), slice);
      if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1) && *((volatile int32_t*)(slice_flags + slice)) == 0 &&
          atomicExch(slice_flags + slice, 1) == 0)
        atomicAdd(flag_count, 1u);
    }
  }
  record_deferred(deferred && inr, (blockIdx.y * gridDim.x + blockIdx.x) * 8u + (threadIdx.x >> 5), list, counter);
}

__global__ void __launch_bounds__(256) k_bs_implied_vol_ladder_retry(uint32_t gx, uint32_t period, const double* __restrict__ price,
                                                                     const double* __restrict__ S, const double* __restrict__ K,
                                                                     const double* __restrict__ T, const double* __restrict__ r,
                                                                     const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                                     const uint32_t* __restrict__ list, const uint32_t* __restrict__ counter) {
  const uint32_t count = *counter;
  for (uint32_t e = blockIdx.x * 8u + (threadIdx.x >> 5); e < count; e += gridDim.x * 8u) {
    const uint32_t w = list[e];
    const uint32_t blk = w >> 3;                                   // (y * gx + x)
    const uint32_t j = (blk % gx) * 256u + (w & 7u) * 32u + (threadIdx.x & 31u);
    const bool inr = j < period;
    const uint32_t i = (blk / gx) * period + (inr ? j : 0u);
    const double p = price[i], s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i];
    const bool call = is_call[i] != 0;
    const double kd = k_ * exp_neg_small(r_ * t_);
    const bool ok = pos_normal(s_) && pos_normal(t_) && pos_normal(kd);   // (Kd normal and positive: so is K, or the quote is outside the domain anyway)
    double out;
    int it;
    solve_quote<false, false>(ok, p, s_, kd, t_, call, 0.0, out, it);
    if (inr) sigma[i] = out;
  }
}


// ---- two quotes per thread -------------------------------------------------------------------------------------------
// The fast kernel is bound by instruction issue (profiles/README.md), and ~65 of its ~400 instructions per quote do nothing
// quote-specific: 46 constant loads (B200 has no constant operand for fp64: every coefficient comes through LDCU/LDC), the
// register moves that assemble immediates, the kernel-parameter loads and the index arithmetic.  A thread that solves TWO
// quotes side by side — D2 below: every fp64 statement of the set-up and of the Halley step written once, executed for both
// components back to back — pays those once per pair.  The fp32 pre-solve has no constants to share (its coefficients are
// immediates) and branches per quote: it stays scalar, called twice.
// Measured (ncu): 370 instead of 406 warp instructions per 32 quotes, 121 instead of 133 of them fp64 — but at 64 registers
// only 32 warps per SM stay resident and the two dependency chains of a thread do not hide each other's latencies as well as
// two warps did: issue slots 69-72 % busy instead of 77 %, 19.4 against 19.6 ms per day.  It ships because it is never slower.
struct D2 { double a, b; };
using ::fma;                                     // the overloads below must not hide the scalar one
__device__ __forceinline__ D2 operator+(D2 x, D2 y) { return {x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ D2 operator-(D2 x, D2 y) { return {x.a - y.a, x.b - y.b}; }
__device__ __forceinline__ D2 operator*(D2 x, D2 y) { return {x.a * y.a, x.b * y.b}; }
__device__ __forceinline__ D2 operator+(D2 x, double y) { return {x.a + y, x.b + y}; }
__device__ __forceinline__ D2 operator*(D2 x, double y) { return {x.a * y, x.b * y}; }
__device__ __forceinline__ D2 operator*(double x, D2 y) { return {x * y.a, x * y.b}; }
__device__ __forceinline__ D2 operator-(D2 x) { return {-x.a, -x.b}; }
__device__ __forceinline__ D2 fma(D2 x, D2 y, D2 z) { return {fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)}; }
__device__ __forceinline__ D2 fma(D2 x, D2 y, double z) { return {fma(x.a, y.a, z), fma(x.b, y.b, z)}; }
__device__ __forceinline__ D2 fma(D2 x, double y, D2 z) { return {fma(x.a, y, z.a), fma(x.b, y, z.b)}; }
__device__ __forceinline__ D2 fma(D2 x, double y, double z) { return {fma(x.a, y, z), fma(x.b, y, z)}; }
__device__ __forceinline__ D2 fma(double x, D2 y, double z) { return {fma(x, y.a, z), fma(x, y.b, z)}; }
template <class F> __device__ __forceinline__ D2 map2(D2 x, F f) { return {f(x.a), f(x.b)}; }
__device__ __forceinline__ D2 add_exp(D2 x, int k) {           // x 2^k on the exponent field (x normal, no overflow: see the callers)
  return {__hiloint2double(__double2hiint(x.a) + (k << 20), __double2loint(x.a)), __hiloint2double(__double2hiint(x.b) + (k << 20), __double2loint(x.b))};
}
__device__ __forceinline__ D2 rcp_fast(D2 a) {
  const D2 y = map2(a, [](double v) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); return r; });
  const D2 e = fma(-a, y, 1.0);
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ D2 rcp_fast1(D2 a) {
  const D2 y = map2(a, [](double v) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); return r; });
  return fma(y, fma(-a, y, 1.0), y);
}
__device__ __forceinline__ D2 rsqrt_fast(D2 a) {
  const D2 y = map2(a, [](double v) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); return r; });
  const D2 e = fma(-(a * y), y, 1.0);
  return fma(y, e * fma(e, 0.375, 0.5), y);
}
__device__ __forceinline__ D2 log_normal(D2 a) {              // see log_normal(double)
  const int hia = __double2hiint(a.a), hib = __double2hiint(a.b);
  const D2 m = {__hiloint2double((hia & 0x000fffff) | 0x3ff00000, __double2loint(a.a)), __hiloint2double((hib & 0x000fffff) | 0x3ff00000, __double2loint(a.b))};
  const double2 ta = __ldg(&LOG_TAB[(hia >> 14) & 63]), tb = __ldg(&LOG_TAB[(hib >> 14) & 63]);
  const D2 u = fma(m, D2{ta.x, tb.x}, -1.0);
  D2 p = {LOG1P_C[1], LOG1P_C[1]};
#pragma unroll
  for (int k = 2; k < 7; ++k) p = fma(p, u, LOG1P_C[k]);
  const D2 l = fma(u * u, p, u);
  const D2 ed = {(double)((hia >> 20) - 1023), (double)((hib >> 20) - 1023)};
  return fma(ed, KF_LN2_HI, D2{ta.y, tb.y} + fma(ed, KF_LN2_LO, l));
}
__device__ __forceinline__ D2 exp_half_neg_over_sqrt2pi(D2 q) {   // see the scalar version
  const double MAGIC = 6755399441055744.0;
  const D2 tn = fma(q, EXPH_C[10], MAGIC);
  const int na = __double2loint(tn.a), nb = __double2loint(tn.b);
  const D2 nd = tn + (-MAGIC);
  const D2 w = fma(nd, EXPH_C[11], -q);
  D2 p = {EXPH_C[0], EXPH_C[0]};
#pragma unroll
  for (int k = 1; k < 10; ++k) p = fma(p, w, EXPH_C[k]);
  p = fma(p, w, 1.0);
  return p * D2{__hiloint2double(((1021 - na) << 20) | 0x98845, 0x33d43651), __hiloint2double(((1021 - nb) << 20) | 0x98845, 0x33d43651)};
}
// four interleaved Horner chains on one stream of coefficients (the fast path's scaled degree-18 fit)
__device__ __forceinline__ void erfcx_pair_pos_scaled(D2 q1, D2 q2, D2& e1, D2& e2) {
  const D2 a1 = q1 + 4.0, a2 = q2 + 4.0;
  const D2 r = rcp_fast(a1 * a2);
  const D2 r1 = r * a2, r2 = r * a1;
  const D2 t1 = fma(r1, -8.0, 1.0), t2 = fma(r2, -8.0, 1.0);
  D2 p1 = {ERFCX_CS[0], ERFCX_CS[0]}, p2 = p1;
#pragma unroll
  for (int k = 1; k <= 18; ++k) { p1 = fma(p1, t1, ERFCX_CS[k]); p2 = fma(p2, t2, ERFCX_CS[k]); }
  e1 = p1 * r1;
  e2 = p2 * r2;
}

struct QuoteIn { double p, s_, k_, t_, r_; bool call; };
// The fast path of solve_quote<false, true> for two quotes (same arithmetic, statement for statement: results are bit-equal to the
// one-quote kernel's).  Returns the deferred flags.
__device__ __forceinline__ void solve_pair(const QuoteIn& A, const QuoteIn& B, double& sigA, double& sigB, bool& defA, bool& defB) {
  // discount factors: the common range of both by one interleaved polynomial, anything else per quote
  const D2 z = {A.r_ * A.t_, B.r_ * B.t_};
  D2 disc;
  if (((__double2hiint(z.a) & 0x7fffffff) < 0x3FB00000) && ((__double2hiint(z.b) & 0x7fffffff) < 0x3FB00000)) {
    const D2 w = -z;
    D2 p = {EXP_C[5], EXP_C[5]};
#pragma unroll
    for (int k = 6; k < 14; ++k) p = fma(p, w, EXP_C[k]);
    disc = p;
  } else {
    disc = {exp_neg_small(z.a), exp_neg_small(z.b)};
  }
  const D2 kd = D2{A.k_, B.k_} * disc;
  bool okA = pos_normal(A.s_) && pos_normal(A.t_) && pos_normal(kd.a), okB = pos_normal(B.s_) && pos_normal(B.t_) && pos_normal(kd.b);
  const bool sgA = A.s_ > kd.a, sgB = B.s_ > kd.b;
  const D2 mx = {sgA ? A.s_ : kd.a, sgB ? B.s_ : kd.b}, mn = {sgA ? kd.a : A.s_, sgB ? kd.b : B.s_};
  const D2 diff = mx - mn;
  const D2 target = D2{A.p, B.p} - D2{(A.call == sgA) ? diff.a : 0.0, (B.call == sgB) ? diff.b : 0.0};
  okA = okA && (target.a > 0.0); okB = okB && (target.b > 0.0);
  const D2 prod = mn * mx;
  okA = okA && pos_normal(prod.a); okB = okB && pos_normal(prod.b);
  const D2 rs = rsqrt_fast(prod);
  const D2 ex2 = mn * rs, beta = target * rs;
  const D2 lx = log_normal(ex2);
  const int keepA = __double2hiint(lx.a) >> 31, keepB = __double2hiint(lx.b) >> 31;
  const D2 hx = {__hiloint2double(__double2hiint(lx.a) & keepA, __double2loint(lx.a) & keepA),
                 __hiloint2double(__double2hiint(lx.b) & keepB, __double2loint(lx.b) & keepB)};
  if (okA && !(beta.a < ex2.a)) okA = false;
  if (okB && !(beta.b < ex2.b)) okB = false;
  // fp32 pre-solve: the table guesses are straight-line code (the scheduler interleaves the two quotes by itself); the Halley step
  // branches on the side of the inflection point — when both quotes sit on the same side (neighbours in a smile nearly always do)
  // the two steps are instantiated for that side back to back in ONE basic block, so that they interleave as well
  const float xfA = 2.0f * (float)hx.a, xfB = 2.0f * (float)hx.b;
  auto lnb = [](double beta_) -> float {
    const int bhi = __double2hiint(beta_);
    return fmaf((float)(((bhi >> 20) & 0x7ff) - 1023), 0.69314718f, f_log((float)__hiloint2double((bhi & 0x000fffff) | 0x3ff00000, __double2loint(beta_))));
  };
  const float lnbA = lnb(beta.a), lnbB = lnb(beta.b);
  const float stA = guess_table(xfA, lnbA), stB = guess_table(xfB, lnbB);
  const bool lowA = stA * stA < -2.0f * xfA, lowB = stB * stB < -2.0f * xfB;
  float s0A, s0B;
  if (lowA == lowB) {
    if (lowA) { s0A = presolve_f32_t<1>(xfA, (float)beta.a, lnbA, (float)ex2.a, stA, true); s0B = presolve_f32_t<1>(xfB, (float)beta.b, lnbB, (float)ex2.b, stB, true); }
    else { s0A = presolve_f32_t<0>(xfA, (float)beta.a, lnbA, (float)ex2.a, stA, false); s0B = presolve_f32_t<0>(xfB, (float)beta.b, lnbB, (float)ex2.b, stB, false); }
  } else {
    s0A = presolve_f32(xfA, (float)beta.a, lnbA, (float)ex2.a, stA, lowA);
    s0B = presolve_f32(xfB, (float)beta.b, lnbB, (float)ex2.b, stB, lowB);
  }
  const bool sokA = s0A >= 1e-30f && s0A < (float)S_MAX, sokB = s0B >= 1e-30f && s0B < (float)S_MAX;
  const D2 s = {sokA ? (double)s0A : 1.0, sokB ? (double)s0B : 1.0};
  // fp64 Halley step, both quotes
  const D2 inv_s = rcp_fast(s);
  const D2 half_inv_s = add_exp(inv_s, -1), two_inv_s = add_exp(inv_s, 1);
  const D2 h = -hx * two_inv_s, t = s * 0.5;
  const D2 hh = h * h;
  const D2 r2h = fma(hh, half_inv_s, s * -0.125);
  const D2 q = fma(t, t, hh);
  const D2 d = h - t;
  D2 e1, e2;
  erfcx_pair_pos_scaled(D2{fabs(d.a), fabs(d.b)} * KF_INV_SQRT2, (h + t) * KF_INV_SQRT2, e1, e2);
  const D2 Ec = exp_half_neg_over_sqrt2pi(q);
  const int dhiA = __double2hiint(d.a), dhiB = __double2hiint(d.b), amA = dhiA >> 31, amB = dhiB >> 31;
  const D2 e1s = {__hiloint2double(__double2hiint(e1.a) | (dhiA & (int)0x80000000), __double2loint(e1.a)),
                  __hiloint2double(__double2hiint(e1.b) | (dhiB & (int)0x80000000), __double2loint(e1.b))};
  const D2 cst = {__hiloint2double(__double2hiint(ex2.a) & amA, __double2loint(ex2.a) & amA),
                  __hiloint2double(__double2hiint(ex2.b) & amB, __double2loint(ex2.b) & amB)};
  const D2 g = fma(Ec, e1s - e2, cst - beta);
  const D2 den = fma(-g, r2h, Ec);
  const D2 step = -g * rcp_fast1(den);
  const D2 sn = s + step;
  const bool accA = ((unsigned)((__double2hiint(step.a) & 0x7fffffff) + (13 << 20)) < (unsigned)__double2hiint(s.a)) && sokA &&
                    (__double2hiint(q.a) < 0x4095E000) && (__double2hiint(den.a) > 0);
  const bool accB = ((unsigned)((__double2hiint(step.b) & 0x7fffffff) + (13 << 20)) < (unsigned)__double2hiint(s.b)) && sokB &&
                    (__double2hiint(q.b) < 0x4095E000) && (__double2hiint(den.b) > 0);
  const D2 sf = {accA ? sn.a : s.a, accB ? sn.b : s.b};
  const D2 sig = sf * rsqrt_fast(D2{A.t_, B.t_});
  sigA = okA ? sig.a : qnan();
  sigB = okB ? sig.b : qnan();
  defA = okA && !accA;
  defB = okB && !accB;
}

// the ladder kernel with two quotes per thread: a CTA takes 512 quotes of one stamp, thread t the quotes t and t + 256
__global__ void __launch_bounds__(256, 4) k_bs_implied_vol_ladder2(uint32_t period, uint32_t n_strike, const double* __restrict__ price,
                                                                const double* __restrict__ S, const double* __restrict__ K,
                                                                const double* __restrict__ T, const double* __restrict__ r,
                                                                const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                                const double* __restrict__ x_ladder, int32_t* __restrict__ slice_flags,
                                                                uint32_t* __restrict__ flag_count, uint32_t* __restrict__ list,
                                                                uint32_t* __restrict__ counter) {
  const uint32_t j0 = blockIdx.x * 512u + threadIdx.x, j1 = j0 + 256u;
  const bool in0 = j0 < period, in1 = j1 < period;
  const uint32_t jj0 = in0 ? j0 : 0u, jj1 = in1 ? j1 : 0u;
  const uint32_t base = blockIdx.y * period, i0 = base + jj0, i1 = base + jj1;
  QuoteIn A, B;
  A.p = ld_stream(price + i0); B.p = ld_stream(price + i1);
  A.s_ = ld_stream(S + i0); B.s_ = ld_stream(S + i1);
  A.k_ = ld_stream(K + i0); B.k_ = ld_stream(K + i1);
  A.t_ = ld_stream(T + i0); B.t_ = ld_stream(T + i1);
  A.r_ = ld_stream(r + i0); B.r_ = ld_stream(r + i1);
  A.call = is_call[i0] != 0; B.call = is_call[i1] != 0;
  const double lad0 = __ldcg(x_ladder + jj0), lad1 = __ldcg(x_ladder + jj1);
  double sg0, sg1;
  bool df0, df1;
  solve_pair(A, B, sg0, sg1, df0, df1);
  if (in0) st_stream(sigma + i0, sg0);
  if (in1) st_stream(sigma + i1, sg1);
  // strikes against the ladder: both quotes in one test, the flagging itself out of the way (rare)
  const bool mm0 = in0 && ((__double2hiint(A.k_) ^ __double2hiint(lad0)) | (__double2loint(A.k_) ^ __double2loint(lad0))) != 0;
  const bool mm1 = in1 && ((__double2hiint(B.k_) ^ __double2hiint(lad1)) | (__double2loint(B.k_) ^ __double2loint(lad1))) != 0;
  if (mm0 | mm1) {
    auto flag = [&](uint32_t i) {
      const uint32_t slice = i / n_strike;
      const uint32_t peers = __match_any_sync(__activemask(), slice);
      if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1) && *((volatile int32_t*)(slice_flags + slice)) == 0 &&
          atomicExch(slice_flags + slice, 1) == 0)
        atomicAdd(flag_count, 1u);
    };
    if (mm0) flag(i0);
    if (mm1) flag(i1);
  }
  // warp ids of the retry list: ((stamp, 512-block) * 2 + half) * 8 + warp; one vote for both halves (deferrals are rare)
  df0 = df0 && in0; df1 = df1 && in1;
  if (__any_sync(IVX_FULL_MASK, df0 | df1)) {
    const uint32_t wbase = (blockIdx.y * gridDim.x + blockIdx.x) * 16u + (threadIdx.x >> 5);
    record_deferred(df0, wbase, list, counter);
    record_deferred(df1, wbase + 8u, list, counter);
  }
}

__global__ void __launch_bounds__(256) k_bs_implied_vol_ladder2_retry(uint32_t gx, uint32_t period, const double* __restrict__ price,
                                                                      const double* __restrict__ S, const double* __restrict__ K,
                                                                      const double* __restrict__ T, const double* __restrict__ r,
                                                                      const uint8_t* __restrict__ is_call, double* __restrict__ sigma,
                                                                      const uint32_t* __restrict__ list, const uint32_t* __restrict__ counter) {
  const uint32_t count = *counter;
  for (uint32_t e = blockIdx.x * 8u + (threadIdx.x >> 5); e < count; e += gridDim.x * 8u) {
    const uint32_t w = list[e];
    const uint32_t blk = w >> 4;                                   // (y * gx + x)
    const uint32_t j = (blk % gx) * 512u + ((w >> 3) & 1u) * 256u + (w & 7u) * 32u + (threadIdx.x & 31u);
    const bool inr = j < period;
    const uint32_t i = (blk / gx) * period + (inr ? j : 0u);
    const double p = price[i], s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i];
    const bool call = is_call[i] != 0;
    const double kd = k_ * exp_neg_small(r_ * t_);
    const bool ok = pos_normal(s_) && pos_normal(t_) && pos_normal(kd);
    double out;
    int it;
    solve_quote<false, false>(ok, p, s_, kd, t_, call, 0.0, out, it);
    if (inr) sigma[i] = out;
  }
}

// ---- chain layout: quote (t, e, k) = price[(t n_exp + e) n_strike + k]; S per stamp, K per contract, T per (stamp, expiry)
// terms[0 .. n_t)            ln S[t]
// terms[n_t .. +n_ek)        ln K[e][k]
// terms[.. +n_te)            r[t] T[t][e]
// terms[.. +n_te)            e^{-r T}
__global__ void __launch_bounds__(256) k_chain_terms(int64_t n_t, int64_t n_exp, int64_t n_strike, const double* __restrict__ S_t,
                                                     const double* __restrict__ K_ek, const double* __restrict__ T_te,
                                                     const double* __restrict__ r_t, double* __restrict__ terms) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n_ek = n_exp * n_strike, n_te = n_t * n_exp;
  if (i < n_t) { terms[i] = log(S_t[i]); return; }
  int64_t j = i - n_t;
  if (j < n_ek) { terms[n_t + j] = log(K_ek[j]); return; }
  j -= n_ek;
  if (j < n_te) {
    const double rt = r_t[j / n_exp] * T_te[j];
    terms[n_t + n_ek + j] = rt;
    terms[n_t + n_ek + n_te + j] = exp(-rt);
  }
}

// one quote of the chain layout: loads + solve; (j, t) = contract and stamp
// TICKS: the price column holds int32 ticks of 1e-4 (the reference's DECIMAL(15,4) price precision: 4 B per quote over PCIe
// instead of 8), INT32_MIN = no quote.  ticks / 1e4 is formed correctly rounded (reciprocal + exact residual, as in the
// bridge's rounding), i.e. it is bit-equal to the double a host gets from round(price, 4).
template <bool DEFER, bool TICKS>
__device__ __forceinline__ bool chain_quote(int32_t j, int64_t t, int32_t n_exp, int32_t n_strike, int64_t n_t, const double* __restrict__ price,
                                            const uint8_t* __restrict__ is_call, int32_t call_per_quote, const double* __restrict__ S_t,
                                            const double* __restrict__ K_ek, const double* __restrict__ T_te, const double* __restrict__ terms,
                                            double* __restrict__ sigma, int32_t* __restrict__ iters) {
  const int32_t n_ek = n_exp * n_strike;
  const bool inr = j < n_ek;
  const int64_t i = t * n_ek + j;
  double p = 0, s_ = 1, k_ = 1, t_ = 1, disc = 1, xi = 0;
  bool call = true;
  if (inr) {
    const int32_t e = j / n_strike;
    const int64_t te = t * n_exp + e;
    const int64_t n_te = n_t * (int64_t)n_exp;
    if (TICKS) {
      const int32_t tk = __ldcs(reinterpret_cast<const int32_t*>(price) + i);
      const double a = (double)tk, q0 = a * 1e-4;
      p = tk == INT32_MIN ? qnan() : fma(fma(-1e4, q0, a), 1e-4, q0);
    } else {
      p = ld_stream(price + i);
    }
    call = (call_per_quote ? is_call[i] : __ldg(is_call + j)) != 0;
    s_ = __ldg(S_t + t); k_ = __ldg(K_ek + j); t_ = __ldg(T_te + te);
    const double lnS = __ldg(terms + t), lnK = __ldg(terms + n_t + j);
    const double rt = __ldg(terms + n_t + n_ek + te);
    disc = __ldg(terms + n_t + n_ek + n_te + te);
    xi = (lnS - lnK) + rt;                                         // ln(S / (K e^{-rT}))
  }
  const double kd = k_ * disc;
  const bool ok = inr && pos_normal(s_) && pos_normal(t_) && pos_normal(kd) && isfinite(xi);
  double out;
  int it;
  const bool deferred = solve_quote<true, DEFER>(ok, p, s_, kd, t_, call, xi, out, it);
  if (inr) {
    st_stream(sigma + i, out);
    if (iters) st_stream(iters + i, it);
  }
  return deferred;
}

// grid: x over the n_exp*n_strike contracts of one stamp, y over stamps
template <bool DEFER, bool TICKS>
__global__ void __launch_bounds__(256, 6) k_bs_implied_vol_chain(int64_t t0, int32_t n_exp, int32_t n_strike, int64_t n_t,
                                                              const double* __restrict__ price, const uint8_t* __restrict__ is_call,
                                                              int32_t call_per_quote, const double* __restrict__ S_t, const double* __restrict__ K_ek,
                                                              const double* __restrict__ T_te, const double* __restrict__ terms,
                                                              double* __restrict__ sigma, int32_t* __restrict__ iters,
                                                              uint32_t* __restrict__ list, uint32_t* __restrict__ counter) {
  const int32_t j = blockIdx.x * 256 + threadIdx.x;               // contract (e, k)
  const bool deferred = chain_quote<DEFER, TICKS>(j, t0 + blockIdx.y, n_exp, n_strike, n_t, price, is_call, call_per_quote, S_t, K_ek, T_te, terms, sigma, iters);
  if (DEFER) record_deferred(deferred, (blockIdx.y * gridDim.x + blockIdx.x) * 8u + (threadIdx.x >> 5), list, counter);
}

template <bool TICKS>
__global__ void __launch_bounds__(256) k_bs_implied_vol_chain_retry(int64_t t0, uint32_t gx, int32_t n_exp, int32_t n_strike, int64_t n_t,
                                                                    const double* __restrict__ price, const uint8_t* __restrict__ is_call,
                                                                    int32_t call_per_quote, const double* __restrict__ S_t, const double* __restrict__ K_ek,
                                                                    const double* __restrict__ T_te, const double* __restrict__ terms,
                                                                    double* __restrict__ sigma, int32_t* __restrict__ iters,
                                                                    const uint32_t* __restrict__ list, const uint32_t* __restrict__ counter) {
  const uint32_t count = *counter;
  for (uint32_t e = blockIdx.x * 8u + (threadIdx.x >> 5); e < count; e += gridDim.x * 8u) {
    const uint32_t w = list[e];
    const uint32_t blk = w >> 3;                                   // (y * gx + x)
    const int32_t j = (int32_t)((blk % gx) * 256u + (w & 7u) * 32u + (threadIdx.x & 31u));
    chain_quote<false, TICKS>(j, t0 + blk / gx, n_exp, n_strike, n_t, price, is_call, call_per_quote, S_t, K_ek, T_te, terms, sigma, iters);
  }
}

__global__ void __launch_bounds__(256) k_bs_price(int64_t n, const double* __restrict__ S, const double* __restrict__ K,
                                                  const double* __restrict__ T, const double* __restrict__ r,
                                                  const double* __restrict__ sigma, const uint8_t* __restrict__ is_call,
                                                  double* __restrict__ price) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i], v = sigma[i];
  const double sq = sqrt(t_);
  const double d1 = (log(s_ / k_) + (r_ + 0.5 * v * v) * t_) / (v * sq);   // greeks.py:21
  const double d2 = d1 - v * sq;                                           // greeks.py:22
  const double kd = k_ * exp(-r_ * t_);
  price[i] = is_call[i] ? s_ * norm_cdf(d1) - kd * norm_cdf(d2) : kd * norm_cdf(-d2) - s_ * norm_cdf(-d1);
}

__global__ void __launch_bounds__(256) k_bs_greeks(int64_t n, const double* __restrict__ S, const double* __restrict__ K,
                                                   const double* __restrict__ T, const double* __restrict__ r,
                                                   const double* __restrict__ sigma, const uint8_t* __restrict__ is_call,
                                                   int32_t ref_put_rho, double* delta, double* gamma, double* theta,
                                                   double* vega, double* rho) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s_ = S[i], k_ = K[i], t_ = T[i], r_ = r[i], v = sigma[i];
  const bool call = is_call[i] != 0;
  const double sq = sqrt(t_);
  const double d1 = (log(s_ / k_) + (r_ + 0.5 * v * v) * t_) / (v * sq);   // greeks.py:21
  const double d2 = d1 - v * sq;                                           // greeks.py:22
  const double pdf = norm_pdf(d1);
  const double disc = exp(-r_ * t_);
  if (delta) delta[i] = call ? norm_cdf(d1) : norm_cdf(d1) - 1.0;          // :25,:29
  if (theta)
    theta[i] = call ? (-s_ * pdf * v / (2.0 * sq) - r_ * k_ * disc * norm_cdf(d2)) / 365.0    // :26-27
                    : (-s_ * pdf * v / (2.0 * sq) + r_ * k_ * disc * norm_cdf(-d2)) / 365.0;  // :30-31
  if (gamma) gamma[i] = pdf / (s_ * v * sq);                               // :33
  if (vega) vega[i] = s_ * pdf * sq / 100.0;                               // :34
  if (rho) {
    double rr = k_ * t_ * disc * norm_cdf(call ? d2 : -d2) / 100.0;        // :35 (put sign as shipped)
    if (!call && !ref_put_rho) rr = -rr;
    rho[i] = rr;
  }
}

// ---- roofline probes ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probe_copy(int64_t n2, const double2* __restrict__ src, double2* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// kernel A's store pattern with nothing else: tiles of 1024 rows, 11 x 8-byte + 4-byte + 1-byte streams per row
__global__ void __launch_bounds__(256) k_probe_write_streams(int64_t n_rows, double* const base, int64_t stride_f64) {
  const int64_t r0 = blockIdx.x * (int64_t)1024;
  for (int it = 0; it < 4; ++it) {
    const int64_t o = r0 + it * 256 + threadIdx.x;
    if (o >= n_rows) return;
#pragma unroll
    for (int k = 0; k < 11; ++k) __stcs(base + k * stride_f64 + o, (double)o);
    __stcs(reinterpret_cast<int32_t*>(base + 11 * stride_f64) + o, (int32_t)o);
    reinterpret_cast<uint8_t*>(base + 12 * stride_f64)[o] = 0;
  }
}

__global__ void __launch_bounds__(256) k_probe_fp64(int32_t iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int k = 0; k < iters; ++k) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double t = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (t == 123.456) sink[0] = t;   // never true; keeps the chain alive
}

}  // namespace
}  // namespace ivx

using namespace ivx;

// workspace of the two-kernel form: a counter (256 B) + one uint32 per warp of the largest launch
extern "C" int64_t ivx_bs_implied_vol_workspace_bytes(int64_t n) {
  if (n < 0) return -1;
  const int64_t CH = (int64_t)1 << 31;
  const int64_t m = n < CH ? n : CH;
  return 256 + 4 * ((m + 31) / 32 + 8);
}

extern "C" int ivx_bs_implied_vol(int64_t n, const double* price, const double* S, const double* K, const double* T,
                                  const double* r, const uint8_t* is_call, double* sigma, int32_t* iters,
                                  void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(price && S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol: null pointer");
  IVX_REQUIRE(!workspace || workspace_bytes >= ivx_bs_implied_vol_workspace_bytes(n), IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol: workspace too small");
  uint32_t* counter = (uint32_t*)workspace;
  uint32_t* list = workspace ? (uint32_t*)((char*)workspace + 256) : nullptr;
  const int64_t CH = (int64_t)1 << 31;
  for (int64_t o = 0; o < n; o += CH) {
    const uint32_t m = (uint32_t)(n - o < CH ? n - o : CH);
    const unsigned blocks = (unsigned)(((int64_t)m + 255) / 256);
    int32_t* it = iters ? iters + o : nullptr;
#define IVX_IV_ARGS m, price + o, S + o, K + o, T + o, r + o, is_call + o, sigma + o, it, list, counter
    if (workspace) {
      // fast kernel (no fallback code in it) + retry kernel over the warps it recorded; stream order makes the pair one operation
      IVX_CUDA(cudaMemsetAsync(counter, 0, 4, stream));
      if (iters) k_bs_implied_vol<true, true><<<blocks, 256, 0, stream>>>(IVX_IV_ARGS);
      else k_bs_implied_vol<false, true><<<blocks, 256, 0, stream>>>(IVX_IV_ARGS);
      IVX_LAUNCH_CHECK("k_bs_implied_vol");
      const unsigned rb = (unsigned)(sm_count() * 2);
      if (iters) k_bs_implied_vol_retry<true><<<rb, 256, 0, stream>>>(IVX_IV_ARGS);
      else k_bs_implied_vol_retry<false><<<rb, 256, 0, stream>>>(IVX_IV_ARGS);
      IVX_LAUNCH_CHECK("k_bs_implied_vol_retry");
    } else {
      if (iters) k_bs_implied_vol<true, false><<<blocks, 256, 0, stream>>>(IVX_IV_ARGS);
      else k_bs_implied_vol<false, false><<<blocks, 256, 0, stream>>>(IVX_IV_ARGS);
      IVX_LAUNCH_CHECK("k_bs_implied_vol");
    }
#undef IVX_IV_ARGS
  }
  return IVX_OK;
}

// stamps per launch of the ladder form: a 32-bit quote index and a 16-bit grid dimension
static int64_t ladder_chunk_stamps(int64_t period) {
  int64_t c = (((int64_t)1 << 31) - 1) / period;
  return c > 65535 ? 65535 : c;
}

extern "C" int64_t ivx_bs_implied_vol_ladder_workspace_bytes(int64_t n_t, int64_t period) {
  if (n_t < 0 || period <= 0 || period >= ((int64_t)1 << 31)) return -1;
  const int64_t gx = (period + 255) / 256, ch = ladder_chunk_stamps(period), gy = n_t < ch ? n_t : ch;
  return 256 + 4 * (gx * gy * 8 + 16 * gy + 8);          // (the two-quote kernel numbers 16 warps per 512-quote block)
}

extern "C" int ivx_bs_implied_vol_ladder(int64_t n_t, int64_t period, int64_t n_strike, const double* price, const double* S, const double* K,
                                         const double* T, const double* r, const uint8_t* is_call, double* sigma, const double* x_ladder,
                                         int32_t* slice_flags, uint32_t* flag_count, void* workspace, int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_t >= 0 && period > 0 && period < ((int64_t)1 << 31) && n_strike > 0 && period % n_strike == 0, IVX_ERR_INVALID_ARG,
              "ivx_bs_implied_vol_ladder: need n_t >= 0 and a period that is a positive multiple of n_strike");
  if (n_t == 0) return IVX_OK;
  IVX_REQUIRE(price && S && K && T && r && is_call && sigma && x_ladder && slice_flags && flag_count && workspace, IVX_ERR_INVALID_ARG,
              "ivx_bs_implied_vol_ladder: null pointer");
  IVX_REQUIRE(workspace_bytes >= ivx_bs_implied_vol_ladder_workspace_bytes(n_t, period), IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol_ladder: workspace too small");
  uint32_t* counter = (uint32_t*)workspace;
  uint32_t* list = (uint32_t*)((char*)workspace + 256);
  IVX_CUDA(cudaMemsetAsync(slice_flags, 0, (size_t)(n_t * (period / n_strike)) * sizeof(int32_t), stream));
  IVX_CUDA(cudaMemsetAsync(flag_count, 0, 4, stream));
  const unsigned gx = (unsigned)((period + 255) / 256), gx2 = (unsigned)((period + 511) / 512);
  const int64_t ch = ladder_chunk_stamps(period);
  // IVX_IV_ONE_QUOTE=1 in the environment keeps the one-quote-per-thread kernel (A/B timing of the two forms; read once; same results)
  static const bool one_quote = [] { const char* e = getenv("IVX_IV_ONE_QUOTE"); return e && e[0] == '1'; }();
  for (int64_t t0 = 0; t0 < n_t; t0 += ch) {
    const unsigned gy = (unsigned)(n_t - t0 < ch ? n_t - t0 : ch);
    const int64_t o = t0 * period;
    IVX_CUDA(cudaMemsetAsync(counter, 0, 4, stream));
    if (!one_quote) {
      k_bs_implied_vol_ladder2<<<dim3(gx2, gy), 256, 0, stream>>>((uint32_t)period, (uint32_t)n_strike, price + o, S + o, K + o, T + o, r + o, is_call + o,
                                                                sigma + o, x_ladder, slice_flags + t0 * (period / n_strike), flag_count, list, counter);
      IVX_LAUNCH_CHECK("k_bs_implied_vol_ladder2");
      k_bs_implied_vol_ladder2_retry<<<(unsigned)(sm_count() * 2), 256, 0, stream>>>(gx2, (uint32_t)period, price + o, S + o, K + o, T + o, r + o, is_call + o,
                                                                                    sigma + o, list, counter);
      IVX_LAUNCH_CHECK("k_bs_implied_vol_ladder2_retry");
      continue;
    }
    k_bs_implied_vol_ladder<<<dim3(gx, gy), 256, 0, stream>>>((uint32_t)period, (uint32_t)n_strike, price + o, S + o, K + o, T + o, r + o, is_call + o,
                                                             sigma + o, x_ladder, slice_flags + t0 * (period / n_strike), flag_count, list, counter);
    IVX_LAUNCH_CHECK("k_bs_implied_vol_ladder");
    k_bs_implied_vol_ladder_retry<<<(unsigned)(sm_count() * 2), 256, 0, stream>>>(gx, (uint32_t)period, price + o, S + o, K + o, T + o, r + o, is_call + o,
                                                                                 sigma + o, list, counter);
    IVX_LAUNCH_CHECK("k_bs_implied_vol_ladder_retry");
  }
  return IVX_OK;
}

// terms (8 B each: n_t + n_ek + 2 n_te) + counter + one uint32 per warp of a launch (at most 65535 stamps per launch)
extern "C" int64_t ivx_bs_chain_workspace_bytes(int64_t n_t, int64_t n_exp, int64_t n_strike) {
  if (n_t < 0 || n_exp < 0 || n_strike < 0) return -1;
  const int64_t gx = (n_exp * n_strike + 255) / 256, gy = n_t < 65535 ? n_t : 65535;
  return 8 * (n_t + n_exp * n_strike + 2 * n_t * n_exp) + 256 + 256 + 4 * (gx * gy * 8 + 8);
}

static int bs_chain_launch(bool ticks, int64_t n_t, int64_t n_exp, int64_t n_strike, const double* price,
                           const uint8_t* is_call, int32_t is_call_per_quote, const double* S_t,
                           const double* K_ek, const double* T_te,
                           const double* r_t, double* sigma, int32_t* iters, void* workspace,
                           int64_t workspace_bytes, ivx_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IVX_REQUIRE(n_t >= 0 && n_exp >= 0 && n_strike >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol_chain: negative size");
  if (n_t == 0 || n_exp == 0 || n_strike == 0) return IVX_OK;
  IVX_REQUIRE(n_exp * n_strike < ((int64_t)1 << 24), IVX_ERR_TOO_LARGE, "ivx_bs_implied_vol_chain: more than 2^24 contracts per stamp");
  IVX_REQUIRE(price && is_call && S_t && K_ek && T_te && r_t && sigma && workspace, IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol_chain: null pointer");
  IVX_REQUIRE(workspace_bytes >= ivx_bs_chain_workspace_bytes(n_t, n_exp, n_strike), IVX_ERR_INVALID_ARG, "ivx_bs_implied_vol_chain: workspace too small");
  double* terms = (double*)workspace;
  const int64_t n_terms = n_t + n_exp * n_strike + n_t * n_exp;
  const int64_t terms_bytes = align_up(8 * (n_terms + n_t * n_exp), 256);
  uint32_t* counter = (uint32_t*)((char*)workspace + terms_bytes);
  uint32_t* list = counter + 64;
  k_chain_terms<<<(unsigned)((n_terms + 255) / 256), 256, 0, stream>>>(n_t, n_exp, n_strike, S_t, K_ek, T_te, r_t, terms);
  IVX_LAUNCH_CHECK("k_chain_terms");
  const unsigned gx = (unsigned)((n_exp * n_strike + 255) / 256);
  for (int64_t t0 = 0; t0 < n_t; t0 += 65535) {
    const unsigned gy = (unsigned)(n_t - t0 < 65535 ? n_t - t0 : 65535);
    IVX_CUDA(cudaMemsetAsync(counter, 0, 4, stream));
#define IVX_CH_ARGS (int32_t)n_exp, (int32_t)n_strike, n_t, price, is_call, is_call_per_quote, S_t, K_ek, T_te, terms, sigma, iters, list, counter
    if (ticks) k_bs_implied_vol_chain<true, true><<<dim3(gx, gy), 256, 0, stream>>>(t0, IVX_CH_ARGS);
    else k_bs_implied_vol_chain<true, false><<<dim3(gx, gy), 256, 0, stream>>>(t0, IVX_CH_ARGS);
    IVX_LAUNCH_CHECK("k_bs_implied_vol_chain");
    if (ticks) k_bs_implied_vol_chain_retry<true><<<(unsigned)(sm_count() * 2), 256, 0, stream>>>(t0, gx, IVX_CH_ARGS);
    else k_bs_implied_vol_chain_retry<false><<<(unsigned)(sm_count() * 2), 256, 0, stream>>>(t0, gx, IVX_CH_ARGS);
    IVX_LAUNCH_CHECK("k_bs_implied_vol_chain_retry");
#undef IVX_CH_ARGS
  }
  return IVX_OK;
}

extern "C" int ivx_bs_implied_vol_chain(int64_t n_t, int64_t n_exp, int64_t n_strike, const double* price,
                                        const uint8_t* is_call, int32_t is_call_per_quote, const double* S_t,
                                        const double* K_ek, const double* T_te,
                                        const double* r_t, double* sigma, int32_t* iters, void* workspace,
                                        int64_t workspace_bytes, ivx_stream_t stream) {
  return bs_chain_launch(false, n_t, n_exp, n_strike, price, is_call, is_call_per_quote, S_t, K_ek, T_te, r_t, sigma, iters, workspace, workspace_bytes, stream);
}

extern "C" int ivx_bs_implied_vol_chain_ticks(int64_t n_t, int64_t n_exp, int64_t n_strike, const int32_t* price_ticks,
                                              const uint8_t* is_call, int32_t is_call_per_quote, const double* S_t,
                                              const double* K_ek, const double* T_te,
                                              const double* r_t, double* sigma, int32_t* iters, void* workspace,
                                              int64_t workspace_bytes, ivx_stream_t stream) {
  return bs_chain_launch(true, n_t, n_exp, n_strike, reinterpret_cast<const double*>(price_ticks), is_call, is_call_per_quote, S_t, K_ek, T_te, r_t, sigma, iters,
                         workspace, workspace_bytes, stream);
}

extern "C" int ivx_bs_price(int64_t n, const double* S, const double* K, const double* T, const double* r,
                            const double* sigma, const uint8_t* is_call, double* price, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_price: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(price && S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_price: null pointer");
  k_bs_price<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, S, K, T, r, sigma, is_call, price);
  IVX_LAUNCH_CHECK("k_bs_price");
  return IVX_OK;
}

extern "C" int ivx_bs_greeks(int64_t n, const double* S, const double* K, const double* T, const double* r,
                             const double* sigma, const uint8_t* is_call, int32_t reference_put_rho_sign, double* delta,
                             double* gamma, double* theta, double* vega, double* rho, ivx_stream_t stream) {
  IVX_REQUIRE(n >= 0, IVX_ERR_INVALID_ARG, "ivx_bs_greeks: negative n");
  if (n == 0) return IVX_OK;
  IVX_REQUIRE(S && K && T && r && is_call && sigma, IVX_ERR_INVALID_ARG, "ivx_bs_greeks: null pointer");
  k_bs_greeks<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, S, K, T, r, sigma, is_call,
                                                                           reference_put_rho_sign, delta, gamma, theta, vega, rho);
  IVX_LAUNCH_CHECK("k_bs_greeks");
  return IVX_OK;
}

extern "C" int ivx_probe_copy(int64_t n_f64, const double* src, double* dst, ivx_stream_t stream) {
  IVX_REQUIRE(n_f64 >= 0 && (n_f64 % 2) == 0 && src && dst, IVX_ERR_INVALID_ARG, "ivx_probe_copy: need an even count and non-null buffers");
  IVX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0, IVX_ERR_INVALID_ARG, "ivx_probe_copy: buffers must be 16-byte aligned");
  k_probe_copy<<<sm_count() * 16, 256, 0, (cudaStream_t)stream>>>(n_f64 / 2, (const double2*)src, (double2*)dst);
  IVX_LAUNCH_CHECK("k_probe_copy");
  return IVX_OK;
}

extern "C" int ivx_probe_write_streams(int64_t n_rows, double* base, int64_t stride_f64, ivx_stream_t stream) {
  IVX_REQUIRE(n_rows > 0 && base && stride_f64 >= n_rows, IVX_ERR_INVALID_ARG, "ivx_probe_write_streams: bad arguments");
  k_probe_write_streams<<<(unsigned)((n_rows + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(n_rows, base, stride_f64);
  IVX_LAUNCH_CHECK("k_probe_write_streams");
  return IVX_OK;
}

extern "C" int ivx_probe_fp64_fma(int64_t n_threads, int32_t iters, double* sink, ivx_stream_t stream) {
  IVX_REQUIRE(n_threads > 0 && iters > 0 && sink, IVX_ERR_INVALID_ARG, "ivx_probe_fp64_fma: bad arguments");
  k_probe_fp64<<<(unsigned)((n_threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(iters, sink);
  IVX_LAUNCH_CHECK("k_probe_fp64");
  return IVX_OK;
}
