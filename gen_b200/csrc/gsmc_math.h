// gsmc_math.h -- fp64 elementary functions that give bit-identical results on the
// host (gcc, -ffp-contract=off) and on the device (nvcc, -fmad=false).
//
// Why: the particle filter's ancestor indices are a discontinuous function of
// the log weights. To make CPU-oracle <-> GPU parity bit-exact (not just
// "within 1e-5") every transcendental that feeds a log weight is built here
// from IEEE-754 operations only (+ - * / sqrt fma, integer bit casts, floor),
// which round identically on x86-64 and on sm_100a. Accuracy is <= 2 ulp
// against glibc (tests/test_math.py), far inside the 1e-5 bar that
// BASELINE.json sets against Julia's own libm.
//
// These replace, on the device, the libm calls the reference reaches through
// Julia Base: `log` in src/modeling_library/distributions/normal.jl:59,
// `exp`/`log` in src/inference/inference.jl:3-11, `randn` in normal.jl:96.
#ifndef GSMC_MATH_H
#define GSMC_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>
#include "gsmc_tables.h"

#if defined(__CUDACC__)
#define GM_HD __host__ __device__ __forceinline__
#else
#define GM_HD static inline
#endif

#define GM_PI 3.141592653589793115997963468544185161590576171875      /* Float64(pi) */
#define GM_RN_SHIFT 6755399441055744.0                               /* 1.5 * 2^52: x + SHIFT rounds x to an integer */
#define GM_TWO_PI 6.28318530717958623199592693708837032318115234375   /* 2.0*pi     */
#define GM_INF_BITS 0x7ff0000000000000ULL


// Polynomial coefficients live in __constant__ memory on the device so that DFMA takes them as
// constant-bank operands (as 64-bit immediates they cost two UMOV issue slots each, 16% of the
// propagate kernel's instructions in the first profile) and in a static table on the host.
#define GM_EXP_COEFFS { 8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5 }   /* 1/5! .. 1/2! */
#define GM_LOG_COEFFS { 4.7619047619047616e-02, 5.2631578947368418e-02, 5.8823529411764705e-02, 6.6666666666666666e-02, \
  7.6923076923076927e-02, 9.0909090909090912e-02, 1.1111111111111110e-01, 1.4285714285714285e-01, 2.0000000000000001e-01, \
  3.3333333333333331e-01 }
#define GM_SIN_COEFFS { -1.9841269841269841e-04, 8.3333333333333332e-03, -1.6666666666666666e-01 }   /* -1/7!, 1/5!, -1/3! */
#define GM_COS_COEFFS { -1.3888888888888889e-03, 4.1666666666666664e-02, -0.5 }                       /* -1/6!, 1/4!, -1/2! */
#define GM_ATAN_COEFFS { -6.6666666666666666e-02, 7.6923076923076927e-02, -9.0909090909090912e-02, 1.1111111111111110e-01, \
  -1.4285714285714285e-01, 2.0000000000000001e-01, -3.3333333333333331e-01 }
#define GM_MISC_CONSTS { 1.44269504088896338700e+00, 6.93147180369123816490e-01, 1.90821492927058770002e-10, \
  GM_EXP_INV_L, GM_EXP_L_HI, GM_EXP_L_LO }
static const double gm_exp_h[] = GM_EXP_COEFFS;
static const double gm_log_h[] = GM_LOG_COEFFS;
static const double gm_sin_h[] = GM_SIN_COEFFS;
static const double gm_cos_h[] = GM_COS_COEFFS;
static const double gm_atan_h[] = GM_ATAN_COEFFS;
static const double gm_misc_h[] = GM_MISC_CONSTS;
#if defined(__CUDACC__)
static __constant__ double gm_exp_d[] = GM_EXP_COEFFS;
static __constant__ double gm_log_d[] = GM_LOG_COEFFS;
static __constant__ double gm_sin_d[] = GM_SIN_COEFFS;
static __constant__ double gm_cos_d[] = GM_COS_COEFFS;
static __constant__ double gm_atan_d[] = GM_ATAN_COEFFS;
static __constant__ double gm_misc_d[] = GM_MISC_CONSTS;
#endif
#if defined(__CUDA_ARCH__)
#define GM_C(tab, i) gm_##tab##_d[i]
#else
#define GM_C(tab, i) gm_##tab##_h[i]
#endif
#define GM_INV_LN2 GM_C(misc, 0)
#define GM_LN2_HI GM_C(misc, 1)
#define GM_LN2_LO GM_C(misc, 2)
#define GM_INV_L64 GM_C(misc, 3)
#define GM_L64_HI GM_C(misc, 4)
#define GM_L64_LO GM_C(misc, 5)

// Lookup tables (gsmc_tables.h): 2^(j/64) and (sin, cos)(pi j/64). Lanes index them divergently, so on the
// device the default copies live in global memory (L1/L2-cached) and the hot kernels stage them in shared
// memory and pass that pointer to the batch forms.
static const double gm_exp2tab_h[64] = GM_EXP2TAB_VALUES;
static const double gm_sincostab_h[256] = GM_SINCOSTAB_VALUES;
#if defined(__CUDACC__)
static __device__ const double gm_exp2tab_g[64] = GM_EXP2TAB_VALUES;
static __device__ __align__(16) const double gm_sincostab_g[256] = GM_SINCOSTAB_VALUES;
#endif
#if defined(__CUDA_ARCH__)
#define GM_EXP2TAB gm_exp2tab_g
#define GM_SINCOSTAB gm_sincostab_g
#else
#define GM_EXP2TAB gm_exp2tab_h
#define GM_SINCOSTAB gm_sincostab_h
#endif

GM_HD double gm_from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double d; memcpy(&d, &b, 8); return d;
#endif
}
GM_HD uint64_t gm_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
/* the two adjacent table entries tab[2i], tab[2i+1] with ONE 16-byte load on the device (all pair tables are 16-byte aligned) */
GM_HD void gm_tab_pair(const double* tab, int i, double* a, double* b) {
#if defined(__CUDA_ARCH__)
  const double2 v = *reinterpret_cast<const double2*>(tab + 2 * i);
  *a = v.x; *b = v.y;
#else
  *a = tab[2 * i]; *b = tab[2 * i + 1];
#endif
}
/* (double)w for a 32-bit unsigned w, exactly, without an int->float conversion instruction: 2^52 + w has w in its low word */
GM_HD double gm_u32_to_f64(uint32_t w) { return gm_from_bits(0x4330000000000000ULL | (uint64_t)w) - 4503599627370496.0; }
/* (w + 0.5) 2^-32 for a 32-bit unsigned w with ONE subtraction: the bit pattern below is 2^20 + w 2^-32, and
   2^20 - 2^-33 is a double (53 significant bits); every step is exact, so the bits equal ((double)w + 0.5) * 2^-32 */
GM_HD double gm_u32_to_unit(uint32_t w) { return gm_from_bits(0x4130000000000000ULL | (uint64_t)w) - 0x1.fffffffffffffp+19; }
GM_HD double gm_inf(void) { return gm_from_bits(GM_INF_BITS); }
GM_HD double gm_nan(void) { return gm_from_bits(0x7ff8000000000000ULL); }
/* p * 2^k by an integer add on the exponent field (the caller guarantees a normal result) */
GM_HD double gm_scale_pow2(double p, int k) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  return gm_from_bits(gm_to_bits(p) + ((uint64_t)(long long)k << 52));
#endif
}
GM_HD double gm_pow2(int k) { return gm_from_bits((uint64_t)(k + 1023) << 52); }  /* -1022<=k<=1023 */

// exp(x) = 2^k 2^(j/64) e^r with x = (64 k + j) ln2/64 + r, |r| <= ln2/128: table value T_j times a
// 5th-order Taylor polynomial, p = T + T (r + r^2 (1/2 + r (1/6 + r (1/24 + r/120)))) (truncation 3.5e-17,
// total error < 1.3 ulp). Results below the smallest normal are flushed to 0 (x < -708.39) so no denormal
// arithmetic is ever involved.
// gm_exp_core: reduced evaluation shared by all variants; returns p in [0.99, 2) and the binary exponent k.
GM_HD double gm_exp_core(double x, const double* tab, long long* k) {
  /* kf = rn(x 64/ln2) by the shift trick: the sum lands in [2^52, 2^53), where doubles are the integers, so the
     low word of z holds the integer (two's complement) and z - SHIFT is kf. No floor, no float->int conversion. */
  const double z = fma(x, GM_INV_L64, GM_RN_SHIFT);
  const double kf = z - GM_RN_SHIFT;
  double r = fma(-kf, GM_L64_HI, x);                    /* exact: L64_HI has 34 significant bits */
  r = fma(-kf, GM_L64_LO, r);
  double p = GM_C(exp, 0);
  p = fma(p, r, GM_C(exp, 1));
  p = fma(p, r, GM_C(exp, 2));
  p = fma(p, r, GM_C(exp, 3));
  const double q = fma(r * r, p, r);
  const int32_t n = (int32_t)(uint32_t)gm_to_bits(z);
  const double t = tab[n & 63];
  *k = (long long)(n >> 6);
  return fma(t, q, t);
}
GM_HD double gm_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893383973096) return gm_inf();
  if (x < -708.3964185322641) return 0.0;
  long long kk;
  const double p = gm_exp_core(x, GM_EXP2TAB, &kk);
  const int k = (int)kk;
  const int k1 = k / 2, k2 = k - k1;
  return (p * gm_pow2(k1)) * gm_pow2(k2);
}

// Same bits as gm_exp(x) for x <= 0 (finite or -inf); no overflow/NaN branches and the 2^k scaling
// is one integer add on the exponent field (the result is normal or flushed to 0, never subnormal).
// Used for exp(lw - max): logsumexp partials and the integer weights of the resampler.
GM_HD double gm_exp_nonpos_t(double x, const double* tab) {
  /* below the flush threshold (down to -inf) the reduced evaluation yields garbage that is never used: no clamp */
  long long k;
  const double p = gm_exp_core(x, tab, &k);
  const double v = gm_scale_pow2(p, (int)k);
  return x < -708.3964185322641 ? 0.0 : v;
}
GM_HD double gm_exp_nonpos(double x) { return gm_exp_nonpos_t(x, GM_EXP2TAB); }

// log(x). x = 2^e * m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1).
// gm_log_core: x positive, finite and normal (no checks); e0 = exponent bias already applied.
GM_HD double gm_log_core(uint64_t b, int e0) {
  int e = e0 + (int)(b >> 52) - 1023;
  double m = gm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > 1.41421356237309514547) { m = m * 0.5; e += 1; }
  const double f = m - 1.0;
  const double s = f / (m + 1.0);
  const double z = s * s;
  double p = GM_C(log, 0);                   /* 1/21, 1/19, ..., 1/3 */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 1; i < 10; ++i) p = fma(p, z, GM_C(log, i));
  const double s2 = s + s;
  const double lm = fma(s2 * z, p, s2);      /* 2s + 2s^3 * P(z) */
  const double ef = (double)e;
  return fma(ef, GM_LN2_HI, fma(ef, GM_LN2_LO, lm));
}
GM_HD double gm_log(double x) {
  if (x != x) return x;
  if (x < 0.0) return gm_nan();
  if (x == 0.0) return -gm_inf();
  uint64_t b = gm_to_bits(x);
  if (b == GM_INF_BITS) return x;
  int e = 0;
  if ((b >> 52) == 0) {                      /* subnormal: scale up by 2^54 (exact) */
    x = x * 18014398509481984.0;
    b = gm_to_bits(x);
    e = -54;
  }
  return gm_log_core(b, e);
}
// same bits as gm_log(x) for positive, finite, normal x (e.g. a uniform in (0,1)); no special cases
GM_HD double gm_log_pos(double x) { return gm_log_core(gm_to_bits(x), 0); }

// x / c for a launch-invariant c with rc = 1.0 / c precomputed: Markstein's sequence
//   q0 = RN(x rc); r = x - q0 c (exact, FMA); q = RN(q0 + r rc)
// gives the correctly rounded quotient, i.e. the same bits as the IEEE division the reference
// formula performs (tests/test_math.py checks 10^7 adversarial operand pairs). Outside the safe
// magnitude range (zero, subnormal results, inf, nan) or when rc == 0 it falls back to x / c.
#define GM_DIV_SPAN (0x7c200000u - 0x03c00000u)
/* span = GM_DIV_SPAN when rc is usable, 0 (= always take the true division) otherwise: gm_div_span(rc) */
GM_HD double gm_div_inv_s(double x, double c, double rc, uint32_t span) {
  /* safe range 2^-963 <= |x| < 2^963, tested on the exponent field with one integer compare
     (0, subnormal, inf and nan fall outside) */
  const uint32_t hx = (uint32_t)(gm_to_bits(x) >> 32) & 0x7fffffffu;
  if (hx - 0x03c00000u >= span) return x / c;
  const double q0 = x * rc;
  const double r = fma(-q0, c, x);
  return fma(r, rc, q0);
}
GM_HD uint32_t gm_div_span(double rc) { return rc == 0.0 ? 0u : GM_DIV_SPAN; }
GM_HD double gm_div_inv(double x, double c, double rc) {
  const uint32_t hx = (uint32_t)(gm_to_bits(x) >> 32) & 0x7fffffffu;
  if (hx - 0x03c00000u >= GM_DIV_SPAN || rc == 0.0) return x / c;
  const double q0 = x * rc;
  const double r = fma(-q0, c, x);
  return fma(r, rc, q0);
}
GM_HD double gm_safe_recip(double c) {       /* rc for gm_div_inv; 0 = "use true division" */
  const double ac = fabs(c);
  return (ac > 1e-150 && ac < 1e150) ? 1.0 / c : 0.0;
}

// sin(pi*t), cos(pi*t) for finite t: t = j/64 + r (exact, |r| <= 1/128), table (S, C) = (sin, cos)(pi j/64)
// with j taken mod 128, short Taylor kernels in x = pi r (|x| <= 0.0246: truncation < 4e-18) and the angle
// addition written around the table value so that its rounding is the dominant error (<= 1 ulp of 1):
//   sin = S + (S (cos x - 1) + C sin x),  cos = C + (C (cos x - 1) - S sin x).
// No quadrant logic; exact at the multiples of 1/2 (the table holds exact 0 and +-1 there).
GM_HD void gm_sincospi_t(double t, const double* tab, double* sn, double* cs) {
  const double zz = fma(t, 64.0, GM_RN_SHIFT);   /* nf = rn(64 t), |t| < 2^24 */
  const double nf = zz - GM_RN_SHIFT;
  const double r = fma(nf, -0.015625, t);    /* exact */
  const double x = r * GM_PI;
  const double z = x * x;
  double ps = GM_C(sin, 0);
  ps = fma(ps, z, GM_C(sin, 1));
  ps = fma(ps, z, GM_C(sin, 2));
  const double sx = fma(x * z, ps, x);       /* sin x */
  double pc = GM_C(cos, 0);
  pc = fma(pc, z, GM_C(cos, 1));
  pc = fma(pc, z, GM_C(cos, 2));
  const double cm = z * pc;                  /* cos x - 1 */
  const int j = (int)((uint32_t)gm_to_bits(zz) & 127u);
  const double S = tab[2 * j], C = tab[2 * j + 1];
  *sn = S + fma(S, cm, C * sx);
  *cs = C + fma(C, cm, -(S * sx));
}
GM_HD void gm_sincospi(double t, double* sn, double* cs) { gm_sincospi_t(t, GM_SINCOSTAB, sn, cs); }

// atan(x) for any finite x and atan2(y, x). Reduction: |x|>1 -> pi/2 - atan(1/|x|);
// then t in [0,1] is shifted by the nearest of atan(k/8), k=0..8, via
// atan(t) = atan(c) + atan((t-c)/(1+t*c)), leaving |u| <= 1/16 for an odd Taylor series.
GM_HD double gm_atan_tab(int k) {
  switch (k) {
    case 0: return 0.0;
    case 1: return 1.2435499454676144e-01;
    case 2: return 2.4497866312686414e-01;
    case 3: return 3.5877067027057225e-01;
    case 4: return 4.6364760900080609e-01;
    case 5: return 5.5859931534356244e-01;
    case 6: return 6.4350110879328437e-01;
    case 7: return 7.1882999962162453e-01;
    default: return 7.8539816339744828e-01;
  }
}
GM_HD double gm_atan(double x) {
  if (x != x) return x;
  const double ax = fabs(x);
  const int inv = ax > 1.0;
  const double t = inv ? 1.0 / ax : ax;
  const double kf = floor(t * 8.0 + 0.5);
  const double c = kf * 0.125;
  const double u = (t - c) / fma(t, c, 1.0);
  const double z = u * u;
  double p = GM_C(atan, 0);                  /* -1/15, 1/13, ..., -1/3 */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 1; i < 7; ++i) p = fma(p, z, GM_C(atan, i));
  double a = gm_atan_tab((int)kf) + fma(u * z, p, u);
  if (inv) a = 1.57079632679489655800 - a;
  return x < 0.0 ? -a : a;
}
GM_HD double gm_atan2(double y, double x) {
  if (x != x || y != y) return gm_nan();
  if (x == 0.0) {
    if (y == 0.0) return 0.0;
    return y > 0.0 ? 1.57079632679489655800 : -1.57079632679489655800;
  }
  const double a = gm_atan(y / x);
  if (x > 0.0) return a;
  return y >= 0.0 ? a + GM_PI : a - GM_PI;
}


// ------------------------------------------------------------------------------------------------
// gm_log_unit: table-driven log for x in (0, 1) with full double precision in absolute terms
// (|error| < 3e-16 * max(1, |log x|)); the result is clamped to <= 0. x = 2^e m; the top 6 mantissa bits
// pick a centre c_i = 1 + (i + 1/2)/64 from a 64-entry table {fl(1/c_i), -log(fl(1/c_i))};
// r = m fl(1/c_i) - 1 with |r| <= 2^-7 and log1p(r) by a 7-term series (truncation 2e-18). No division and
// no compare/select on the mantissa: about half the instructions of gm_log. Used for the radius of the
// Box-Muller transform, sqrt(-2 log u1): close to u1 = 1 the RELATIVE error of the log grows (the absolute
// error does not), which moves a normal that is itself < 1e-7 by < 1e-8.
// ------------------------------------------------------------------------------------------------
#define GM_LOGTAB64_VALUES { \
  0.99224806201550386, 0.0077821404420549628, 0.97709923664122134, 0.023167059281534418, \
  0.96240601503759393, 0.038318864302136657, 0.94814814814814818, 0.053244514518812243, \
  0.93430656934306566, 0.067950661908507778, 0.92086330935251803, 0.082443669211074544, \
  0.90780141843971629, 0.096729626458551141, 0.8951048951048951, 0.11081436634029011, \
  0.88275862068965516, 0.12470347850095725, 0.87074829931972786, 0.13840232285911919, \
  0.85906040268456374, 0.151916042025842, 0.84768211920529801, 0.16524957289530717, \
  0.83660130718954251, 0.17840765747281825, 0.82580645161290323, 0.19139485299962947, \
  0.8152866242038217, 0.20421554142869083, 0.80503144654088055, 0.2168739383006143, \
  0.79503105590062106, 0.2293741010648459, 0.78527607361963192, 0.24171993688714513, \
  0.77575757575757576, 0.25391520998096345, 0.76646706586826352, 0.26596354849713788, \
  0.75739644970414199, 0.27786845100345631, 0.74853801169590639, 0.28963329258304271, \
  0.73988439306358378, 0.30126133057816185, 0.73142857142857143, 0.3127557100038969, \
  0.7231638418079096, 0.32411946865421198, 0.71508379888268159, 0.33535554192113781, \
  0.70718232044198892, 0.34646676734620863, 0.69945355191256831, 0.3574558889218038, \
  0.69189189189189193, 0.36832556115870757, 0.68449197860962563, 0.3790783529349695, \
  0.67724867724867721, 0.38971675114002524, 0.67015706806282727, 0.40024316412701266, \
  0.66321243523316065, 0.41065992498526832, 0.65641025641025641, 0.42096929464412963, \
  0.64974619289340096, 0.43117346481837143, 0.64321608040201006, 0.4412745608048752, \
  0.63681592039800994, 0.45127464413945861, 0.63054187192118227, 0.46117571512217015, \
  0.62439024390243902, 0.47097971521879101, 0.61835748792270528, 0.48068852934575196, \
  0.61244019138755978, 0.49030398804519387, 0.60663507109004744, 0.49982786955644926, \
  0.60093896713615025, 0.5092619017898079, 0.59534883720930232, 0.51860776420804566, \
  0.58986175115207373, 0.52786708962084239, 0.58447488584474883, 0.53704146589688373, \
  0.579185520361991, 0.54613243759813557, 0.57399103139013452, 0.55514150754050162, \
  0.56888888888888889, 0.56407013828480301, 0.56387665198237891, 0.57291975356178537, \
  0.55895196506550215, 0.58169173963462251, 0.55411255411255411, 0.59038744660217635, \
  0.54935622317596566, 0.59900818964608338, 0.5446808510638298, 0.60755525022454182, \
  0.54008438818565396, 0.61602987721551405, 0.53556485355648531, 0.62443328801189357, \
  0.53112033195020747, 0.63276666957103778, 0.52674897119341568, 0.64103117942093124, \
  0.52244897959183678, 0.64922794662510974, 0.51821862348178138, 0.65735807270836, \
  0.51405622489959835, 0.66542263254509049, 0.50996015936254979, 0.6734226752121667, \
  0.50592885375494068, 0.68135922480790312, 0.50196078431372548, 0.689233281238809 }
static const double gm_logtab64_h[128] = GM_LOGTAB64_VALUES;
#if defined(__CUDACC__)
static __constant__ __align__(16) double gm_logtab64_d[128] = GM_LOGTAB64_VALUES;
static __device__ __align__(16) const double gm_logtab64_g[128] = GM_LOGTAB64_VALUES;
#endif
// tab: the 128-entry table (shared-memory copy on the device: lanes index it divergently)
GM_HD double gm_log_unit(double x, const double* tab) {
  const uint64_t b = gm_to_bits(x);
  const int e = (int)(b >> 52) - 1023;
  const int idx = (int)((b >> 46) & 63);
  const double m = gm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  const double r = fma(m, tab[2 * idx], -1.0);
  double p = 1.4285714285714285e-01;                 /* 1/7 */
  p = fma(p, r, -1.6666666666666666e-01);
  p = fma(p, r, 2.0000000000000001e-01);
  p = fma(p, r, -0.25);
  p = fma(p, r, 3.3333333333333331e-01);
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  const double ef = (double)e;
  const double l = fma(ef, GM_LN2_HI, fma(ef, GM_LN2_LO, fma(p, r, tab[2 * idx + 1])));
  return l > 0.0 ? 0.0 : l;
}

// ------------------------------------------------------------------------------------------------
// fp32 pieces of the Box-Muller transform behind the per-particle normal draws (gsmc_rng.cuh, normal_quad): the radius
// and angle come from 32-bit Philox words, so the transform is evaluated in IEEE fp32 (+ - * fma sqrt and bit casts
// only: the same bits on the host and on the device) and the two normals are widened to fp64 afterwards. All MODEL
// arithmetic (transition, logpdf, logsumexp) stays in fp64; only the resolution of the random numbers is 32 bits,
// as in cuRAND's Philox normal generator.
//   gm_nlog_u32f(w)   = -ln((w + 1/2) 2^-32), w a 32-bit word: v = 2w + 1 = 2^n m (m in [1, 2), kept to 32 bits), so
//                       -ln x = lz ln2 + [ln2 - ln m] with lz = 32 - n. The top 6 mantissa bits pick a centre c_i from
//                       a 64-entry table {ic = fl(1/c_i), fl(ln(2 ic))}; r = m ic - 1 (|r| <= 2^-7), log1p(r) by three
//                       terms (truncation < 1e-9), -ln x = fma(lz, ln2, ln(2 ic)) - log1p(r). Next to x = 1 the 24-bit
//                       mantissa would quantise the result at 6e-8 absolute, so x > 1 - 2^-7 (0.8% of the words) goes
//                       through the complement, -log1p(-(1 - x)). Relative error < 1e-5 everywhere (2e-7 typical).
//   gm_sincos_u32f(a) = (sin, cos)(2 pi a 2^-32): a = 2^25 j + d, |d| <= 2^24 (exactly a float), table (S, C) =
//                       (sin, cos)(pi j/64), x = d pi 2^-31 (|x| <= pi/128), sin x and cos x - 1 by two terms each and the
//                       angle addition written around the table values (as gm_sincospi_t). Absolute error < 1e-7.
// ------------------------------------------------------------------------------------------------
static const float gm_logtabf_h[128] = GM_LOGTABF_VALUES;
static const float gm_sincostabf_h[256] = GM_SINCOSTABF_VALUES;
#if defined(__CUDACC__)
static __device__ __align__(16) const float gm_logtabf_g[128] = GM_LOGTABF_VALUES;
static __device__ __align__(16) const float gm_sincostabf_g[256] = GM_SINCOSTABF_VALUES;
#endif
#if defined(__CUDA_ARCH__)
#define GM_LOGTABF gm_logtabf_g
#define GM_SINCOSTABF gm_sincostabf_g
#else
#define GM_LOGTABF gm_logtabf_h
#define GM_SINCOSTABF gm_sincostabf_h
#endif
GM_HD float gm_f32_from_bits(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}
/* the two adjacent float table entries tab[2i], tab[2i+1] with one 8-byte load on the device */
GM_HD void gm_tab_pairf(const float* tab, uint32_t i, float* a, float* b) {
#if defined(__CUDA_ARCH__)
  const float2 v = *reinterpret_cast<const float2*>(tab + 2 * i);
  *a = v.x; *b = v.y;
#else
  *a = tab[2 * i]; *b = tab[2 * i + 1];
#endif
}
GM_HD float gm_nlog_u32f(uint32_t w, const float* tab) {
  if (w >= 0xfe000000u) {
    /* x > 1 - 2^-7: the 24-bit mantissa of x would quantise -ln x at 6e-8 ABSOLUTE; the complement y = 1 - x =
       (~w + 1/2) 2^-32 <= 2^-7 is (nearly) exact instead, -ln x = -log1p(-y) = y (1 + y/2 + y^2/3 + y^3/4) */
    const float y = ((float)(~w) + 0.5f) * 2.32830644e-10f;
    float q = fmaf(0.25f, y, 0.333333343f);
    q = fmaf(q, y, 0.5f);
    q = fmaf(q, y, 1.0f);
    return q * y;
  }
  /* t = the top 32 bits (leading one included) of the 33-bit odd integer 2w + 1: m = t 2^-31 in [1, 2) */
#if defined(__CUDA_ARCH__)
  const int lz = __clz((int)w);                                   /* 32 for w = 0 */
  const uint32_t t = __funnelshift_lc(0x80000000u, w, (uint32_t)lz);
#else
  const int lz = w ? __builtin_clz(w) : 32;
  const uint32_t t = (uint32_t)(((2 * (uint64_t)w + 1) << lz) >> 1);
#endif
  const float m = gm_f32_from_bits(0x3f800000u | ((t >> 8) & 0x007fffffu));
  float ic, t2;
  gm_tab_pairf(tab, (t >> 25) & 63u, &ic, &t2);
  const float r = fmaf(m, ic, -1.0f);
  float q = fmaf(0.333333343f, r, -0.5f);
  q = fmaf(q, r, 1.0f);
  const float base = fmaf((float)lz, GM_LN2F, t2);
  const float l = fmaf(-q, r, base);
  return l < 0.0f ? 0.0f : l;
}
GM_HD void gm_sincos_u32f(uint32_t a, const float* tab, float* sn, float* cs) {
  const uint32_t b = a + 0x01000000u;                              /* wraps with the full turn */
  const int32_t d = (int32_t)(b & 0x01ffffffu) - 0x01000000;
  const float x = (float)d * GM_PI_64_2M25F;
  const float z = x * x;
  const float sx = fmaf(x * z, -0.166666672f, x);                  /* sin x */
  const float cm = z * fmaf(z, 0.0416666679f, -0.5f);              /* cos x - 1 */
  float S, C;
  gm_tab_pairf(tab, b >> 25, &S, &C);
  *sn = S + fmaf(S, cm, C * sx);
  *cs = C + fmaf(C, cm, -(S * sx));
}
/* two standard normals (cos branch, sin branch) from a radius word and an angle word */
GM_HD void gm_box_muller_u32(uint32_t wr, uint32_t wa, const float* ltab, const float* sctab, double* z0, double* z1) {
  const float l = gm_nlog_u32f(wr, ltab);
  const float r = sqrtf(l + l);
  float sn, cs;
  gm_sincos_u32f(wa, sctab, &sn, &cs);
  *z0 = (double)(r * cs);
  *z1 = (double)(r * sn);
}

#if defined(__cplusplus)
// ------------------------------------------------------------------------------------------------
// Batch forms: K independent evaluations with the coefficient loop outside and the element loop
// inside. Element by element the arithmetic is identical to the scalar functions above (same bits);
// the batch form lets one constant load feed K FMAs and gives the scheduler K independent Horner
// chains to interleave (the scalar chains stall on their own FMA latency).
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define GM_UNROLL _Pragma("unroll")
#else
#define GM_UNROLL
#endif

template <int K> GM_HD void gm_exp_nonpos_v(const double* x, double* out, const double* tab) {
  double z[K], r[K], p[K];
  GM_UNROLL for (int k = 0; k < K; ++k) {
    z[k] = fma(x[k], GM_INV_L64, GM_RN_SHIFT);
    const double kf = z[k] - GM_RN_SHIFT;
    r[k] = fma(-kf, GM_L64_HI, x[k]);
    r[k] = fma(-kf, GM_L64_LO, r[k]);
    p[k] = GM_C(exp, 0);
  }
  GM_UNROLL for (int i = 1; i < 4; ++i) {
    const double c = GM_C(exp, i);
    GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], c);
  }
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const double q = fma(r[k] * r[k], p[k], r[k]);
    const int32_t n = (int32_t)(uint32_t)gm_to_bits(z[k]);
    const double t = tab[n & 63];
    const double v = gm_scale_pow2(fma(t, q, t), n >> 6);
    out[k] = x[k] < -708.3964185322641 ? 0.0 : v;
  }
}

// gm_log_pos for K positive, finite, normal arguments
template <int K> GM_HD void gm_log_pos_v(const double* x, double* out) {
  double s[K], z[K], p[K], ef[K];
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const uint64_t b = gm_to_bits(x[k]);
    int e = (int)(b >> 52) - 1023;
    double m = gm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.41421356237309514547) { m = m * 0.5; e += 1; }
    const double f = m - 1.0;
    s[k] = f / (m + 1.0);
    z[k] = s[k] * s[k];
    ef[k] = (double)e;
    p[k] = GM_C(log, 0);
  }
  GM_UNROLL for (int i = 1; i < 10; ++i) {
    const double c = GM_C(log, i);
    GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], z[k], c);
  }
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const double s2 = s[k] + s[k];
    const double lm = fma(s2 * z[k], p[k], s2);
    out[k] = fma(ef[k], GM_LN2_HI, fma(ef[k], GM_LN2_LO, lm));
  }
}

template <int K> GM_HD void gm_log_unit_v(const double* x, const double* tab, double* out) {
  double r[K], p[K], c[K], ef[K];
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const uint64_t b = gm_to_bits(x[k]);
    const int idx = (int)((b >> 46) & 63);
    const double m = gm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    ef[k] = (double)((int)(b >> 52) - 1023);
    double ic;
    gm_tab_pair(tab, idx, &ic, &c[k]);
    r[k] = fma(m, ic, -1.0);
    p[k] = 1.4285714285714285e-01;
  }
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], -1.6666666666666666e-01);
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], 2.0000000000000001e-01);
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], -0.25);
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], 3.3333333333333331e-01);
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], -0.5);
  GM_UNROLL for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], 1.0);
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const double l = fma(ef[k], GM_LN2_HI, fma(ef[k], GM_LN2_LO, fma(p[k], r[k], c[k])));
    out[k] = l > 0.0 ? 0.0 : l;
  }
}

template <int K> GM_HD void gm_sincospi_v(const double* t, double* sn, double* cs, const double* tab) {
  double zz[K], x[K], z[K], ps[K], pc[K];
  GM_UNROLL for (int k = 0; k < K; ++k) {
    zz[k] = fma(t[k], 64.0, GM_RN_SHIFT);
    const double r = fma(zz[k] - GM_RN_SHIFT, -0.015625, t[k]);
    x[k] = r * GM_PI;
    z[k] = x[k] * x[k];
    ps[k] = GM_C(sin, 0);
    pc[k] = GM_C(cos, 0);
  }
  GM_UNROLL for (int i = 1; i < 3; ++i) {
    const double a = GM_C(sin, i), b = GM_C(cos, i);
    GM_UNROLL for (int k = 0; k < K; ++k) { ps[k] = fma(ps[k], z[k], a); pc[k] = fma(pc[k], z[k], b); }
  }
  GM_UNROLL for (int k = 0; k < K; ++k) {
    const double sx = fma(x[k] * z[k], ps[k], x[k]);
    const double cm = z[k] * pc[k];
    const int j = (int)((uint32_t)gm_to_bits(zz[k]) & 127u);
    double S, C;
    gm_tab_pair(tab, j, &S, &C);
    sn[k] = S + fma(S, cm, C * sx);
    cs[k] = C + fma(C, cm, -(S * sx));
  }
}
#endif  /* __cplusplus */

#endif  /* GSMC_MATH_H */
