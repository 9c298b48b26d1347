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

#if defined(__CUDACC__)
#define GM_HD __host__ __device__ __forceinline__
#else
#define GM_HD static inline
#endif

#define GM_PI 3.141592653589793115997963468544185161590576171875      /* Float64(pi) */
#define GM_TWO_PI 6.28318530717958623199592693708837032318115234375   /* 2.0*pi     */
#define GM_INF_BITS 0x7ff0000000000000ULL

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
GM_HD double gm_inf(void) { return gm_from_bits(GM_INF_BITS); }
GM_HD double gm_nan(void) { return gm_from_bits(0x7ff8000000000000ULL); }
GM_HD double gm_pow2(int k) { return gm_from_bits((uint64_t)(k + 1023) << 52); }  /* -1022<=k<=1023 */

// exp(x). Results below the smallest normal are flushed to 0 (x < -708.39) so no
// denormal arithmetic is ever involved. 13th-order Taylor on |r| <= ln2/2.
GM_HD double gm_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893383973096) return gm_inf();
  if (x < -708.3964185322641) return 0.0;
  const double kf = floor(x * 1.44269504088896338700e+00 + 0.5);
  double r = fma(-kf, 6.93147180369123816490e-01, x);   /* ln2 hi (fdlibm split) */
  r = fma(-kf, 1.90821492927058770002e-10, r);          /* ln2 lo */
  double p = 1.6059043836821613e-10;                    /* 1/13! */
  p = fma(p, r, 2.08767569878681e-09);                  /* 1/12! */
  p = fma(p, r, 2.505210838544172e-08);                 /* 1/11! */
  p = fma(p, r, 2.755731922398589e-07);                 /* 1/10! */
  p = fma(p, r, 2.7557319223985893e-06);                /* 1/9!  */
  p = fma(p, r, 2.48015873015873e-05);                  /* 1/8!  */
  p = fma(p, r, 1.984126984126984e-04);                 /* 1/7!  */
  p = fma(p, r, 1.388888888888889e-03);                 /* 1/6!  */
  p = fma(p, r, 8.333333333333333e-03);                 /* 1/5!  */
  p = fma(p, r, 4.1666666666666664e-02);                /* 1/4!  */
  p = fma(p, r, 1.6666666666666666e-01);                /* 1/3!  */
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k = (int)kf;
  const int k1 = k / 2, k2 = k - k1;
  return (p * gm_pow2(k1)) * gm_pow2(k2);
}

// log(x). x = 2^e * m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1).
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
  e += (int)(b >> 52) - 1023;
  double m = gm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > 1.41421356237309514547) { m = m * 0.5; e += 1; }
  const double f = m - 1.0;
  const double s = f / (m + 1.0);
  const double z = s * s;
  double p = 4.7619047619047616e-02;         /* 1/21 */
  p = fma(p, z, 5.2631578947368418e-02);     /* 1/19 */
  p = fma(p, z, 5.8823529411764705e-02);     /* 1/17 */
  p = fma(p, z, 6.6666666666666666e-02);     /* 1/15 */
  p = fma(p, z, 7.6923076923076927e-02);     /* 1/13 */
  p = fma(p, z, 9.0909090909090912e-02);     /* 1/11 */
  p = fma(p, z, 1.1111111111111110e-01);     /* 1/9  */
  p = fma(p, z, 1.4285714285714285e-01);     /* 1/7  */
  p = fma(p, z, 2.0000000000000001e-01);     /* 1/5  */
  p = fma(p, z, 3.3333333333333331e-01);     /* 1/3  */
  const double s2 = s + s;
  const double lm = fma(s2 * z, p, s2);      /* 2s + 2s^3 * P(z) */
  const double ef = (double)e;
  return fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, lm));
}

// sin(pi*t), cos(pi*t) for finite t. Range reduction is exact; kernels are Taylor
// polynomials on |pi*r| <= pi/4.
GM_HD void gm_sincospi(double t, double* sn, double* cs) {
  const double nf = floor(t + t + 0.5);
  const double r = fma(nf, -0.5, t);         /* exact: |r| <= 1/4 */
  const double x = r * GM_PI;
  const double z = x * x;
  double ps = -8.2206352466243295e-18;       /* -1/19! */
  ps = fma(ps, z, 2.8114572543455206e-15);   /*  1/17! */
  ps = fma(ps, z, -7.6471637318198164e-13);  /* -1/15! */
  ps = fma(ps, z, 1.6059043836821613e-10);   /*  1/13! */
  ps = fma(ps, z, -2.5052108385441720e-08);  /* -1/11! */
  ps = fma(ps, z, 2.7557319223985893e-06);   /*  1/9!  */
  ps = fma(ps, z, -1.9841269841269841e-04);  /* -1/7!  */
  ps = fma(ps, z, 8.3333333333333332e-03);   /*  1/5!  */
  ps = fma(ps, z, -1.6666666666666666e-01);  /* -1/3!  */
  const double s0 = fma(x * z, ps, x);
  double pc = -1.5619206968586225e-16;       /* -1/18! */
  pc = fma(pc, z, 4.7794773323873853e-14);   /*  1/16! */
  pc = fma(pc, z, -1.1470745597729725e-11);  /* -1/14! */
  pc = fma(pc, z, 2.0876756987868100e-09);   /*  1/12! */
  pc = fma(pc, z, -2.7557319223985888e-07);  /* -1/10! */
  pc = fma(pc, z, 2.4801587301587302e-05);   /*  1/8!  */
  pc = fma(pc, z, -1.3888888888888889e-03);  /* -1/6!  */
  pc = fma(pc, z, 4.1666666666666664e-02);   /*  1/4!  */
  pc = fma(pc, z, -0.5);
  const double c0 = fma(z, pc, 1.0);
  const long long n = (long long)nf;
  const int q = (int)(n & 3);
  double so, co;
  if (q == 0) { so = s0; co = c0; }
  else if (q == 1) { so = c0; co = -s0; }
  else if (q == 2) { so = -s0; co = -c0; }
  else { so = -c0; co = s0; }
  *sn = so; *cs = co;
}

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
  double p = -6.6666666666666666e-02;        /* -1/15 */
  p = fma(p, z, 7.6923076923076927e-02);     /*  1/13 */
  p = fma(p, z, -9.0909090909090912e-02);    /* -1/11 */
  p = fma(p, z, 1.1111111111111110e-01);     /*  1/9  */
  p = fma(p, z, -1.4285714285714285e-01);    /* -1/7  */
  p = fma(p, z, 2.0000000000000001e-01);     /*  1/5  */
  p = fma(p, z, -3.3333333333333331e-01);    /* -1/3  */
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

#endif  /* GSMC_MATH_H */
