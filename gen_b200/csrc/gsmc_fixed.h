// gsmc_fixed.h -- exact 64-bit fixed-point helpers of the resampler (host + device).
//
// muldiv_floor(a, b, d) = floor(a*b / d) for a <= d (so the quotient fits in 64 bits), d < 2^63.
// Used for the sorted-uniform thresholds T_k = floor(S_k * C_N / S_tot): computing T_k once lets
// every CDF probe be a plain 64-bit compare instead of a 128-bit multiply-compare. The quotient is
// found with two floating-point refinement steps and an exact integer fix-up, so the result is the
// exact integer floor (tests/test_math.py checks it against unsigned __int128 division).
#ifndef GSMC_FIXED_H
#define GSMC_FIXED_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define GF_HD __host__ __device__ __forceinline__
#else
#define GF_HD static inline
#endif

GF_HD uint64_t gf_mulhi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

typedef struct MulDiv {    // per-launch constants for a fixed multiplier b and divisor d
  uint64_t b, d;
  double ratio;            // (double)b / (double)d
  double inv_d;            // 1 / (double)d
} MulDiv;
GF_HD MulDiv make_muldiv(uint64_t b, uint64_t d) {
  MulDiv m;
  m.b = b; m.d = d;
  m.inv_d = 1.0 / (double)d;
  m.ratio = (double)b / (double)d;
  return m;
}

// floor(a * m.b / m.d), requires a <= m.d, m.d < 2^63, m.b < 2^63
GF_HD uint64_t muldiv_floor(uint64_t a, const MulDiv m) {
  const uint64_t phi = gf_mulhi(a, m.b), plo = a * m.b;          // P = a*b (128 bit)
  // first guess: relative error ~2^-51 -> absolute error up to ~2^12
  uint64_t q = (uint64_t)((double)a * m.ratio);
  // remainder r = P - q*d as a signed 128-bit number (|r| <~ 2^13 * d)
  uint64_t qhi = gf_mulhi(q, m.d), qlo = q * m.d;
  uint64_t rlo = plo - qlo;
  int64_t rhi = (int64_t)(phi - qhi - (plo < qlo ? 1u : 0u));
  // second step: r / d is small (|.| < 2^14), so double arithmetic gets it to within 1
  const double rd = (double)rhi * 18446744073709551616.0 + (double)rlo;
  const int64_t adj = (int64_t)floor(rd * m.inv_d);
  q += (uint64_t)adj;
  // exact remainder now fits comfortably in a signed 64-bit value: r1 = P - q*d in (-2d, 2d)
  int64_t r1 = (int64_t)(plo - q * m.d);
  while (r1 < 0) { q -= 1; r1 += (int64_t)m.d; }
  while (r1 >= (int64_t)m.d) { q += 1; r1 -= (int64_t)m.d; }
  return q;
}

#endif
