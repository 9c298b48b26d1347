// gsmc_rng.cuh -- counter-based Philox4x32-10 and the draw layout of the filter.
//
// Replaces the reference's global MersenneTwister (`randn()` in
// src/modeling_library/distributions/normal.jl:96, `rand()` in bernoulli.jl:19 and
// Distributions.jl's categorical sampler): every draw is a pure function of
// (seed, global element index, time step, stream), so the result does not depend on
// the thread/block/GPU that computes it.
//
// Layout (shared with the oracle's independent restatement, oracle/gsmc_oracle.c):
//   call c of (seed, t, stream)    = Philox(ctr = {lo32(c), hi32(c), t, stream}, key = {lo32(seed), hi32(seed)})
//   word a = out0 | out1<<32, word b = out2 | out3<<32
//   normals  (stream 0): element e -> call e>>2, pair (e>>1)&1 of the call: radius word out[2 pair], angle word
//                        out[2 pair + 1]; Box-Muller in fp32 on the 32-bit words (gsmc_math.h, gm_box_muller_u32):
//                        r = sqrtf(-2 ln((wr+.5)2^-32)); e even -> r cos(2 pi wa 2^-32), e odd -> r sin(...); widened to
//                        fp64. One call feeds FOUR consecutive elements (a thread owns quads of neighbouring particles).
//   uniforms (stream 1,3): element e -> call e>>1; (e odd ? b : a)>>11 * 2^-53   in [0,1)
//   resampling draws (stream 2): output slot k -> call k>>2, 32-bit word k&3 (out0..out3); u = (w+.5)2^-32 places
//                        the draw inside its group's interval (see "grouped order statistics" below)
//   observation choices of unobserved steps (stream 5): particle i -> element i (a normal by the 53-bit fp64 Box-Muller
//                        below, normal_pair: element e -> call e>>1; or a uniform for discrete emissions)
//   group gaps (stream 4): Marsaglia-Tsang Gamma variate of group j, attempt a: normal = cos branch of call
//                        (j<<5 | 2a), uniform = ((word a of call (j<<5 | 2a+1)) >> 11 + .5) 2^-53
// Particle i's j-th normal at a step that needs nz normals per particle is element i*nz + j.
#ifndef GSMC_RNG_CUH
#define GSMC_RNG_CUH

#include <stdint.h>
#include "gsmc_math.h"

enum { GSMC_STREAM_NORMAL = 0, GSMC_STREAM_UNIFORM = 1, GSMC_STREAM_RESAMPLE = 2, GSMC_STREAM_SAMPLE = 3, GSMC_STREAM_GAP = 4, GSMC_STREAM_OBS = 5 };

struct PhiloxOut { uint64_t a, b; };

__host__ __device__ __forceinline__ PhiloxOut philox_call(uint64_t seed, uint64_t call, uint32_t t, uint32_t stream) {
  uint32_t c0 = (uint32_t)call, c1 = (uint32_t)(call >> 32), c2 = t, c3 = stream;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;     // one IMAD.WIDE.U32 each
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  PhiloxOut o;
  o.a = (uint64_t)c0 | ((uint64_t)c1 << 32);
  o.b = (uint64_t)c2 | ((uint64_t)c3 << 32);
  return o;
}

// The ten round keys of a seed, precomputed on the host and passed as a kernel parameter: the rounds then take
// them as constant-bank operands instead of recomputing k += W per round and call.
struct PhiloxKeys { uint32_t k[20]; };
static inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { K.k[2 * r] = k0; K.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  return K;
}
__device__ __forceinline__ PhiloxOut philox_call(const PhiloxKeys& K, uint64_t call, uint32_t t, uint32_t stream) {
  uint32_t c0 = (uint32_t)call, c1 = (uint32_t)(call >> 32), c2 = t, c3 = stream;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
  }
  PhiloxOut o;
  o.a = (uint64_t)c0 | ((uint64_t)c1 << 32);
  o.b = (uint64_t)c2 | ((uint64_t)c3 << 32);
  return o;
}

// two standard normals from one call, 53-bit uniforms and fp64 Box-Muller: the Gamma gaps of the resampler (stream 4)
// and the observation choices of unobserved steps (stream 5). The per-particle draws of the models use normal_quad.
__host__ __device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t call, uint32_t t, const double* ltab, double* z0, double* z1,
                                                      uint32_t stream = GSMC_STREAM_NORMAL) {
  const PhiloxOut o = philox_call(seed, call, t, stream);
  const double u1 = ((double)(o.a >> 11) + 0.5) * 0x1p-53;
  const double u2 = (double)(o.b >> 11) * 0x1p-53;
  const double r = sqrt(-2.0 * gm_log_unit(u1, ltab));   // u1 in (0,1): log <= 0
  double s, c;
  gm_sincospi(2.0 * u2, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}

__host__ __device__ __forceinline__ void uniform_pair(uint64_t seed, uint64_t call, uint32_t t, uint32_t stream, double* u0, double* u1) {
  const PhiloxOut o = philox_call(seed, call, t, stream);
  *u0 = (double)(o.a >> 11) * 0x1p-53;
  *u1 = (double)(o.b >> 11) * 0x1p-53;
}

// ------------------------------------------------------------------------------------------------
// Grouped order statistics: the M sorted uniforms behind the M iid categorical draws of a resampling event
// (particle_filter.jl:200) are generated group by group, GSMC_GROUP output slots per group.
//   * Group j opens at the order statistic A_j / S_tot with A_0 = head ~ Exp(1), A_{j+1} = A_j + g_j and
//     g_j ~ Gamma(r_j), r_j = number of slots of group j (the sum of r consecutive Exp(1) spacings), S_tot = A_last.
//   * The other r_j - 1 draws of the group are iid uniform between the two order statistics that bracket it
//     (Markov property of order statistics), in the order they are drawn: x_k = fma(u_k, g_j, A_j).
// The multiset of all M values is an exact sample of M iid uniforms; every group maps to the CDF window between the
// ancestors of its bracketing order statistics, so search and gather stream through memory without a per-draw
// logarithm or a prefix sum over the draws. Gaps are integers (fixed point, scale 2^20, < 2^53 in total) so their
// prefix sums are associative: the result does not depend on the block or GPU count.
// ------------------------------------------------------------------------------------------------
#define GSMC_GROUP 256
#define GSMC_GROUP_SHIFT 8
#define GM_GAP_SCALE 1048576.0          /* 2^20 */
#if defined(__CUDA_ARCH__)
#define GM_LOGTAB64 gm_logtab64_g
#else
#define GM_LOGTAB64 gm_logtab64_h
#endif
// floor(Gamma(shape) * 2^20), shape >= 1: Marsaglia & Tsang (2000); after 16 rejections (p < 1e-20) the mean.
__host__ __device__ inline uint64_t gap_variate(uint64_t seed, uint64_t group, uint32_t shape, uint32_t rho) {
  const double d = (double)shape - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  double v = 1.0;
  for (int attempt = 0; attempt < 16; ++attempt) {
    const uint64_t call = (group << 5) | (uint64_t)(2 * attempt);
    const PhiloxOut o = philox_call(seed, call, rho, GSMC_STREAM_GAP);
    const double u1 = ((double)(o.a >> 11) + 0.5) * 0x1p-53;
    const double u2 = (double)(o.b >> 11) * 0x1p-53;
    const double r = sqrt(-2.0 * gm_log_unit(u1, GM_LOGTAB64));
    double sn, cs;
    gm_sincospi(2.0 * u2, &sn, &cs);
    const double z = r * cs;
    const double w = 1.0 + c * z;
    if (!(w > 0.0)) continue;
    const double w3 = (w * w) * w;
    const PhiloxOut p = philox_call(seed, call + 1, rho, GSMC_STREAM_GAP);
    const double u = ((double)(p.a >> 11) + 0.5) * 0x1p-53;
    const double lhs = gm_log(u);
    const double zz = (0.5 * z) * z;
    const double rhs = ((zz + d) - d * w3) + d * gm_log(w3);
    if (lhs < rhs) { v = w3; break; }
  }
  return (uint64_t)((d * v) * GM_GAP_SCALE);
}
// gap of the group that starts at output slot k0 of an event with m draws (0 beyond the last group)
__host__ __device__ inline uint64_t gap_of_group(uint64_t seed, uint64_t k0, uint64_t m, uint32_t rho) {
  if (k0 >= m) return 0;
  const uint64_t left = m - k0;
  return gap_variate(seed, k0 >> GSMC_GROUP_SHIFT, (uint32_t)(left < GSMC_GROUP ? left : GSMC_GROUP), rho);
}
// A_0: the Exp(1) gap below the first order statistic, drawn as the group one past the last
__host__ __device__ inline uint64_t gap_head(uint64_t seed, uint64_t m, uint32_t rho) {
  return gap_variate(seed, (m + GSMC_GROUP - 1) >> GSMC_GROUP_SHIFT, 1, rho);
}
// threshold of a draw at position x (in gap units) against the integer CDF: min(trunc(x * ratio), C_N - 1)
__host__ __device__ __forceinline__ uint64_t threshold_u64(double x, double ratio, uint64_t cn) {
  const uint64_t T = (uint64_t)(x * ratio);
  return T < cn ? T : cn - 1;
}

// four standard normals (elements 4 call .. 4 call + 3 of the step's virtual normal array) from one call
template <class Key>
__host__ __device__ __forceinline__ void normal_quad(const Key& seed, uint64_t call, uint32_t t, const float* ltab, const float* sctab, double* z) {
  const PhiloxOut o = philox_call(seed, call, t, GSMC_STREAM_NORMAL);
  gm_box_muller_u32((uint32_t)o.a, (uint32_t)(o.a >> 32), ltab, sctab, z, z + 1);
  gm_box_muller_u32((uint32_t)o.b, (uint32_t)(o.b >> 32), ltab, sctab, z + 2, z + 3);
}
// K calls -> 4K standard normals, z[4m .. 4m+3] from calls[m] (same bits as normal_quad; the K independent
// evaluations are written one after the other and interleaved by the compiler)
template <int K, class Key>
__device__ __forceinline__ void normal_quads_v(const Key& seed, const uint64_t* calls, uint32_t t, const float* ltab, const float* sctab, double* z) {
#pragma unroll
  for (int m = 0; m < K; ++m) normal_quad(seed, calls[m], t, ltab, sctab, z + 4 * m);
}
#endif
