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
//   normals  (stream 0): element e -> call e>>1; Box-Muller u1=((a>>11)+.5)2^-53, u2=(b>>11)2^-53,
//                        r=sqrt(-2 gm_log_unit(u1)); e even -> r cos(2 pi u2), e odd -> r sin(2 pi u2)
//   uniforms (stream 1,3): element e -> call e>>1; (e odd ? b : a)>>11 * 2^-53   in [0,1)
//   spacings (stream 2): element e -> call e>>2, 32-bit word e&3 (out0..out3); u = (w+.5)2^-32;
//                        floor(-gm_log_tab(u) * 2^27)   (fixed-point Exp(1) variate; < 2^32 because
//                        -log u <= 33 ln 2, so a spacing is stored in 4 bytes)
// Particle i's j-th normal at a step that needs nz normals per particle is element i*nz + j.
#ifndef GSMC_RNG_CUH
#define GSMC_RNG_CUH

#include <stdint.h>
#include "gsmc_math.h"

enum { GSMC_STREAM_NORMAL = 0, GSMC_STREAM_UNIFORM = 1, GSMC_STREAM_RESAMPLE = 2, GSMC_STREAM_SAMPLE = 3 };

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

// two standard normals from one call
__host__ __device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t call, uint32_t t, const double* ltab, double* z0, double* z1) {
  const PhiloxOut o = philox_call(seed, call, t, GSMC_STREAM_NORMAL);
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

__host__ __device__ __forceinline__ uint32_t spacing_from_word(uint32_t w, const double* tab) {
  const double u = gm_u32_to_unit(w);                           // (w + 0.5) 2^-32
  return (uint32_t)(-gm_log_tab(u, tab) * GM_SPACING_SCALE);     // = floor: the product lies in (0, 2^32)
}
// the four spacings 4c .. 4c+3 of call c
template <class Key>
__host__ __device__ __forceinline__ void spacing_quad(const Key& seed, uint64_t call, uint32_t rho, const double* tab, uint32_t* e) {
  const PhiloxOut o = philox_call(seed, call, rho, GSMC_STREAM_RESAMPLE);
  const uint32_t w[4] = {(uint32_t)o.a, (uint32_t)(o.a >> 32), (uint32_t)o.b, (uint32_t)(o.b >> 32)};
  double u[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) u[j] = gm_u32_to_unit(w[j]);        // (w + 0.5) 2^-32
  gm_log_tab_v<4>(u, tab, l);
#pragma unroll
  for (int j = 0; j < 4; ++j) e[j] = (uint32_t)(-l[j] * GM_SPACING_SCALE);    // = floor: the product lies in (0, 2^32)
}
// spacing of a single element (uniform across the calling warp / a single thread)
__host__ __device__ __forceinline__ uint64_t spacing_one(uint64_t seed, uint64_t element, uint32_t rho, const double* tab) {
  const PhiloxOut o = philox_call(seed, element >> 2, rho, GSMC_STREAM_RESAMPLE);
  const uint32_t w[4] = {(uint32_t)o.a, (uint32_t)(o.a >> 32), (uint32_t)o.b, (uint32_t)(o.b >> 32)};
  return spacing_from_word(w[element & 3], tab);
}

// Batch forms (same bits as the scalar functions, see gsmc_math.h "Batch forms").
// K Philox calls -> 2K standard normals z[2m] (cos branch), z[2m+1] (sin branch)
template <int K, class Key>
__host__ __device__ __forceinline__ void normal_pairs_v(const Key& seed, const uint64_t* calls, uint32_t t, const double* ltab, const double* sctab, double* z) {
  double u1[K], t2[K], l[K], sn[K], cs[K];
#pragma unroll
  for (int m = 0; m < K; ++m) {
    const PhiloxOut o = philox_call(seed, calls[m], t, GSMC_STREAM_NORMAL);
    u1[m] = ((double)(o.a >> 11) + 0.5) * 0x1p-53;
    const double u2 = (double)(o.b >> 11) * 0x1p-53;
    t2[m] = 2.0 * u2;
  }
  gm_log_unit_v<K>(u1, ltab, l);
  gm_sincospi_v<K>(t2, sn, cs, sctab);
#pragma unroll
  for (int m = 0; m < K; ++m) {
    const double r = sqrt(-2.0 * l[m]);
    z[2 * m] = r * cs[m];
    z[2 * m + 1] = r * sn[m];
  }
}
#endif
