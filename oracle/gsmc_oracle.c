/* gsmc_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see gsmc_oracle.h).
 *
 * Sequential fp64 restatement of Gen.jl's particle filter / importance sampling
 * (paths relative to /root/reference). Particles are looped over exactly like
 * `for i=1:num_particles` in src/inference/particle_filter.jl:84,103,143,165,202;
 * an optional OpenMP pragma spreads that loop over host threads for the
 * `bench.py --impl reference` leg (each particle's draws are counter-based, so the
 * result does not depend on the thread count).
 *
 * Transcendentals come from gen_b200/csrc/gsmc_math.h (IEEE-only arithmetic, so the
 * same bits come out on the GPU); compile with -DORC_USE_LIBM to use glibc's
 * exp/log/sincos/atan2 instead (tests/test_oracle.py checks both builds agree to
 * 1e-12 and that both reproduce the reference's golden values).
 */
#include "gsmc_oracle.h"
#include "../gen_b200/csrc/gsmc_math.h"
#include "../gen_b200/csrc/gsmc_fixed.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

static __thread char g_err[512];
const char* orc_last_error(void) { return g_err; }
#define ORC_FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return -1; } while (0)

/* ------------------------------------------------------------------------- */
/* math                                                                       */
/* ------------------------------------------------------------------------- */
#ifdef ORC_USE_LIBM
double orc_exp(double x) { return exp(x); }
double orc_log(double x) { return log(x); }
double orc_atan2(double y, double x) { return atan2(y, x); }
void orc_sincospi(double t, double* s, double* c) { *s = sin(M_PI * t); *c = cos(M_PI * t); }
#else
double orc_exp(double x) { return gm_exp(x); }
double orc_log(double x) { return gm_log(x); }
double orc_atan2(double y, double x) { return gm_atan2(y, x); }
void orc_sincospi(double t, double* s, double* c) { gm_sincospi(t, s, c); }
#endif

/* checks of gsmc_math.h helpers the device uses (tests/test_math.py) */
double orc_div_inv(double x, double c) { return gm_div_inv(x, c, gm_safe_recip(c)); }
double orc_log_pos(double x) { return gm_log_pos(x); }
/* log of a uniform in (0,1) for the Box-Muller radius: table-driven (gsmc_math.h), libm under -DORC_USE_LIBM */
#ifdef ORC_USE_LIBM
double orc_log_unit(double x) { return log(x); }
#else
double orc_log_unit(double x) { return gm_log_unit(x, gm_logtab64_h); }
#endif
double orc_exp_nonpos(double x) { return gm_exp_nonpos(x); }
/* muldiv_floor vs unsigned __int128 division */
int64_t orc_muldiv_mismatches(uint64_t seed, int64_t n) {
  uint64_t s = seed ? seed : 1; int64_t bad = 0;
  for (int64_t it = 0; it < n; ++it) {
    uint64_t r[4];
    for (int j = 0; j < 4; ++j) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; r[j] = s; }
    int sb = 1 + (int)(r[0] % 62), sd = 1 + (int)(r[1] % 62);
    uint64_t b = (r[2] >> (64 - sb)) | 1, d = (r[3] >> (64 - sd)) | 1;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    uint64_t a = s % (d + 1);
    int mode = (int)(it & 7);
    if (mode == 0) a = d;
    if (mode == 1) a = d - 1;
    if (mode == 2) a = 0;
    if (mode == 4) b = ((uint64_t)1 << 62) - (s % 1000);
    if (mode == 5) { d = ((uint64_t)1 << 62) + (s % 100000); a = r[0] % (d + 1); }
    if (muldiv_floor(a, make_muldiv(b, d)) != (uint64_t)(((u128)a * b) / d)) ++bad;
  }
  return bad;
}
/* number of operand pairs (adversarial mantissas included) where the Markstein sequence differs from x / c */
int64_t orc_div_inv_mismatches(uint64_t seed, int64_t n) {
  uint64_t s = seed ? seed : 88172645463325252ULL;
  int64_t bad = 0;
  for (int64_t it = 0; it < n; ++it) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t mc = s;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t mx = s;
    int mode = (int)(it & 7);
    if (mode == 0) mc |= 0x000fffffffff0000ULL;
    if (mode == 1) mc &= 0xfff0000000000fffULL;
    if (mode == 2) mc |= 0x000fffffffffffffULL;
    double c = gm_from_bits((mc & 0x000fffffffffffffULL) | ((uint64_t)(1023 - 40 + (mc >> 57)) << 52));
    double x = gm_from_bits((mx & 0x000fffffffffffffULL) | ((uint64_t)(1023 - 60 + (mx >> 57)) << 52));
    if (mx & (1ULL << 56)) x = -x;
    if (gm_to_bits(orc_div_inv(x, c)) != gm_to_bits(x / c)) ++bad;
  }
  return bad;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011), written out independently of the       */
/* product's device version; pinned by the Random123 known-answer vectors.    */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one call -> two 64-bit words */
static void philox_pair(uint64_t seed, uint64_t call, uint32_t t, uint32_t stream, uint64_t* a, uint64_t* b) {
  uint32_t ctr[4] = { (uint32_t)call, (uint32_t)(call >> 32), t, stream };
  uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  *a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
  *b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

/* fp64 Box-Muller on 53-bit uniforms: the Gamma gaps of the resampler and the observation choices of unobserved steps
 * (call c yields elements 2c (cos branch) and 2c+1 (sin branch) of those streams). */
static void box_muller(uint64_t a, uint64_t b, double* z0, double* z1) {
  const double u1 = ((double)(a >> 11) + 0.5) * 0x1p-53;   /* (0,1) */
  const double u2 = (double)(b >> 11) * 0x1p-53;           /* [0,1) */
  const double r = sqrt(-2.0 * orc_log_unit(u1));
  double s, c;
  orc_sincospi(2.0 * u2, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}

/* Per-particle normal draws (stream 0): call c yields elements 4c .. 4c+3; pair 0 = (radius word out[0], angle word
 * out[1]), pair 1 = (out[2], out[3]); cos branch first. The transform is evaluated in fp32 on the 32-bit words
 * (gm_box_muller_u32 of the shared gsmc_math.h, also under -DORC_USE_LIBM: its accuracy is checked directly against
 * glibc in tests/test_math.py) and widened to fp64. */
void orc_fill_normals(uint64_t seed, uint32_t t, uint64_t first, uint64_t count, double* out) {
  for (uint64_t e = first; e < first + count; ++e) {
    uint32_t ctr[4] = { (uint32_t)(e >> 2), (uint32_t)((e >> 2) >> 32), t, ORC_STREAM_NORMAL };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    double z0, z1;
    orc_philox4x32_10(ctr, key, o);
    const int pair = (int)((e >> 1) & 1);
    gm_box_muller_u32(o[2 * pair], o[2 * pair + 1], gm_logtabf_h, gm_sincostabf_h, &z0, &z1);
    out[e - first] = (e & 1) ? z1 : z0;
  }
}
/* pieces of the fp32 transform for tests/test_math.py */
float orc_nlog_u32f(uint32_t w) { return gm_nlog_u32f(w, gm_logtabf_h); }
void orc_sincos_u32f(uint32_t a, float* s, float* c) { gm_sincos_u32f(a, gm_sincostabf_h, s, c); }

void orc_fill_uniforms(uint64_t seed, uint32_t t, uint32_t stream, uint64_t first, uint64_t count, double* out) {
  for (uint64_t e = first; e < first + count; ++e) {
    uint64_t a, b;
    philox_pair(seed, e >> 1, t, stream, &a, &b);
    out[e - first] = (double)(((e & 1) ? b : a) >> 11) * 0x1p-53;   /* [0,1), like Julia's rand() */
  }
}

/* ------------------------------------------------------------------------- */
/* Sorted uniforms as GROUPED order statistics (oracle-defined; DESIGN.md "Resampling arithmetic").
 * The M sorted uniforms behind the M iid categorical draws of particle_filter.jl:200 are generated
 * group by group, ORC_GROUP = 256 output slots per group:
 *   - the order statistic that opens group j sits at A_j / S_tot, where A_0 = head ~ Exp(1),
 *     A_{j+1} = A_j + g_j and g_j ~ Gamma(r_j) (r_j = slots of group j: 256, fewer in the last group), because the
 *     sum of r consecutive Exp(1) spacings is Gamma(r); S_tot = A_{n_groups};
 *   - given the two order statistics that bracket a group, the r_j - 1 draws between them are iid uniform on that
 *     interval (Markov property of order statistics) and are stored in the order they are drawn.
 * So the multiset of all M values is an exact sample of M iid uniforms, every group's ancestors lie between the
 * ancestors of its bracketing order statistics, and no per-draw logarithm or global prefix sum is needed.
 * Gaps are fixed-point integers (scale 2^20) so that their prefix sums are associative (shard-independent).
 * Gamma variates: Marsaglia & Tsang (2000), one normal + one uniform per attempt, Philox stream ORC_STREAM_GAP,
 * call (group << 5 | 2*attempt) for the normal (cos branch of Box-Muller) and the next call for the uniform. */
/* ------------------------------------------------------------------------- */
#define ORC_GROUP 256
#define ORC_GAP_SCALE 1048576.0
uint64_t orc_gap_variate(uint64_t seed, uint32_t rho, uint64_t group, uint32_t shape) {
  const double d = (double)shape - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  double v = 1.0;                                           /* after 16 rejections (probability < 1e-20): the mean */
  for (int attempt = 0; attempt < 16; ++attempt) {
    const uint64_t call = (group << 5) | (uint64_t)(2 * attempt);
    uint64_t a, b; double z, z1;
    philox_pair(seed, call, rho, ORC_STREAM_GAP, &a, &b);
    box_muller(a, b, &z, &z1);
    const double w = 1.0 + c * z;
    if (!(w > 0.0)) continue;
    const double w3 = (w * w) * w;
    philox_pair(seed, call + 1, rho, ORC_STREAM_GAP, &a, &b);
    const double u = ((double)(a >> 11) + 0.5) * 0x1p-53;
    const double lhs = orc_log(u);
    const double zz = (0.5 * z) * z;
    const double rhs = ((zz + d) - d * w3) + d * orc_log(w3);
    if (lhs < rhs) { v = w3; break; }
  }
  return (uint64_t)((d * v) * ORC_GAP_SCALE);
}
/* gaps g_j of groups [first, first+count) of an event with m draws (0 beyond the last group) */
void orc_fill_gaps(uint64_t seed, uint32_t rho, uint64_t m, uint64_t first, uint64_t count, uint64_t* out) {
  for (uint64_t j = first; j < first + count; ++j) {
    const uint64_t k0 = j * ORC_GROUP;
    out[j - first] = k0 < m ? orc_gap_variate(seed, rho, j, (uint32_t)(m - k0 < ORC_GROUP ? m - k0 : ORC_GROUP)) : 0;
  }
}
/* A_0: the Exp(1) gap below the first order statistic, drawn as the group one past the last */
uint64_t orc_gap_head(uint64_t seed, uint32_t rho, uint64_t m) {
  return orc_gap_variate(seed, rho, (m + ORC_GROUP - 1) / ORC_GROUP, 1);
}

/* ------------------------------------------------------------------------- */
/* distributions (src/modeling_library/distributions/NAME.jl)                    */
/* ------------------------------------------------------------------------- */
/* normal.jl:56-60 */
double orc_logpdf_normal(double x, double mu, double std) {
  double var = std * std;
  double diff = x - mu;
  return -(diff * diff) / (2.0 * var) - 0.5 * orc_log(2.0 * GM_PI * var);
}
/* normal.jl:96  random(::Normal, mu, std) = mu + std * randn() */
static double random_normal(double mu, double std, double z) { return mu + std * z; }

/* categorical.jl:10-12 (x is 1-based) */
double orc_logpdf_categorical(int64_t x, const double* probs, int64_t n) {
  return (x > 0 && x <= n) ? orc_log(probs[x - 1]) : -INFINITY;
}
/* categorical.jl:20-22 -> Distributions.jl 0.24.10 (Manifest.toml:71-75, not vendored)
 * rand(::DiscreteNonParametric): linear inverse-CDF scan
 *   cp = p[1]; i = 1; while cp <= draw && i < n: cp += p[i += 1]; return i   (restated from memory) */
static int64_t random_categorical(const double* probs, int64_t n, double draw) {
  double cp = probs[0];
  int64_t i = 1;
  while (cp <= draw && i < n) { cp += probs[i]; i += 1; }
  return i;
}
/* uniform_continuous.jl:12-14 */
double orc_logpdf_uniform(double x, double low, double high) {
  return (x >= low && x <= high) ? -orc_log(high - low) : -INFINITY;
}
/* bernoulli.jl:10-12 */
double orc_logpdf_bernoulli(int x, double p) { return x ? orc_log(p) : orc_log(1. - p); }

/* ------------------------------------------------------------------------- */
/* inference.jl:3-11, particle_filter.jl:3-12                                 */
/* ------------------------------------------------------------------------- */
/* Julia's sum() over a Vector is pairwise with a 1024-element base case. */
static double pairwise_sum_exp(const double* a, int64_t lo, int64_t hi, double mx, double scale) {
  if (hi - lo <= 1024) {
    double s = 0.0;
    for (int64_t i = lo; i < hi; ++i) s += orc_exp(scale * a[i] - mx);
    return s;
  }
  int64_t mid = lo + ((hi - lo) >> 1);
  return pairwise_sum_exp(a, lo, mid, mx, scale) + pairwise_sum_exp(a, mid, hi, mx, scale);
}
static double logsumexp_scaled(const double* arr, int64_t n, double scale) {
  double max_arr = -INFINITY;
  for (int64_t i = 0; i < n; ++i) { double v = scale * arr[i]; if (v > max_arr || v != v) max_arr = v; }
  if (max_arr == -INFINITY) return -INFINITY;
  return max_arr + orc_log(pairwise_sum_exp(arr, 0, n, max_arr, scale));
}
double orc_logsumexp(const double* arr, int64_t n) { return logsumexp_scaled(arr, n, 1.0); }
double orc_logsumexp2(double x1, double x2) {
  double m = x1 > x2 ? x1 : x2;
  return m == -INFINITY ? m : m + orc_log(orc_exp(x1 - m) + orc_exp(x2 - m));
}
/* particle_filter.jl:3-6: log_ess = -logsumexp(2. * log_normalized_weights) */
double orc_effective_sample_size(const double* lnw, int64_t n) {
  return orc_exp(-logsumexp_scaled(lnw, n, 2.0));
}

/* ------------------------------------------------------------------------- */
/* integer resampling arithmetic (oracle-defined where the reference defers   */
/* to Distributions.jl; DESIGN.md "Resampling arithmetic")                    */
/* ------------------------------------------------------------------------- */
int orc_weight_shift(uint64_t n_global) {
  int lg = 0;
  while (((uint64_t)1 << lg) < n_global) ++lg;
  int k = 62 - lg;
  return k > 52 ? 52 : k;
}
void orc_quantise_weights(const double* lw, int64_t n, uint64_t n_global, uint64_t* q, double* max_out) {
  double m = -INFINITY;
  for (int64_t i = 0; i < n; ++i) if (lw[i] > m) m = lw[i];
  const double scale = gm_pow2(orc_weight_shift(n_global));
  for (int64_t i = 0; i < n; ++i) q[i] = (uint64_t)floor(orc_exp(lw[i] - m) * scale);
  if (max_out) *max_out = m;
}
static int64_t upper_bound_u64(const uint64_t* cdf, int64_t n, uint64_t T) {  /* min{i: cdf[i] > T} */
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (cdf[mid] > T) hi = mid; else lo = mid + 1; }
  return lo;
}
void orc_search_iid(const uint64_t* cdf, int64_t n, const double* u, int64_t m, int64_t* anc) {
  const uint64_t total = cdf[n - 1];
  for (int64_t j = 0; j < m; ++j) {
    uint64_t t = (uint64_t)floor(u[j] * 0x1p53);
    uint64_t T = (uint64_t)(((u128)t * total) >> 53);
    int64_t i = upper_bound_u64(cdf, n, T);
    anc[j] = i < n ? i : n - 1;
  }
}
/* Ancestors of the m draws of event rho against the integer CDF (see "grouped order statistics" above).
 * Position x (gap units) maps to the integer threshold thr(x) = min(trunc((double)x * ratio), C_N - 1),
 * ratio = (double)C_N / (double)S_tot. Group j = slots [256 j, 256 j + r_j) is bracketed by TL = thr(A_j) and
 * TH = thr(A_j + g_j), i.e. by the CDF positions p_lo = min{i : C_i > TL} and p_hi = min(min{i : C_i > TH}, n - 1).
 *   slot k that opens the group:  anc_k = p_lo  (the order statistic itself)
 *   any other slot: w_k = 32-bit word (k & 3) of Philox call k >> 2 of stream ORC_STREAM_RESAMPLE places the draw at
 *                   TL + w_k (TH - TL) / 2^32; compared in the domain of the 32-bit words:
 *                   anc_k = p_lo + #{p in [p_lo, p_hi) : K_p < w_k},
 *                   K_p = (C_p - TL) 2^32 / (TH - TL) in integer arithmetic: with D = TH - TL, lz = clz64(D),
 *                   Dn = (D << lz) >> 32 (the top 32 bits, in [2^31, 2^32)), M = trunc((2^63 - 2^10) / (double)Dn) and
 *                   c' = ((C_p - TL) << lz) >> 32 (<= Dn):  K_p = (c' * M) >> 31  (< 2^32; all keys 0 when TH <= TL)
 * (K_p is, up to a few units of 2^-32, the largest word that still selects a particle <= p; it is non-decreasing in
 * p, so the count is a binary search. No per-draw arithmetic is left: two integer multiplies per CDF entry.) */
static uint64_t thr_of(uint64_t x, double ratio, uint64_t total) {
  uint64_t T = (uint64_t)((double)x * ratio);
  return T < total ? T : total - 1;
}
/* scale of a bracket: lz in the high word, M in the low word; 0 when TH <= TL */
uint64_t orc_bracket_scale(uint64_t tl, uint64_t th) {
  if (th <= tl) return 0;
  const uint64_t d = th - tl;
  const int lz = __builtin_clzll(d);
  const uint32_t dn = (uint32_t)((d << lz) >> 32);
  const uint32_t mul = (uint32_t)(9223372036854774784.0 / (double)dn);      /* 2^63 - 2^10, exactly a double */
  return ((uint64_t)lz << 32) | mul;
}
uint32_t orc_bracket_key(uint64_t c_minus_tl, uint64_t scale) {
  const int lz = (int)(scale >> 32);
  const uint32_t mul = (uint32_t)scale;
  const uint32_t cn = (uint32_t)((c_minus_tl << lz) >> 32);
  return (uint32_t)(((uint64_t)cn * mul) >> 31);
}
void orc_search_sorted(const uint64_t* cdf, int64_t n, uint64_t seed, uint32_t rho, int64_t m, int64_t* anc) {
  const uint64_t total = cdf[n - 1];
  const int64_t n_groups = (m + ORC_GROUP - 1) / ORC_GROUP;
  uint64_t* g = (uint64_t*)malloc(sizeof(uint64_t) * (n_groups + 1));
  orc_fill_gaps(seed, rho, (uint64_t)m, 0, (uint64_t)n_groups, g);
  const uint64_t head = orc_gap_head(seed, rho, (uint64_t)m);
  uint64_t stot = head;
  for (int64_t j = 0; j < n_groups; ++j) stot += g[j];
  const double ratio = stot ? (double)total / (double)stot : 0.0;
  uint64_t A = head;
  for (int64_t j = 0; j < n_groups; ++j) {
    const uint64_t TL = thr_of(A, ratio, total), TH = thr_of(A + g[j], ratio, total);
    const int64_t p_lo = upper_bound_u64(cdf, n, TL);
    int64_t p_hi = upper_bound_u64(cdf, n, TH);
    if (p_hi > n - 1) p_hi = n - 1;
    const uint64_t r32 = orc_bracket_scale(TL, TH);
    for (int64_t k = j * ORC_GROUP; k < (j + 1) * ORC_GROUP && k < m; ++k) {
      int64_t pos = p_lo;
      if (k > j * ORC_GROUP) {
        uint32_t ctr[4] = { (uint32_t)((uint64_t)k >> 2), (uint32_t)(((uint64_t)k >> 2) >> 32), rho, ORC_STREAM_RESAMPLE };
        uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
        uint32_t o[4];
        orc_philox4x32_10(ctr, key, o);
        const uint32_t w = o[k & 3];
        int64_t lo = p_lo, hi = p_hi;                      /* first p in [p_lo, p_hi) with K_p >= w, else p_hi */
        while (lo < hi) {
          const int64_t mid = lo + ((hi - lo) >> 1);
          if (orc_bracket_key(cdf[mid] - TL, r32) < w) lo = mid + 1; else hi = mid;
        }
        pos = lo;
      }
      anc[k] = pos;
    }
    A += g[j];
  }
  free(g);
}

/* ------------------------------------------------------------------------- */
/* model families                                                             */
/* ------------------------------------------------------------------------- */
#define MAXP 4096
struct orc_pf {
  int family;
  int n_params;
  double params[MAXP];
  int64_t N;
  int D;
  uint64_t seed;
  int keep_history;
  int nthreads;
  int64_t T;            /* number of time steps in the traces (0 before init) */
  double* cur;          /* D x N column-major current latent */
  double* nxt;
  double** hist;        /* hist[t-1] = D x N, physically permuted at every resample */
  double** obs_hist;    /* obs_hist[t-1] = N sampled observation choices of an UNOBSERVED step (NULL if it was observed); permuted alike */
  int64_t hist_cap;
  double* lw;
  double log_ml_est;
  int64_t* parents;
  uint32_t n_resamples;
  uint32_t n_sample_calls;
};

static int family_dim(int family) {
  switch (family) {
    case ORC_HMM: case ORC_LGSSM: case ORC_SV: case ORC_NORMAL_NORMAL: return 1;
    case ORC_BEARINGS: return 4;
    case ORC_REGRESSION: return 2;
    case ORC_OUTLIER_REGRESSION: return 4 + ORC_OUTLIER_ZWORDS;
    case ORC_UNIFORM_NORMAL: return 1;
    default: return -1;
  }
}
static int family_normals(int family, int proposal, int is_init) {
  (void)proposal;
  switch (family) {
    case ORC_HMM: return 0;
    case ORC_LGSSM: case ORC_SV: case ORC_NORMAL_NORMAL: return 1;
    case ORC_BEARINGS: return is_init ? 4 : 2;
    case ORC_REGRESSION: return 2;
    case ORC_OUTLIER_REGRESSION: return 4;
    case ORC_UNIFORM_NORMAL: return 0;
    default: return -1;
  }
}
static int family_uniforms(int family, int proposal, int is_init) {
  (void)proposal; (void)is_init;
  return family == ORC_HMM ? 1 : 0;
}
int orc_pf_state_dim(const orc_pf* pf) { return pf->D; }
int orc_pf_num_normals(const orc_pf* pf, int proposal, int is_init) { return family_normals(pf->family, proposal, is_init); }
int orc_pf_num_uniforms(const orc_pf* pf, int proposal, int is_init) { return family_uniforms(pf->family, proposal, is_init); }

/* --- HMM: test/inference/particle_filter.jl:52-78 (dynamic-DSL kernel + Unfold), proposals :104-127 ---
 * params = [K, V, prior[K], trans[K][K] (row zp: P(z | zp) = transition_dists[:,zp]),
 *           emis[K][V] (row z: P(x | z) = emission_dists[:,z])]; obs = [x] (1-based) */
typedef struct { int K, V; const double *prior, *trans, *emis; } hmm_t;
static hmm_t hmm_view(const double* p) {
  hmm_t h; h.K = (int)p[0]; h.V = (int)p[1]; h.prior = p + 2; h.trans = h.prior + h.K; h.emis = h.trans + h.K * h.K; return h;
}
/* one particle, one time step. zp = previous latent (1-based; ignored when is_init). */
static double hmm_particle(const hmm_t* h, int is_init, int proposal, int64_t zp, int64_t x, double u, double* z_out) {
  const double* pz = is_init ? h->prior : h->trans + (zp - 1) * h->K;   /* categorical(prior) / categorical(transition_dists[:,prev_z]) */
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    /* generate(): :z unconstrained -> random(); :x constrained -> weight += logpdf (dynamic/generate.jl:26-41) */
    int64_t z = random_categorical(pz, h->K, u);
    double weight = 0.0;
    weight += orc_logpdf_categorical(x, h->emis + (z - 1) * h->V, h->V);
    *z_out = (double)z;
    return weight;
  }
  /* custom: dist = pz .* emission_dists[x,:]; z ~ categorical(dist ./ sum(dist))  (:104-108, :117-127) */
  double dist[64], s = 0.0;
  for (int k = 0; k < h->K; ++k) { dist[k] = pz[k] * h->emis[k * h->V + (x - 1)]; }
  for (int k = 0; k < h->K; ++k) s += dist[k];
  for (int k = 0; k < h->K; ++k) dist[k] = dist[k] / s;
  int64_t z = random_categorical(dist, h->K, u);
  double q_score = orc_logpdf_categorical(z, dist, h->K);          /* propose()/simulate() score */
  double model_w = 0.0;                                            /* generate() with :z and :x constrained */
  model_w += orc_logpdf_categorical(z, pz, h->K);
  model_w += orc_logpdf_categorical(x, h->emis + (z - 1) * h->V, h->V);
  *z_out = (double)z;
  return model_w - q_score;                                        /* particle_filter.jl:87, trace_translators.jl:800 */
}

/* --- LGSSM: static kernel of test/modeling_library/unfold.jl:5-8 plus an observation choice ---
 *   x_init ~ normal(m0, s0);            y_init ~ normal(c * x_init, r)
 *   x      ~ normal(x_prev * a + b, q); y      ~ normal(c * x, r)
 * params = [m0, s0, a, b, q, c, r]; obs = [y]
 * custom proposal (locally optimal Gaussian, written as a Gen user would):
 *   prec = 1/(sd*sd) + (c*c)/(r*r); var = 1/prec; mu = var * (mean/(sd*sd) + (c*y)/(r*r)); x ~ normal(mu, sqrt(var)) */
static double lgssm_particle(const double* p, int is_init, int proposal, double xp, double y, double z, double* x_out) {
  const double m0 = p[0], s0 = p[1], a = p[2], b = p[3], q = p[4], c = p[5], r = p[6];
  const double mean = is_init ? m0 : xp * a + b;
  const double sd = is_init ? s0 : q;
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    double x = random_normal(mean, sd, z);
    double weight = 0.0;
    weight += orc_logpdf_normal(y, c * x, r);
    *x_out = x;
    return weight;
  }
  double prec = 1.0 / (sd * sd) + (c * c) / (r * r);
  double var = 1.0 / prec;
  double mu = var * (mean / (sd * sd) + (c * y) / (r * r));
  double sdq = sqrt(var);
  double x = random_normal(mu, sdq, z);
  double q_score = orc_logpdf_normal(x, mu, sdq);
  double model_w = 0.0;
  model_w += orc_logpdf_normal(x, mean, sd);
  model_w += orc_logpdf_normal(y, c * x, r);
  *x_out = x;
  return model_w - q_score;
}

/* --- stochastic volatility (static kernel + Unfold pattern of examples/pmmh/model.jl:40-50) ---
 *   h_init ~ normal(mu, sigma / sqrt(1 - phi*phi));  h ~ normal(mu + phi * (h_prev - mu), sigma)
 *   y ~ normal(0, exp(h / 2))
 * params = [mu, phi, sigma]; obs = [y] */
static double sv_particle(const double* p, int is_init, double hp, double y, double z, double* h_out) {
  const double mu = p[0], phi = p[1], sigma = p[2];
  double mean = is_init ? mu : mu + phi * (hp - mu);
  double sd = is_init ? sigma / sqrt(1.0 - phi * phi) : sigma;
  double h = random_normal(mean, sd, z);
  double weight = 0.0;
  weight += orc_logpdf_normal(y, 0.0, orc_exp(h / 2.0));
  *h_out = h;
  return weight;
}

/* --- bearings-only tracking (custom-proposal pattern of examples/planning/filtering.jl:79-91,175-187) ---
 * state s = (x, vx, y, vy); params = [m[4], sd[4], sigma_w, sigma_theta]; obs = [bearing]
 *   init: x ~ normal(m1,sd1), vx ~ normal(m2,sd2), y ~ normal(m3,sd3), vy ~ normal(m4,sd4)
 *   step: wx ~ normal(0, sigma_w), wy ~ normal(0, sigma_w)
 *         x = xp + vxp + 0.5*wx; vx = vxp + wx; y = yp + vyp + 0.5*wy; vy = vyp + wy
 *   bearing ~ normal(atan(y, x), sigma_theta)        (plain `normal`: not wrapped)
 * custom step proposal (independent Gaussians from a one-step EKF update of (wx, wy)):
 *   xb = xp + vxp; yb = yp + vyp; rho2 = xb*xb + yb*yb; nu = bearing - atan(yb, xb)
 *   hx = -yb / rho2; hy = xb / rho2
 *   S  = 0.25*sw2*(hx*hx + hy*hy) + st2
 *   kx = 0.5*sw2*hx / S; ky = 0.5*sw2*hy / S
 *   wx ~ normal(kx*nu, sqrt(sw2*(1 - 0.5*kx*hx))); wy ~ normal(ky*nu, sqrt(sw2*(1 - 0.5*ky*hy))) */
static double bearings_particle(const double* p, int is_init, int proposal, const double sp[4], double obs,
                                const double* z, double s_out[4]) {
  const double sw = p[8], st = p[9];
  double weight;
  if (is_init) {
    double x = random_normal(p[0], p[4], z[0]);
    double vx = random_normal(p[1], p[5], z[1]);
    double y = random_normal(p[2], p[6], z[2]);
    double vy = random_normal(p[3], p[7], z[3]);
    weight = 0.0;
    weight += orc_logpdf_normal(obs, orc_atan2(y, x), st);
    s_out[0] = x; s_out[1] = vx; s_out[2] = y; s_out[3] = vy;
    return weight;
  }
  double wx, wy, q_score = 0.0, model_w = 0.0;
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    wx = random_normal(0.0, sw, z[0]);
    wy = random_normal(0.0, sw, z[1]);
  } else {
    double sw2 = sw * sw, st2 = st * st;
    double xb = sp[0] + sp[1], yb = sp[2] + sp[3];
    double rho2 = xb * xb + yb * yb;
    double nu = obs - orc_atan2(yb, xb);
    double hx = -yb / rho2, hy = xb / rho2;
    double S = 0.25 * sw2 * (hx * hx + hy * hy) + st2;
    double kx = 0.5 * sw2 * hx / S, ky = 0.5 * sw2 * hy / S;
    double mx = kx * nu, my = ky * nu;
    double sx = sqrt(sw2 * (1.0 - 0.5 * kx * hx)), sy = sqrt(sw2 * (1.0 - 0.5 * ky * hy));
    wx = random_normal(mx, sx, z[0]);
    wy = random_normal(my, sy, z[1]);
    q_score += orc_logpdf_normal(wx, mx, sx);
    q_score += orc_logpdf_normal(wy, my, sy);
    model_w += orc_logpdf_normal(wx, 0.0, sw);
    model_w += orc_logpdf_normal(wy, 0.0, sw);
  }
  double x = sp[0] + sp[1] + 0.5 * wx;
  double vx = sp[1] + wx;
  double y = sp[2] + sp[3] + 0.5 * wy;
  double vy = sp[3] + wy;
  model_w += orc_logpdf_normal(obs, orc_atan2(y, x), st);
  s_out[0] = x; s_out[1] = vx; s_out[2] = y; s_out[3] = vy;
  return model_w - q_score;
}

/* --- IS families ---
 * regression (examples/regression/quickstart.jl:3-9):
 *   slope ~ normal(0, sd_s); intercept ~ normal(0, sd_i); y_i ~ normal(slope * x_i + intercept, sd_n)
 *   params = [n, sd_s, sd_i, sd_n, xs[n]]; obs = ys[n]; custom proposal params = [mu_s, sd_s', mu_i, sd_i']
 * normal-normal (test/inference/importance_sampling.jl:3-12):
 *   x ~ normal(mu0, sd0); y ~ normal(x, sd_y); params = [mu0, sd0, sd_y]; obs=[y]; proposal params = [mu_q, sd_q] */
static double regression_sample(const double* p, const double* ys, int proposal, const double* pp, const double* z, double* lat) {
  const int n = (int)p[0];
  const double* xs = p + 4;
  double slope, intercept, prop_w = 0.0, model_w = 0.0;
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    slope = random_normal(0.0, p[1], z[0]);
    intercept = random_normal(0.0, p[2], z[1]);
  } else {
    slope = random_normal(pp[0], pp[1], z[0]);
    intercept = random_normal(pp[2], pp[3], z[1]);
    prop_w += orc_logpdf_normal(slope, pp[0], pp[1]);
    prop_w += orc_logpdf_normal(intercept, pp[2], pp[3]);
    model_w += orc_logpdf_normal(slope, 0.0, p[1]);
    model_w += orc_logpdf_normal(intercept, 0.0, p[2]);
  }
  for (int i = 0; i < n; ++i) model_w += orc_logpdf_normal(ys[i], slope * xs[i] + intercept, p[3]);
  lat[0] = slope; lat[1] = intercept;
  return model_w - prop_w;                    /* importance.jl:27 (prop_w = 0) / :46 */
}
static double normal_normal_sample(const double* p, const double* obs, int proposal, const double* pp, const double* z, double* lat) {
  double x, prop_w = 0.0, model_w = 0.0;
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    x = random_normal(p[0], p[1], z[0]);
  } else {
    x = random_normal(pp[0], pp[1], z[0]);
    prop_w += orc_logpdf_normal(x, pp[0], pp[1]);
    model_w += orc_logpdf_normal(x, p[0], p[1]);
  }
  model_w += orc_logpdf_normal(obs[0], x, p[2]);
  lat[0] = x;
  return model_w - prop_w;
}

/* examples/regression/static_model.jl:3-23: a static model whose data points are a Map of the static `datum`:
 *   inlier_log_std ~ normal(0, sd) :log_inlier_std; outlier_log_std ~ normal(0, sd) :log_outlier_std;
 *   inlier_std = exp(inlier_log_std); outlier_std = exp(outlier_log_std); slope ~ normal(0, sd); intercept ~ normal(0, sd);
 *   datum i (:data => i): is_outlier ~ bernoulli(prob) :z;  std = ifelse(is_outlier, inlier_std, outlier_std)  [literally :6];
 *                         y ~ normal(x_i * slope + intercept, std) :y   (constrained)
 * params = [n, prob, sd, xs[n]] (the reference: prob = 0.5, sd = 2). generate(): unconstrained choices are sampled in
 * trace order (static_ir/generate.jl:36-42; Map visits the data points in order, map/generate.jl), the weight is the sum
 * of the logpdfs of the constrained :y's. random(bernoulli, p) = rand() < p (bernoulli.jl:19).
 * latents: [log_inlier_std, log_outlier_std, slope, intercept, z bits packed 32 per word as exact integers]. */
static double outlier_regression_sample(const double* p, const double* ys, const double* z, const double* u, double* lat) {
  const int n = (int)p[0];
  const double prob = p[1], sd = p[2];
  const double* xs = p + 3;
  const double inlier_log_std = random_normal(0.0, sd, z[0]);
  const double outlier_log_std = random_normal(0.0, sd, z[1]);
  const double inlier_std = orc_exp(inlier_log_std), outlier_std = orc_exp(outlier_log_std);
  const double slope = random_normal(0.0, sd, z[2]);
  const double intercept = random_normal(0.0, sd, z[3]);
  double w = 0.0;
  uint32_t words[ORC_OUTLIER_ZWORDS] = {0};
  for (int i = 0; i < n; ++i) {
    const int is_outlier = u[i] < prob;                                        /* bernoulli.jl:19 */
    const double std = is_outlier ? inlier_std : outlier_std;                  /* static_model.jl:6 */
    w += orc_logpdf_normal(ys[i], xs[i] * slope + intercept, std);
    if (is_outlier) words[i >> 5] |= 1u << (i & 31);
  }
  lat[0] = inlier_log_std; lat[1] = outlier_log_std; lat[2] = slope; lat[3] = intercept;
  for (int k = 0; k < ORC_OUTLIER_ZWORDS; ++k) lat[4 + k] = (double)words[k];
  return w;
}
/* x ~ uniform(lo, hi) (uniform_continuous.jl:12-23: random = rand() * (high - low) + low); y ~ normal(x, sd_y).
 * params = [lo, hi, sd_y]; obs = [y]; custom proposal x ~ uniform(pp[0], pp[1]). */
static double uniform_normal_sample(const double* p, const double* obs, int proposal, const double* pp, const double* u, double* lat) {
  double x, prop_w = 0.0, model_w = 0.0;
  if (proposal == ORC_PROPOSAL_DEFAULT) {
    x = u[0] * (p[1] - p[0]) + p[0];
  } else {
    x = u[0] * (pp[1] - pp[0]) + pp[0];
    prop_w += orc_logpdf_uniform(x, pp[0], pp[1]);
    model_w += orc_logpdf_uniform(x, p[0], p[1]);
  }
  model_w += orc_logpdf_normal(obs[0], x, p[2]);
  lat[0] = x;
  return model_w - prop_w;
}

/* ------------------------------------------------------------------------- */
/* particle filter                                                            */
/* ------------------------------------------------------------------------- */
orc_pf* orc_pf_create(int family, const double* params, int n_params, int64_t N, uint64_t seed, int keep_history, int nthreads) {
  int D = family_dim(family);
  if (D < 0 || family == ORC_REGRESSION || family == ORC_NORMAL_NORMAL || family == ORC_OUTLIER_REGRESSION || family == ORC_UNIFORM_NORMAL) { snprintf(g_err, sizeof g_err, "family %d is not a state-space family", family); return NULL; }
  if (n_params > MAXP || N <= 0) { snprintf(g_err, sizeof g_err, "bad arguments"); return NULL; }
  orc_pf* pf = (orc_pf*)calloc(1, sizeof(orc_pf));
  pf->family = family; pf->n_params = n_params; memcpy(pf->params, params, sizeof(double) * n_params);
  pf->N = N; pf->D = D; pf->seed = seed; pf->keep_history = keep_history; pf->nthreads = nthreads > 0 ? nthreads : 1;
  pf->cur = (double*)malloc(sizeof(double) * D * N);
  pf->nxt = (double*)malloc(sizeof(double) * D * N);
  pf->lw = (double*)malloc(sizeof(double) * N);
  pf->parents = (int64_t*)malloc(sizeof(int64_t) * N);
  return pf;
}
void orc_pf_destroy(orc_pf* pf) {
  if (!pf) return;
  for (int64_t t = 0; t < pf->T && pf->hist; ++t) free(pf->hist[t]);
  for (int64_t t = 0; t < pf->T && pf->obs_hist; ++t) free(pf->obs_hist[t]);
  free(pf->obs_hist);
  free(pf->hist); free(pf->cur); free(pf->nxt); free(pf->lw); free(pf->parents); free(pf);
}
static void push_history(orc_pf* pf, double* sampled_obs) {
  if (!pf->keep_history) { free(sampled_obs); return; }
  if (pf->T > pf->hist_cap) {
    pf->hist_cap = pf->hist_cap ? pf->hist_cap * 2 : 16;
    if (pf->hist_cap < pf->T) pf->hist_cap = pf->T;
    pf->hist = (double**)realloc(pf->hist, sizeof(double*) * pf->hist_cap);
    pf->obs_hist = (double**)realloc(pf->obs_hist, sizeof(double*) * pf->hist_cap);
  }
  pf->hist[pf->T - 1] = (double*)malloc(sizeof(double) * pf->D * pf->N);
  memcpy(pf->hist[pf->T - 1], pf->cur, sizeof(double) * pf->D * pf->N);
  pf->obs_hist[pf->T - 1] = sampled_obs;
}
/* The observation choice of an UNOBSERVED step: static_ir/generate.jl:36-42 / dynamic/generate.jl:26-33 sample an
 * unconstrained choice with random(dist, args...) and add nothing to the weight. One draw per particle from Philox
 * stream ORC_STREAM_OBS of the step (element = particle index): a normal for the continuous families, a uniform for the
 * HMM's categorical emission. lat = the particle's NEW latent (D values, stride N). */
static double sample_observation(const orc_pf* pf, const double* lat, int64_t stride, double draw) {
  const double* p = pf->params;
  switch (pf->family) {
    case ORC_LGSSM: return random_normal(p[5] * lat[0], p[6], draw);                       /* y ~ normal(c*x, r) */
    case ORC_SV: return random_normal(0.0, orc_exp(lat[0] / 2.0), draw);                    /* y ~ normal(0, exp(h/2)) */
    case ORC_BEARINGS: return random_normal(orc_atan2(lat[2 * stride], lat[0]), p[9], draw);  /* bearing ~ normal(atan(y, x), sigma_theta) */
    case ORC_HMM: { hmm_t h = hmm_view(p); return (double)random_categorical(h.emis + ((int64_t)lat[0] - 1) * h.V, h.V, draw); }
  }
  return 0.0;
}

/* one pass of `for i=1:num_particles` for init (particle_filter.jl:84-88,103-105) or
 * step (:143-146,165-172). Adds the increment to lw (init: sets it). */
static int propagate(orc_pf* pf, int is_init, const double* obs, int n_obs, int proposal,
                     const double* zrep, const double* urep) {
  const int64_t N = pf->N; const int D = pf->D;
  const int nz = family_normals(pf->family, proposal, is_init), nu = family_uniforms(pf->family, proposal, is_init);
  const int unobserved = (n_obs < 1 || !obs);
  const double dummy_obs[1] = {1.0};
  if (unobserved) {
    /* no constraint at this step: the latent is sampled as usual, the observation choice is sampled too, weight += 0 */
    if (proposal != ORC_PROPOSAL_DEFAULT) ORC_FAIL("the catalogue's custom proposals condition on the observation");
    obs = dummy_obs;
  }
  if (proposal != ORC_PROPOSAL_DEFAULT && pf->family == ORC_SV) ORC_FAIL("no custom proposal for this family");
  const uint32_t t = (uint32_t)(pf->T + 1);
  double* Z = NULL; double* U = NULL;
  if (nz) { if (zrep) Z = (double*)zrep; else { Z = (double*)malloc(sizeof(double) * nz * N); } }
  if (nu) { if (urep) U = (double*)urep; else { U = (double*)malloc(sizeof(double) * nu * N); } }
  if (nz && !zrep) {
    #pragma omp parallel for num_threads(pf->nthreads) schedule(static)
    for (int64_t blk = 0; blk < (nz * N + 4095) / 4096; ++blk) {
      uint64_t f = (uint64_t)blk * 4096, c = (uint64_t)(nz * N) - f; if (c > 4096) c = 4096;
      orc_fill_normals(pf->seed, t, f, c, Z + f);
    }
  }
  if (nu && !urep) {
    #pragma omp parallel for num_threads(pf->nthreads) schedule(static)
    for (int64_t blk = 0; blk < (nu * N + 4095) / 4096; ++blk) {
      uint64_t f = (uint64_t)blk * 4096, c = (uint64_t)(nu * N) - f; if (c > 4096) c = 4096;
      orc_fill_uniforms(pf->seed, t, ORC_STREAM_UNIFORM, f, c, U + f);
    }
  }
  const double* p = pf->params;
  hmm_t hv; if (pf->family == ORC_HMM) hv = hmm_view(p);
  #pragma omp parallel for num_threads(pf->nthreads) schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    double w = 0.0;
    switch (pf->family) {
      case ORC_HMM:
        w = hmm_particle(&hv, is_init, proposal, is_init ? 0 : (int64_t)pf->cur[i], (int64_t)obs[0], U[i], &pf->nxt[i]);
        break;
      case ORC_LGSSM:
        w = lgssm_particle(p, is_init, proposal, pf->cur[i], obs[0], Z[i], &pf->nxt[i]);
        break;
      case ORC_SV:
        w = sv_particle(p, is_init, pf->cur[i], obs[0], Z[i], &pf->nxt[i]);
        break;
      case ORC_BEARINGS: {
        double sp[4], so[4];
        for (int d = 0; d < 4; ++d) sp[d] = pf->cur[d * N + i];
        w = bearings_particle(p, is_init, proposal, sp, obs[0], Z + (int64_t)nz * i, so);
        for (int d = 0; d < 4; ++d) pf->nxt[d * N + i] = so[d];
      } break;
    }
    if (unobserved) w = 0.0;               /* generate/update weight = sum over CONSTRAINED choices: none */
    if (is_init) pf->lw[i] = w;            /* particle_filter.jl:87,104 */
    else pf->lw[i] += w;                   /* :145,171 */
  }
  (void)D;
  if (nz && !zrep) free(Z);
  if (nu && !urep) free(U);
  double* sampled = NULL;
  if (unobserved) {
    sampled = (double*)malloc(sizeof(double) * N);
    for (int64_t i = 0; i < N; ++i) {
      uint64_t a, b; double z0, z1, draw;
      philox_pair(pf->seed, (uint64_t)i >> 1, t, ORC_STREAM_OBS, &a, &b);
      if (pf->family == ORC_HMM) draw = (double)(((i & 1) ? b : a) >> 11) * 0x1p-53;
      else { box_muller(a, b, &z0, &z1); draw = (i & 1) ? z1 : z0; }
      sampled[i] = sample_observation(pf, pf->nxt + i, N, draw);
    }
  }
  double* tmp = pf->cur; pf->cur = pf->nxt; pf->nxt = tmp;   /* swap references, :148-151,174-177 */
  pf->T += 1;
  push_history(pf, sampled);
  return 0;
}

int orc_pf_init(orc_pf* pf, const double* obs, int n_obs, int proposal, const double* pp, int npp,
                const double* zrep, const double* urep) {
  (void)pp; (void)npp;
  if (pf->T != 0) ORC_FAIL("already initialised");
  int rc = propagate(pf, 1, obs, n_obs, proposal, zrep, urep);
  if (rc) return rc;
  pf->log_ml_est = 0.;                                        /* :90,107 */
  for (int64_t i = 0; i < pf->N; ++i) pf->parents[i] = i;     /* collect(1:num_particles), 0-based here */
  return 0;
}
int orc_pf_step(orc_pf* pf, const double* obs, int n_obs, int proposal, const double* pp, int npp,
                const double* zrep, const double* urep) {
  (void)pp; (void)npp;
  if (pf->T < 1) ORC_FAIL("not initialised");
  return propagate(pf, 0, obs, n_obs, proposal, zrep, urep);
}

int orc_pf_maybe_resample(orc_pf* pf, double ess_threshold, int scheme, const double* urep,
                          int* did, double* ess_out, double* log_total_out) {
  const int64_t N = pf->N; const int D = pf->D;
  /* (log_total_weight, log_normalized_weights) = normalize_weights(state.log_weights)   :192 */
  double log_total = orc_logsumexp(pf->lw, N);
  double* lnw = (double*)malloc(sizeof(double) * N);
  for (int64_t i = 0; i < N; ++i) lnw[i] = pf->lw[i] - log_total;
  double ess = orc_effective_sample_size(lnw, N);                                      /* :193 */
  free(lnw);
  int do_resample = ess < ess_threshold;                                               /* :194 */
  if (ess_out) *ess_out = ess;
  if (log_total_out) *log_total_out = log_total;
  if (did) *did = do_resample;
  if (!do_resample) return 0;
  if (!(log_total > -INFINITY) || log_total != log_total) ORC_FAIL("total weight is zero or not finite");
  /* weights = exp.(lnw); rand!(Categorical(weights / sum(weights)), parents)           :199-200
   * -> integer CDF (orc_quantise_weights) + inverse-CDF search */
  uint64_t* q = (uint64_t*)malloc(sizeof(uint64_t) * N);
  orc_quantise_weights(pf->lw, N, (uint64_t)N, q, NULL);
  int64_t* anc = pf->parents;
  if (scheme == ORC_RESAMPLE_MULTINOMIAL) {
    for (int64_t i = 1; i < N; ++i) q[i] += q[i - 1];
    if (urep) orc_search_iid(q, N, urep, N, anc);
    else {
      orc_search_sorted(q, N, pf->seed, pf->n_resamples, N, anc);
    }
  } else {
    /* residual (not in the reference): c_i = floor(N p_i) copies in index order, then
     * M = N - sum c_i multinomial draws on the residual fractions */
    uint64_t total = 0;
    for (int64_t i = 0; i < N; ++i) total += q[i];
    const double scale = ((double)N * 4294967296.0) / (double)total;
    int64_t Dn = 0;
    uint64_t* G = (uint64_t*)malloc(sizeof(uint64_t) * N);
    uint64_t g = 0;
    for (int64_t i = 0; i < N; ++i) {
      uint64_t e = (uint64_t)floor((double)q[i] * scale);
      uint64_t c = e >> 32;
      g += (e & 0xffffffffULL);
      G[i] = g;
      for (uint64_t j = 0; j < c; ++j) anc[Dn++] = i;
    }
    int64_t M = N - Dn;
    if (M > 0) {
      if (urep) orc_search_iid(G, N, urep, M, anc + Dn);
      else {
        orc_search_sorted(G, N, pf->seed, pf->n_resamples, M, anc + Dn);
      }
    }
    free(G);
  }
  free(q);
  pf->log_ml_est += log_total - orc_log((double)N);                                     /* :201 */
  /* new_traces[i] = traces[parents[i]]; log_weights[i] = 0.                            :202-205 */
  for (int d = 0; d < D; ++d)
    for (int64_t i = 0; i < N; ++i) pf->nxt[d * N + i] = pf->cur[d * N + anc[i]];
  if (pf->keep_history) {
    for (int64_t t = 0; t < pf->T; ++t) {
      double* h = pf->hist[t];
      double* tmp = (double*)malloc(sizeof(double) * D * N);
      for (int d = 0; d < D; ++d)
        for (int64_t i = 0; i < N; ++i) tmp[d * N + i] = h[d * N + anc[i]];
      memcpy(h, tmp, sizeof(double) * D * N);
      free(tmp);
      double* o = pf->obs_hist[t];
      if (o) {
        double* tmo = (double*)malloc(sizeof(double) * N);
        for (int64_t i = 0; i < N; ++i) tmo[i] = o[anc[i]];
        memcpy(o, tmo, sizeof(double) * N);
        free(tmo);
      }
    }
  }
  for (int64_t i = 0; i < N; ++i) pf->lw[i] = 0.;
  double* tmp = pf->cur; pf->cur = pf->nxt; pf->nxt = tmp;                              /* :207-210 */
  pf->n_resamples += 1;
  return 0;
}

/* particle_filter.jl:52-55 */
double orc_pf_log_ml_estimate(const orc_pf* pf) {
  return pf->log_ml_est + orc_logsumexp(pf->lw, pf->N) - orc_log((double)pf->N);
}
const double* orc_pf_log_weights(const orc_pf* pf) { return pf->lw; }
void orc_pf_set_log_weights(orc_pf* pf, const double* lw) { memcpy(pf->lw, lw, sizeof(double) * pf->N); }
const int64_t* orc_pf_parents(const orc_pf* pf) { return pf->parents; }
const double* orc_pf_state(const orc_pf* pf) { return pf->cur; }
int64_t orc_pf_num_steps(const orc_pf* pf) { return pf->T; }
int orc_pf_history(const orc_pf* pf, int64_t t, double* out) {
  if (!pf->keep_history) ORC_FAIL("history not kept");
  if (t < 1 || t > pf->T) ORC_FAIL("t out of range");
  memcpy(out, pf->hist[t - 1], sizeof(double) * pf->D * pf->N);
  return 0;
}

/* sampled observation choices of unobserved step t (in the particles' current order); fails if step t was observed */
int orc_pf_sampled_observation(const orc_pf* pf, int64_t t, double* out) {
  if (!pf->keep_history) ORC_FAIL("history not kept");
  if (t < 1 || t > pf->T) ORC_FAIL("t out of range");
  if (!pf->obs_hist[t - 1]) ORC_FAIL("step %lld was observed", (long long)t);
  memcpy(out, pf->obs_hist[t - 1], sizeof(double) * pf->N);
  return 0;
}

/* particle_filter.jl:62-70: weights = exp.(lnw); traces[categorical(weights)] num_samples times.
 * Integer CDF + one iid uniform per sample (stream SAMPLE, event = call counter). */
int orc_pf_sample_unweighted(orc_pf* pf, int64_t num_samples, const double* urep, int64_t* idx) {
  const int64_t N = pf->N;
  double log_total = orc_logsumexp(pf->lw, N);
  if (!(log_total > -INFINITY) || log_total != log_total) ORC_FAIL("total weight is zero or not finite");
  uint64_t* q = (uint64_t*)malloc(sizeof(uint64_t) * N);
  orc_quantise_weights(pf->lw, N, (uint64_t)N, q, NULL);
  for (int64_t i = 1; i < N; ++i) q[i] += q[i - 1];
  double* u = (double*)urep;
  if (!urep) {
    u = (double*)malloc(sizeof(double) * num_samples);
    orc_fill_uniforms(pf->seed, pf->n_sample_calls, ORC_STREAM_SAMPLE, 0, (uint64_t)num_samples, u);
  }
  orc_search_iid(q, N, u, num_samples, idx);
  if (!urep) free(u);
  free(q);
  pf->n_sample_calls += 1;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* importance sampling (importance.jl:20-52)                                  */
/* ------------------------------------------------------------------------- */
int orc_importance_sampling(int family, const double* params, int n_params, const double* obs, int n_obs,
                            int proposal, const double* pp, int npp, int64_t n, uint64_t seed,
                            const double* zrep, double* lat_out, double* lnw_out, double* lml_out, int nthreads) {
  (void)n_params; (void)npp;
  int D = family_dim(family);
  if (family != ORC_REGRESSION && family != ORC_NORMAL_NORMAL && family != ORC_OUTLIER_REGRESSION && family != ORC_UNIFORM_NORMAL)
    ORC_FAIL("family %d is not an importance-sampling family", family);
  if ((family == ORC_REGRESSION || family == ORC_OUTLIER_REGRESSION) && n_obs != (int)params[0]) ORC_FAIL("need one observation per data point");
  if (family == ORC_OUTLIER_REGRESSION && ((int)params[0] < 1 || (int)params[0] > 32 * ORC_OUTLIER_ZWORDS)) ORC_FAIL("1 <= n <= %d data points", 32 * ORC_OUTLIER_ZWORDS);
  if (family == ORC_OUTLIER_REGRESSION && proposal != ORC_PROPOSAL_DEFAULT) ORC_FAIL("the outlier model has no custom proposal in the catalogue");
  if ((family == ORC_NORMAL_NORMAL || family == ORC_UNIFORM_NORMAL) && n_obs != 1) ORC_FAIL("need exactly one observation");
  if (proposal != ORC_PROPOSAL_DEFAULT && !pp) ORC_FAIL("custom proposal needs parameters");
  const int nz = family_normals(family, proposal, 1);
  const int nu = family == ORC_OUTLIER_REGRESSION ? (int)params[0] : (family == ORC_UNIFORM_NORMAL ? 1 : 0);
  if (nthreads < 1) nthreads = 1;
  double* Z = (double*)zrep;
  if (!zrep && nz > 0) {
    Z = (double*)malloc(sizeof(double) * nz * n);
    #pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t blk = 0; blk < (nz * n + 4095) / 4096; ++blk) {
      uint64_t f = (uint64_t)blk * 4096, c = (uint64_t)(nz * n) - f; if (c > 4096) c = 4096;
      orc_fill_normals(seed, 1, f, c, Z + f);
    }
  }
  double* U = NULL;
  if (nu > 0) {                                            /* element i*nu + j of the step's virtual uniform array */
    U = (double*)malloc(sizeof(double) * (size_t)nu * n);
    orc_fill_uniforms(seed, 1, ORC_STREAM_UNIFORM, 0, (uint64_t)nu * n, U);
  }
  double* lw = lnw_out;
  #pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t i = 0; i < n; ++i) {                       /* for i=1:num_samples  :25,41 */
    double lat[4 + ORC_OUTLIER_ZWORDS];
    if (family == ORC_REGRESSION) lw[i] = regression_sample(params, obs, proposal, pp, Z + (int64_t)nz * i, lat);
    else if (family == ORC_OUTLIER_REGRESSION) lw[i] = outlier_regression_sample(params, obs, Z + (int64_t)nz * i, U + (int64_t)nu * i, lat);
    else if (family == ORC_UNIFORM_NORMAL) lw[i] = uniform_normal_sample(params, obs, proposal, pp, U + i, lat);
    else lw[i] = normal_normal_sample(params, obs, proposal, pp, Z + (int64_t)nz * i, lat);
    for (int d = 0; d < D; ++d) lat_out[d * n + i] = lat[d];
  }
  if (!zrep && Z) free(Z);
  free(U);
  double log_total = orc_logsumexp(lw, n);                /* :29,48 */
  *lml_out = log_total - orc_log((double)n);              /* :30,49 */
  for (int64_t i = 0; i < n; ++i) lnw_out[i] = lw[i] - log_total;   /* :31,50 */
  return 0;
}
