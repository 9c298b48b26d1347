// models.cuh -- the model catalogue as per-particle device functors.
//
// Each functor is the collapsed form of what the reference executes per particle per
// time step through the GFI: Unfold's process_new! (src/modeling_library/unfold/update.jl:54-78)
// calling generate() on a static-IR kernel (src/static_ir/generate.jl:24-43: constrained choice ->
// weight += logpdf; unconstrained choice -> random()), and for custom proposals the
// SimpleExtendingTraceTranslator (src/inference/trace_translators.jl:783-802: weight =
// model_weight - proposal_score). Arithmetic follows the reference's distribution code
// literally (normal.jl:56-60,96; categorical.jl:10-12) so that log weights agree bit for bit
// with the CPU oracle; constants that do not depend on the particle (std*std, log(2 pi var))
// are hoisted into ModelArgs::k by prepare() with the same IEEE operations.
#ifndef GSMC_MODELS_CUH
#define GSMC_MODELS_CUH

#include "gsmc_math.h"

#define GSMC_MAX_INLINE_PARAMS 32
#define GSMC_MAX_INLINE_OBS 16
#define GSMC_HMM_MAX_K 16

struct NormC { double two_var, half_log, inv_two_var; uint32_t span, pad_; };   // hoisted constants of one normal logpdf

struct ModelArgs {
  double p[GSMC_MAX_INLINE_PARAMS];   // model parameters (prefix; all of them live at p_dev too)
  double obs[GSMC_MAX_INLINE_OBS];    // this step's observations (prefix; obs_dev when longer)
  double pp[8];                       // proposal parameters
  double k[16];                       // per-launch derived constants (prepare())
  NormC nc[6];                        // per-launch hoisted normal-logpdf constants (prepare())
  const double* p_dev;
  const double* obs_dev;
  int n_p, n_obs;
  int unobserved;                     // this step has no constraint: the weight increment is 0 (the observation choice is sampled)
};

// normal.jl:56-60 with var = std*std hoisted:  -(diff*diff)/(2.0*var) - 0.5*log(2.0*pi*var)
GM_HD NormC make_normc(double std) {
  NormC c;
  const double var = std * std;
  c.two_var = 2.0 * var;
  c.half_log = 0.5 * gm_log(2.0 * GM_PI * var);
  c.inv_two_var = gm_safe_recip(c.two_var);
  c.span = gm_div_span(c.inv_two_var); c.pad_ = 0;
  return c;
}
// the division by the launch-invariant 2*var is done with gm_div_inv: same bits as `/`, 3 instructions
GM_HD double logpdf_normal_c(double x, double mu, NormC c) {
  const double diff = x - mu;
  return gm_div_inv_s(-(diff * diff), c.two_var, c.inv_two_var, c.span) - c.half_log;
}
// per-particle std (nothing to hoist): the literal formula
GM_HD double logpdf_normal(double x, double mu, double std) {
  const double var = std * std;
  const double diff = x - mu;
  return -(diff * diff) / (2.0 * var) - 0.5 * gm_log(2.0 * GM_PI * var);
}
GM_HD double random_normal(double mu, double std, double z) { return mu + std * z; }   // normal.jl:96
// bernoulli.jl:10-12,19: logpdf = x ? log(prob) : log(1. - prob); random = rand() < prob
GM_HD double logpdf_bernoulli(bool x, double prob) { return x ? gm_log(prob) : gm_log(1. - prob); }
GM_HD bool random_bernoulli(double prob, double u) { return u < prob; }
// uniform_continuous.jl:12-14,21-23: logpdf = (x >= low && x <= high) ? -log(high-low) : -Inf; random = rand() * (high - low) + low
GM_HD double logpdf_uniform(double x, double low, double high) { return (x >= low && x <= high) ? -gm_log(high - low) : -gm_inf(); }
GM_HD double random_uniform(double low, double high, double u) { return u * (high - low) + low; }

// What a model with a run-time number of uniform draws per particle needs to draw them itself: element
// global_index * n + j of the step's virtual uniform array (same layout as the fixed-count path), or the replayed values.
struct DrawCtx {
  uint64_t seed;
  uint64_t global_index;
  uint32_t t;
  const double* urep;       // this particle's replayed uniforms, or NULL
};

// ---------------------------------------------------------------------------------------------
// LGSSM: x_init ~ normal(m0,s0); x ~ normal(x_prev*a + b, q); y ~ normal(c*x, r)
// p = [m0, s0, a, b, q, c, r]; obs = [y]
// k = [mean_sd, obs.two_var, obs.half_log, lat.two_var, lat.half_log, prop_var, prop_sd, q.two_var, q.half_log,
//      c*c/(r*r) numerator terms ...]
// ---------------------------------------------------------------------------------------------
struct LgssmModel {
  static constexpr int D = 1;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool, int) { return 1; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs& a, bool init, int prop) {
    const double* p = a.p;
    const double sd = init ? p[1] : p[4], c = p[5], r = p[6];
    a.k[0] = sd; a.nc[0] = make_normc(r); a.nc[1] = make_normc(sd);
    if (prop == 1) {
      const double prec = 1.0 / (sd * sd) + (c * c) / (r * r);
      const double var = 1.0 / prec;
      const double sdq = sqrt(var);
      a.nc[2] = make_normc(sdq);
      a.k[5] = var; a.k[6] = sdq;
      a.k[9] = sd * sd; a.k[10] = r * r;
    }
  }
  // the observation choice of an unobserved step: y ~ normal(c*x, r)
  static constexpr bool OBS_DRAW_UNIFORM = false;
  __device__ __forceinline__ static double sample_obs(const ModelArgs& a, const double* lat, double z) { return random_normal(a.p[5] * lat[0], a.p[6], z); }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double* prev,
                                                    const double* z, const double*, double* out) {
    const double* p = a.p;
    const double mean = INIT ? p[0] : prev[0] * p[2] + p[3];
    const double y = a.obs[0], c = p[5];
    const NormC on = a.nc[0];
    if (PROP == 0) {
      const double x = random_normal(mean, a.k[0], z[0]);
      double w = 0.0;
      w += logpdf_normal_c(y, c * x, on);
      out[0] = x;
      return w;
    } else {
      const double mu = a.k[5] * (mean / a.k[9] + (c * y) / a.k[10]);
      const double x = random_normal(mu, a.k[6], z[0]);
      const NormC qn = a.nc[2], ln = a.nc[1];
      const double q_score = logpdf_normal_c(x, mu, qn);
      double mw = 0.0;
      mw += logpdf_normal_c(x, mean, ln);
      mw += logpdf_normal_c(y, c * x, on);
      out[0] = x;
      return mw - q_score;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Stochastic volatility: h_init ~ normal(mu, sigma/sqrt(1-phi*phi)); h ~ normal(mu + phi*(h_prev-mu), sigma);
// y ~ normal(0, exp(h/2)).  p = [mu, phi, sigma]; obs = [y]
// ---------------------------------------------------------------------------------------------
struct SvModel {
  static constexpr int D = 1;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool, int) { return 1; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0; }
  static void prepare(ModelArgs& a, bool init, int) {
    const double phi = a.p[1], sigma = a.p[2];
    a.k[0] = init ? sigma / sqrt(1.0 - phi * phi) : sigma;
  }
  static constexpr bool OBS_DRAW_UNIFORM = false;       // y ~ normal(0, exp(h/2))
  __device__ __forceinline__ static double sample_obs(const ModelArgs&, const double* lat, double z) { return random_normal(0.0, gm_exp(lat[0] / 2.0), z); }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double* prev,
                                                    const double* z, const double*, double* out) {
    const double mu = a.p[0], phi = a.p[1];
    const double mean = INIT ? mu : mu + phi * (prev[0] - mu);
    const double h = random_normal(mean, a.k[0], z[0]);
    double w = 0.0;
    w += logpdf_normal(a.obs[0], 0.0, gm_exp(h / 2.0));
    out[0] = h;
    return w;
  }
};

// ---------------------------------------------------------------------------------------------
// Bearings-only tracking, state (x, vx, y, vy).  p = [m[4], sd[4], sigma_w, sigma_theta]; obs=[bearing]
// ---------------------------------------------------------------------------------------------
struct BearingsModel {
  static constexpr int D = 4;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool init, int) { return init ? 4 : 2; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs& a, bool, int) {
    const double sw = a.p[8], st = a.p[9];
    a.nc[0] = make_normc(st); a.nc[1] = make_normc(sw);
    a.k[4] = sw * sw; a.k[5] = st * st;
  }
  static constexpr bool OBS_DRAW_UNIFORM = false;       // bearing ~ normal(atan(y, x), sigma_theta)
  __device__ __forceinline__ static double sample_obs(const ModelArgs& a, const double* lat, double z) { return random_normal(gm_atan2(lat[2], lat[0]), a.p[9], z); }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double* sp,
                                                    const double* z, const double*, double* out) {
    const double* p = a.p;
    const NormC tn = a.nc[0];
    const double obs = a.obs[0];
    if (INIT) {
      const double x = random_normal(p[0], p[4], z[0]);
      const double vx = random_normal(p[1], p[5], z[1]);
      const double y = random_normal(p[2], p[6], z[2]);
      const double vy = random_normal(p[3], p[7], z[3]);
      double w = 0.0;
      w += logpdf_normal_c(obs, gm_atan2(y, x), tn);
      out[0] = x; out[1] = vx; out[2] = y; out[3] = vy;
      return w;
    }
    const double sw = p[8];
    double wx, wy, q_score = 0.0, mw = 0.0;
    if (PROP == 0) {
      wx = random_normal(0.0, sw, z[0]);
      wy = random_normal(0.0, sw, z[1]);
    } else {
      const double sw2 = a.k[4], st2 = a.k[5];
      const double xb = sp[0] + sp[1], yb = sp[2] + sp[3];
      const double rho2 = xb * xb + yb * yb;
      const double nu = obs - gm_atan2(yb, xb);
      const double hx = -yb / rho2, hy = xb / rho2;
      const double S = 0.25 * sw2 * (hx * hx + hy * hy) + st2;
      const double kx = 0.5 * sw2 * hx / S, ky = 0.5 * sw2 * hy / S;
      const double mx = kx * nu, my = ky * nu;
      const double sx = sqrt(sw2 * (1.0 - 0.5 * kx * hx)), sy = sqrt(sw2 * (1.0 - 0.5 * ky * hy));
      wx = random_normal(mx, sx, z[0]);
      wy = random_normal(my, sy, z[1]);
      q_score += logpdf_normal(wx, mx, sx);
      q_score += logpdf_normal(wy, my, sy);
      const NormC wn = a.nc[1];
      mw += logpdf_normal_c(wx, 0.0, wn);
      mw += logpdf_normal_c(wy, 0.0, wn);
    }
    const double x = sp[0] + sp[1] + 0.5 * wx;
    const double vx = sp[1] + wx;
    const double y = sp[2] + sp[3] + 0.5 * wy;
    const double vy = sp[3] + wy;
    mw += logpdf_normal_c(obs, gm_atan2(y, x), tn);
    out[0] = x; out[1] = vx; out[2] = y; out[3] = vy;
    return mw - q_score;
  }
};

// ---------------------------------------------------------------------------------------------
// HMM (test/inference/particle_filter.jl:52-78; proposals :104-127).
// p = [K, V, prior[K], trans[K][K] (row z_prev), emis[K][V] (row z)]; obs = [x] (1-based); latent stored 1-based.
// Shared-memory tables built per block by prologue():
//   cum[zp][k]  running sums of the sampling distribution (prior / transition row, or the
//               normalised locally-optimal proposal), exactly the `cp` sequence of the linear scan
//   wt[zp][z]   the weight increment for choosing z from zp
// ---------------------------------------------------------------------------------------------
struct HmmModel {
  static constexpr int D = 1;
  static constexpr int SMEM_DOUBLES = 2 * GSMC_HMM_MAX_K * GSMC_HMM_MAX_K;
  __host__ __device__ static constexpr int nz(bool, int) { return 0; }
  __host__ __device__ static constexpr int nu(bool, int) { return 1; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs&, bool, int) {}
  static constexpr bool OBS_DRAW_UNIFORM = true;        // x ~ categorical(emission_dists[:, z]): linear scan with one uniform
  __device__ __forceinline__ static double sample_obs(const ModelArgs& a, const double* lat, double u) {
    const double* p = a.p_dev;
    const int K = (int)p[0], V = (int)p[1];
    const double* e = p + 2 + K + K * K + ((int)lat[0] - 1) * V;
    double cp = e[0];
    int i = 1;
    while (cp <= u && i < V) { cp += e[i]; i += 1; }
    return (double)i;
  }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs& a, double* sm) {
    const double* p = a.p_dev;
    const int K = (int)p[0], V = (int)p[1];
    const double* prior = p + 2;
    const double* trans = prior + K;
    const double* emis = trans + K * K;
    const int x = (int)a.obs[0];
    double* cum = sm;
    double* wt = sm + GSMC_HMM_MAX_K * GSMC_HMM_MAX_K;
    const int rows = INIT ? 1 : K;
    for (int zp = threadIdx.x; zp < rows; zp += blockDim.x) {
      const double* pz = INIT ? prior : trans + zp * K;
      double dist[GSMC_HMM_MAX_K];
      if (PROP == 0) {
        for (int k = 0; k < K; ++k) dist[k] = pz[k];
      } else {
        double s = 0.0;
        for (int k = 0; k < K; ++k) dist[k] = pz[k] * emis[k * V + (x - 1)];
        for (int k = 0; k < K; ++k) s += dist[k];
        for (int k = 0; k < K; ++k) dist[k] = dist[k] / s;
      }
      double cp = dist[0];
      cum[zp * GSMC_HMM_MAX_K] = cp;
      for (int k = 1; k < K; ++k) { cp += dist[k]; cum[zp * GSMC_HMM_MAX_K + k] = cp; }
      for (int z = 0; z < K; ++z) {
        const bool x_ok = (x > 0 && x <= V);
        const double le = x_ok ? gm_log(emis[z * V + (x - 1)]) : -gm_inf();
        double w;
        if (PROP == 0) {
          w = 0.0;
          w += le;
        } else {
          const double q_score = gm_log(dist[z]);
          double mw = 0.0;
          mw += gm_log(pz[z]);
          mw += le;
          w = mw - q_score;
        }
        wt[zp * GSMC_HMM_MAX_K + z] = w;
      }
    }
  }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double* sm, const double* prev,
                                                    const double*, const double* u, double* out) {
    const int K = (int)a.p[0];
    const int zp = INIT ? 0 : (int)prev[0] - 1;
    const double* cum = sm + zp * GSMC_HMM_MAX_K;
    // Distributions.jl linear scan: cp = p[1]; i = 1; while cp <= draw && i < n: cp += p[i += 1]
    int i = 1;
    while (cum[i - 1] <= u[0] && i < K) ++i;
    out[0] = (double)i;
    return sm[GSMC_HMM_MAX_K * GSMC_HMM_MAX_K + zp * GSMC_HMM_MAX_K + (i - 1)];
  }
};

// ---------------------------------------------------------------------------------------------
// Importance-sampling families (run through the init kernel: one "time step", no resampling).
// regression: p = [n, sd_s, sd_i, sd_n, xs[n]]; obs = ys[n]; pp = [mu_s, sd_s', mu_i, sd_i']
// ---------------------------------------------------------------------------------------------
struct RegressionModel {
  static constexpr int D = 2;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool, int) { return 2; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs& a, bool, int prop) {
    a.nc[0] = make_normc(a.p[3]);
    if (prop == 1) {
      a.nc[1] = make_normc(a.pp[1]); a.nc[2] = make_normc(a.pp[3]); a.nc[3] = make_normc(a.p[1]); a.nc[4] = make_normc(a.p[2]);
    }
  }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double*,
                                                    const double* z, const double*, double* out) {
    const int n = (int)a.p[0];
    const double* xs = a.p_dev + 4;
    const double* ys = a.obs_dev;
    double slope, intercept, pw = 0.0, mw = 0.0;
    if (PROP == 0) {
      slope = random_normal(0.0, a.p[1], z[0]);
      intercept = random_normal(0.0, a.p[2], z[1]);
    } else {
      slope = random_normal(a.pp[0], a.pp[1], z[0]);
      intercept = random_normal(a.pp[2], a.pp[3], z[1]);
      const NormC a0 = a.nc[1], a1 = a.nc[2], b0 = a.nc[3], b1 = a.nc[4];
      pw += logpdf_normal_c(slope, a.pp[0], a0);
      pw += logpdf_normal_c(intercept, a.pp[2], a1);
      mw += logpdf_normal_c(slope, 0.0, b0);
      mw += logpdf_normal_c(intercept, 0.0, b1);
    }
    const NormC nn = a.nc[0];
    for (int i = 0; i < n; ++i) mw += logpdf_normal_c(__ldg(ys + i), slope * __ldg(xs + i) + intercept, nn);
    out[0] = slope; out[1] = intercept;
    return mw - pw;
  }
};

// ---------------------------------------------------------------------------------------------
// Outlier regression, examples/regression/static_model.jl:3-23 (a static model over a Map of the static `datum`):
//   log_inlier_std, log_outlier_std, slope, intercept ~ normal(0, sd); per datum z ~ bernoulli(prob),
//   std = ifelse(z, inlier_std, outlier_std) [literally :6], y ~ normal(x * slope + intercept, std) (constrained).
// p = [n, prob, sd, xs[n]]; obs = ys[n]. Latent columns: the four reals, then the flags packed 32 per column as exact
// integers (GSMC_OUTLIER_ZWORDS columns: n <= 256). The n uniforms of a particle come through DrawCtx.
// ---------------------------------------------------------------------------------------------
#define GSMC_OUTLIER_ZWORDS 8
struct OutlierRegressionModel {
  static constexpr int D = 4 + GSMC_OUTLIER_ZWORDS;
  static constexpr int SMEM_DOUBLES = 0;
  static constexpr bool CTX_UNIFORMS = true;
  __host__ __device__ static constexpr int nz(bool, int) { return 4; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0; }
  static void prepare(ModelArgs&, bool, int) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle_ctx(const ModelArgs& a, const double*, const double*, const double* z,
                                                        const DrawCtx& c, double* out) {
    const int n = (int)a.p[0];
    const double prob = a.p[1], sd = a.p[2];
    const double* xs = a.p_dev + 3;
    const double* ys = a.obs_dev;
    const double inlier_log_std = random_normal(0.0, sd, z[0]);
    const double outlier_log_std = random_normal(0.0, sd, z[1]);
    const double inlier_std = gm_exp(inlier_log_std), outlier_std = gm_exp(outlier_log_std);
    const double slope = random_normal(0.0, sd, z[2]);
    const double intercept = random_normal(0.0, sd, z[3]);
    double w = 0.0;
    uint32_t words[GSMC_OUTLIER_ZWORDS];
#pragma unroll
    for (int k = 0; k < GSMC_OUTLIER_ZWORDS; ++k) words[k] = 0;
    const uint64_t e0 = c.global_index * (uint64_t)n;
    double ua = 0.0, ub = 0.0;
    uint64_t have = ~(uint64_t)0;                        // Philox call whose two uniforms are in (ua, ub)
    for (int i = 0; i < n; ++i) {
      double u;
      if (c.urep) u = c.urep[i];
      else {
        const uint64_t e = e0 + (uint64_t)i;
        if ((e >> 1) != have) { have = e >> 1; uniform_pair(c.seed, have, c.t, GSMC_STREAM_UNIFORM, &ua, &ub); }
        u = (e & 1) ? ub : ua;
      }
      const bool is_outlier = random_bernoulli(prob, u);
      const double std = is_outlier ? inlier_std : outlier_std;
      w += logpdf_normal(__ldg(ys + i), __ldg(xs + i) * slope + intercept, std);
      // the flag goes into word i >> 5 without indexing the register array dynamically
#pragma unroll
      for (int k = 0; k < GSMC_OUTLIER_ZWORDS; ++k) if ((i >> 5) == k && is_outlier) words[k] |= 1u << (i & 31);
    }
    out[0] = inlier_log_std; out[1] = outlier_log_std; out[2] = slope; out[3] = intercept;
#pragma unroll
    for (int k = 0; k < GSMC_OUTLIER_ZWORDS; ++k) out[4 + k] = (double)words[k];
    return w;
  }
};

// uniform-normal: x ~ uniform(lo, hi); y ~ normal(x, sd_y). p = [lo, hi, sd_y]; obs = [y]; pp = [lo_q, hi_q]
struct UniformNormalModel {
  static constexpr int D = 1;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool, int) { return 0; }
  __host__ __device__ static constexpr int nu(bool, int) { return 1; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs& a, bool, int) { a.nc[0] = make_normc(a.p[2]); }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double*,
                                                    const double*, const double* u, double* out) {
    double x, pw = 0.0, mw = 0.0;
    if (PROP == 0) {
      x = random_uniform(a.p[0], a.p[1], u[0]);
    } else {
      x = random_uniform(a.pp[0], a.pp[1], u[0]);
      pw += logpdf_uniform(x, a.pp[0], a.pp[1]);
      mw += logpdf_uniform(x, a.p[0], a.p[1]);
    }
    const NormC yn = a.nc[0];
    mw += logpdf_normal_c(a.obs[0], x, yn);
    out[0] = x;
    return mw - pw;
  }
};

// normal-normal: x ~ normal(mu0, sd0); y ~ normal(x, sd_y). p = [mu0, sd0, sd_y]; obs = [y]; pp = [mu_q, sd_q]
struct NormalNormalModel {
  static constexpr int D = 1;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool, int) { return 1; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static bool has_proposal(int prop) { return prop == 0 || prop == 1; }
  static void prepare(ModelArgs& a, bool, int prop) {
    a.nc[0] = make_normc(a.p[2]);
    if (prop == 1) { a.nc[1] = make_normc(a.pp[1]); a.nc[2] = make_normc(a.p[1]); }
  }
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double*,
                                                    const double* z, const double*, double* out) {
    double x, pw = 0.0, mw = 0.0;
    if (PROP == 0) {
      x = random_normal(a.p[0], a.p[1], z[0]);
    } else {
      x = random_normal(a.pp[0], a.pp[1], z[0]);
      const NormC qn = a.nc[1], xn = a.nc[2];
      pw += logpdf_normal_c(x, a.pp[0], qn);
      mw += logpdf_normal_c(x, a.p[0], xn);
    }
    const NormC yn = a.nc[0];
    mw += logpdf_normal_c(a.obs[0], x, yn);
    out[0] = x;
    return mw - pw;
  }
};

#endif
