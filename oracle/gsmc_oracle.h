/* gsmc_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, sequential, fp64 restatement of the reference's particle-filter and
 * importance-sampling code for the catalogue model families. Nothing in the
 * product (gen_b200/) links, imports or calls this; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may.
 *
 * PARITY STATUS: the reference is pure Julia and Julia is not available in the
 * build or GPU environment, so the oracle cannot be checked against outputs of
 * the reference itself. It is pinned against every golden value the reference's
 * own tests hold for this path (HMM forward-algorithm log-ML -4.87645083351704,
 * test/inference/particle_filter.jl:52-81,140-142; the hand-computed forward
 * example :29-48; Unfold weight identities test/modeling_library/unfold.jl:43-72,
 * 196-234; IS invariants test/inference/importance_sampling.jl:18-34) and against
 * closed forms (Kalman, conjugate regression). Individual random draws and
 * ancestor indices of the Julia implementation are NOT pinned (they come from
 * Julia's MersenneTwister + Distributions.jl 0.24.10 alias sampler, neither of
 * which is in /root/reference): "parity unpinned at the draw level".
 *
 * Function-by-function map (all paths relative to /root/reference):
 *   orc_logsumexp              src/inference/inference.jl:3-6
 *   orc_normalize_weights      src/inference/particle_filter.jl:8-12
 *   orc_effective_sample_size  src/inference/particle_filter.jl:3-6
 *   orc_pf_init                src/inference/particle_filter.jl:79-108
 *   orc_pf_step                src/inference/particle_filter.jl:139-180,
 *                              src/inference/trace_translators.jl:783-802,
 *                              src/modeling_library/unfold/update.jl:54-78
 *   orc_pf_maybe_resample      src/inference/particle_filter.jl:189-213
 *   orc_pf_log_ml_estimate     src/inference/particle_filter.jl:52-55
 *   orc_pf_sample_unweighted   src/inference/particle_filter.jl:62-70
 *   orc_importance_sampling    src/inference/importance.jl:20-52
 *   orc_logpdf_normal          src/modeling_library/distributions/normal.jl:56-60
 *   orc_logpdf_categorical     src/modeling_library/distributions/categorical.jl:10-12
 */
#ifndef GSMC_ORACLE_H
#define GSMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model families (same numbering as include/gen_b200.h) */
enum { ORC_HMM = 1, ORC_LGSSM = 2, ORC_SV = 3, ORC_BEARINGS = 4, ORC_REGRESSION = 5, ORC_NORMAL_NORMAL = 6,
       ORC_OUTLIER_REGRESSION = 7,   /* examples/regression/static_model.jl:3-23 (bernoulli outlier flags, Map of `datum`) */
       ORC_UNIFORM_NORMAL = 8 };     /* x ~ uniform(lo, hi); y ~ normal(x, sd): uniform_continuous.jl:12-23 on the path */
#define ORC_OUTLIER_ZWORDS 8          /* outlier flags packed 32 per latent word: at most 256 data points */
enum { ORC_PROPOSAL_DEFAULT = 0, ORC_PROPOSAL_CUSTOM = 1 };
enum { ORC_RESAMPLE_MULTINOMIAL = 0, ORC_RESAMPLE_RESIDUAL = 1 };
/* Philox stream ids (counter word 3) */
enum { ORC_STREAM_NORMAL = 0, ORC_STREAM_UNIFORM = 1, ORC_STREAM_RESAMPLE = 2, ORC_STREAM_SAMPLE = 3, ORC_STREAM_GAP = 4, ORC_STREAM_OBS = 5 };

typedef struct orc_pf orc_pf;

/* ---- primitives ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* element e of the virtual normal / uniform array of (seed, t, stream) */
void orc_fill_normals(uint64_t seed, uint32_t t, uint64_t first, uint64_t count, double* out);
void orc_fill_uniforms(uint64_t seed, uint32_t t, uint32_t stream, uint64_t first, uint64_t count, double* out);
/* grouped order statistics behind the sorted resampling draws (see gsmc_oracle.c): Gamma(shape) gap variate of a
 * group in fixed point (scale 2^20), the gaps of groups [first, first+count) of an event with m draws, and the Exp(1)
 * head gap below the first order statistic */
uint64_t orc_gap_variate(uint64_t seed, uint32_t rho, uint64_t group, uint32_t shape);
void orc_fill_gaps(uint64_t seed, uint32_t rho, uint64_t m, uint64_t first, uint64_t count, uint64_t* out);
uint64_t orc_gap_head(uint64_t seed, uint32_t rho, uint64_t m);
double orc_div_inv(double x, double c);
double orc_log_pos(double x);
double orc_exp_nonpos(double x);
float orc_nlog_u32f(uint32_t w);                       /* -ln((w + 1/2) 2^-32) in fp32: radius of the per-particle normals */
void orc_sincos_u32f(uint32_t a, float* s, float* c);  /* (sin, cos)(2 pi a 2^-32) in fp32 */
double orc_log_unit(double x);   /* gm_log_unit: log of a uniform in (0,1), Box-Muller radius */
int64_t orc_muldiv_mismatches(uint64_t seed, int64_t n);
int64_t orc_div_inv_mismatches(uint64_t seed, int64_t n);
double orc_exp(double x);
double orc_log(double x);
double orc_atan2(double y, double x);
void orc_sincospi(double t, double* s, double* c);
double orc_logpdf_normal(double x, double mu, double std);
double orc_logpdf_categorical(int64_t x, const double* probs, int64_t n);
double orc_logpdf_uniform(double x, double low, double high);
double orc_logpdf_bernoulli(int x, double p);
double orc_logsumexp(const double* arr, int64_t n);
double orc_logsumexp2(double x1, double x2);
double orc_effective_sample_size(const double* log_normalized_weights, int64_t n);

/* integer resampling arithmetic (oracle-defined; see DESIGN.md "Resampling arithmetic") */
int orc_weight_shift(uint64_t n_global);   /* k: weights are quantised to floor(w * 2^k) */
void orc_quantise_weights(const double* lw, int64_t n, uint64_t n_global, uint64_t* q_out, double* max_out);
/* iid-uniform ("replay") search: anc[j] = min{ i : C_i > floor(floor(u_j*2^53) * C_N / 2^53) } (0-based) */
void orc_search_iid(const uint64_t* cdf, int64_t n, const double* u, int64_t m, int64_t* anc);
/* grouped-order-statistics search of the m draws of event rho (0-based ancestors): the slot that opens a group takes
 * the group's lower bracket position; every other slot counts the 32-bit bracket keys below its Philox word */
uint64_t orc_bracket_scale(uint64_t tl, uint64_t th);              /* (clz64(TH - TL) << 32) | M, 0 when TH <= TL */
uint32_t orc_bracket_key(uint64_t c_minus_tl, uint64_t scale);    /* ((((C_p - TL) << lz) >> 32) * M) >> 31 */
void orc_search_sorted(const uint64_t* cdf, int64_t n, uint64_t seed, uint32_t rho, int64_t m, int64_t* anc);

/* ---- particle filter ---- */
orc_pf* orc_pf_create(int family, const double* params, int n_params, int64_t num_particles,
                      uint64_t seed, int keep_history, int num_threads);
void orc_pf_destroy(orc_pf* pf);
int orc_pf_state_dim(const orc_pf* pf);
/* init/step with obs == NULL (n_obs == 0): an UNOBSERVED step -- the observation choice is sampled, the weight is unchanged */
int orc_pf_sampled_observation(const orc_pf* pf, int64_t t, double* out);
int orc_pf_num_normals(const orc_pf* pf, int proposal, int is_init);
int orc_pf_num_uniforms(const orc_pf* pf, int proposal, int is_init);
/* initialize_particle_filter. obs: family-specific observation vector (may be NULL = no constraint).
 * z_replay/u_replay: optional exported draws (NULL -> Philox). returns 0 or <0 on error. */
int orc_pf_init(orc_pf* pf, const double* obs, int n_obs, int proposal, const double* prop_params, int n_prop,
                const double* z_replay, const double* u_replay);
int orc_pf_step(orc_pf* pf, const double* obs, int n_obs, int proposal, const double* prop_params, int n_prop,
                const double* z_replay, const double* u_replay);
/* maybe_resample!. u_replay: optional N iid uniforms (one per output slot). */
int orc_pf_maybe_resample(orc_pf* pf, double ess_threshold, int scheme, const double* u_replay,
                          int* did_resample, double* ess_out, double* log_total_out);
double orc_pf_log_ml_estimate(const orc_pf* pf);
const double* orc_pf_log_weights(const orc_pf* pf);
void orc_pf_set_log_weights(orc_pf* pf, const double* lw);
const int64_t* orc_pf_parents(const orc_pf* pf);          /* 0-based ancestors of the last resample */
const double* orc_pf_state(const orc_pf* pf);             /* current latent, column-major D x N */
/* full trajectory value: time index t (1-based), dim d, particle i -- needs keep_history */
int orc_pf_history(const orc_pf* pf, int64_t t, double* out /* D x N */);
int64_t orc_pf_num_steps(const orc_pf* pf);
int orc_pf_sample_unweighted(orc_pf* pf, int64_t num_samples, const double* u_replay, int64_t* idx_out);

/* ---- importance sampling ---- */
/* returns normalised log weights, latents (column-major D x n) and the log-ML estimate */
int orc_importance_sampling(int family, const double* params, int n_params, const double* obs, int n_obs,
                            int proposal, const double* prop_params, int n_prop,
                            int64_t num_samples, uint64_t seed, const double* z_replay,
                            double* latents_out, double* log_norm_weights_out, double* lml_out, int num_threads);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
