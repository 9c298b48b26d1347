/* gen_b200.h -- C ABI of libgensmc.so: B200-native sequential Monte Carlo and
 * importance sampling behind Gen.jl's src/inference API.
 *
 * The reference (Gen.jl v0.4.1, pure Julia) has no FFI for this path; its
 * extension mechanism is multiple dispatch on the exported generic functions
 * (src/inference/particle_filter.jl:215-216, src/inference/importance.jl:110).
 * Each entry point below is what a Julia method of that generic function binds
 * with `ccall` (julia/GenB200.jl); the comment above each names
 * the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions: every call returns 0 or a negative GSMC_E_* code and never
 * throws/aborts; gsmc_last_error() gives the message. Host pointers are borrowed
 * for the duration of the call; the library owns all device memory behind the
 * opaque handle. A handle is not thread-safe (the reference is single-threaded).
 * Indices are 0-based here; the Julia wrapper adds 1. Particle state is stored
 * as structure-of-arrays device columns; host copies are column-major [D][n].
 */
#ifndef GEN_B200_H
#define GEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GSMC_API
#else
#define GSMC_API __attribute__((visibility("default")))
#endif

typedef struct gsmc_filter* gsmc_handle;

/* model catalogue: static-IR kernels + Unfold, see DESIGN.md "Model catalogue" */
enum {
  GSMC_MODEL_HMM = 1,            /* test/inference/particle_filter.jl:52-78 */
  GSMC_MODEL_LGSSM = 2,          /* kernel form of test/modeling_library/unfold.jl:5-8 + observation */
  GSMC_MODEL_SV = 3,             /* stochastic volatility; SSM pattern of examples/pmmh/model.jl:40-50 */
  GSMC_MODEL_BEARINGS = 4,       /* bearings-only tracking; pattern of examples/planning/filtering.jl:79-91 */
  GSMC_MODEL_REGRESSION = 5,     /* examples/regression/quickstart.jl:3-9 (importance sampling) */
  GSMC_MODEL_NORMAL_NORMAL = 6,  /* test/inference/importance_sampling.jl:3-12 (importance sampling) */
  GSMC_MODEL_OUTLIER_REGRESSION = 7, /* examples/regression/static_model.jl:3-23: bernoulli outlier flags, Map of a static kernel (importance sampling) */
  GSMC_MODEL_UNIFORM_NORMAL = 8  /* x ~ uniform(low, high); y ~ normal(x, sd): uniform_continuous.jl:12-23 on the device (importance sampling) */
};
/* models generated from a static-IR description (gsmc_register_model_plugin) get ids GSMC_MODEL_PLUGIN_BASE + k */
#define GSMC_MODEL_PLUGIN_BASE 1000
enum { GSMC_PROPOSAL_DEFAULT = 0, GSMC_PROPOSAL_CUSTOM = 1 };
enum { GSMC_RESAMPLE_MULTINOMIAL = 0, GSMC_RESAMPLE_RESIDUAL = 1 };
enum { GSMC_F64 = 0, GSMC_F32 = 1 };   /* storage type of state columns and log weights; arithmetic is f64 */

enum {
  GSMC_OK = 0,
  GSMC_E_BADARG = -1,        /* invalid argument / call order (Julia: error(...)) */
  GSMC_E_CUDA = -2,          /* CUDA runtime error */
  GSMC_E_NCCL = -3,          /* NCCL error / NCCL not loadable */
  GSMC_E_DEGENERATE = -4,    /* total weight zero or not finite (reference: Categorical constructor throws) */
  GSMC_E_UNSUPPORTED = -5,   /* model/proposal combination not in the catalogue */
  GSMC_E_NOMEM = -6,
  GSMC_E_PEER = -7           /* sharded filter: a peer GPU did not answer a scalar exchange in time (lost / hung rank) */
};

typedef struct gsmc_config {
  uint32_t struct_size;      /* = sizeof(gsmc_config) */
  int32_t model_id;          /* GSMC_MODEL_* */
  int32_t dtype;             /* GSMC_F64 | GSMC_F32 */
  int32_t resample_scheme;   /* GSMC_RESAMPLE_* */
  uint64_t num_particles;    /* global particle count N (all ranks together) */
  uint64_t seed;             /* Philox4x32-10 key. Every draw is a function of (seed, global particle index, time step): the
                              * result does not depend on the GPU count. Standard normals have the resolution of their two
                              * 32-bit words (Box-Muller evaluated in fp32, widened to fp64); all model arithmetic is fp64. */
  int32_t device;            /* CUDA device ordinal, -1 = current device */
  int32_t keep_history;      /* 1: keep every step's state + ancestor columns (get_traces semantics) */
  int64_t history_capacity;  /* number of time steps to preallocate when keep_history (0 = 128) */
  void* stream;              /* cudaStream_t to run on; NULL = the library creates its own */
} gsmc_config;

typedef struct gsmc_stats {
  double last_ess;           /* of the last maybe_resample / log_ml_estimate */
  double last_log_total;     /* logsumexp(log_weights) */
  double log_ml_est;         /* accumulated state.log_ml_est */
  int64_t num_steps;         /* time steps in the traces (1 after init) */
  int64_t num_resamples;     /* resampling events so far */
  int64_t kernel_launches;   /* kernels launched by this handle so far */
  /* per-kernel-class device time in ms and launch counts, filled while profiling is on.
   * propagate = init/step launches that read the previous state directly; propagate_gather = step
   * launches that read it through the ancestor column of a pending resample (per-call API only;
   * gsmc_run_steps cannot tell them apart without a host round trip and books all as propagate). */
  double ms_propagate, ms_propagate_gather, ms_finalize, ms_scan, ms_spacings, ms_search, ms_other;
  int64_t n_propagate, n_propagate_gather, n_finalize, n_scan, n_spacings, n_search, n_other;
  int64_t graph_replays;     /* gsmc_run_steps calls served by the captured CUDA graph of a repeated run shape */
} gsmc_stats;

GSMC_API const char* gsmc_version(void);
/* message of the last failed call on this thread (h may be NULL) */
GSMC_API const char* gsmc_last_error(gsmc_handle h);

/* Builds the device-side ParticleFilterState{U} (src/inference/particle_filter.jl:18-24).
 * params: model parameter vector in the catalogue layout. */
GSMC_API int gsmc_create(const gsmc_config* cfg, const double* params, size_t n_params, gsmc_handle* out);
GSMC_API void gsmc_destroy(gsmc_handle h);
/* Back to the state right after gsmc_create (+ attach): no time steps, log_ml_est = 0, the resample
 * event counter back to 0 so that a repeated run redraws the same Philox stream. Buffers are kept. */
GSMC_API int gsmc_reset(gsmc_handle h);

/* Multi-GPU (one process per GPU, NCCL over NVLink): rank 0 makes an id and hands it to every rank
 * (the host does that, e.g. through torch.distributed); each rank creates one communicator and
 * attaches its filters to it before gsmc_init. After attach the handle owns particles
 * [rank*N/R, (rank+1)*N/R). No reference equivalent (the reference is single-process);
 * SURVEY.md section 8(e). NCCL creates the communicator and carries the one-time exchange of the
 * CUDA-IPC handles; the per-step scalars (logsumexp triples, integer weight and spacing totals) travel
 * as peer-memory stores fused into the finalize/scan kernels (set GSMC_NCCL_SCALARS=1 to use
 * ncclAllGather for them instead), and remote ancestors are read with peer-memory loads. */
typedef struct gsmc_comm_s* gsmc_comm;
GSMC_API int gsmc_comm_unique_id(void* id_out, size_t nbytes /* >= 128 */);
GSMC_API int gsmc_comm_create(const void* unique_id, size_t nbytes, int rank, int nranks, int device, gsmc_comm* out);
GSMC_API void gsmc_comm_destroy(gsmc_comm c);
GSMC_API int gsmc_comm_attach(gsmc_handle h, gsmc_comm c);

/* Shard emulation: the R ranks of ONE sharded filter as R handles in one process on one device, sharing one stream.
 * The data path is the multi-GPU code (rank offsets of the integer CDF and of the group gaps, CDF windows and ancestor
 * gathers that cross shard boundaries, "peer" pointers); only the scalar exchanges differ: a rank reads its peers'
 * scalars directly, and the gsmc_group_* calls enqueue every rank's producer kernels before any rank's consumers.
 * Used by the 1-GPU parity tests of the sharded path; no reference equivalent.
 * Attach every rank before gsmc_group_init; per-handle getters (log weights, state, ancestors, log-ML estimate,
 * trajectories) work on the members as on any sharded handle. gsmc_group_destroy destroys the members too. */
typedef struct gsmc_group_s* gsmc_group;
GSMC_API int gsmc_group_create(int nranks, int device, gsmc_group* out);
GSMC_API void gsmc_group_destroy(gsmc_group g);
GSMC_API int gsmc_group_attach(gsmc_group g, int rank, gsmc_handle h);
GSMC_API int gsmc_group_init(gsmc_group g, const double* obs, size_t n_obs,
                             int proposal_id, const double* proposal_params, size_t n_proposal_params);
GSMC_API int gsmc_group_step(gsmc_group g, const double* obs, size_t n_obs,
                             int proposal_id, const double* proposal_params, size_t n_proposal_params);
GSMC_API int gsmc_group_maybe_resample(gsmc_group g, double ess_threshold, int* did_resample, double* ess_out);
GSMC_API int gsmc_group_sample_unweighted(gsmc_group g, uint64_t num_samples, int64_t* idx_out);

/* Replay mode: the draws the next init/step/maybe_resample/sample_unweighted call consumes,
 * instead of Philox (this rank's slice). normals: n_local*n_norm values ordered
 * [particle][draw]; uniforms: n_local*n_unif for init/step, one per output slot for
 * maybe_resample, one per sample for sample_unweighted. Cleared by the call that uses them. */
GSMC_API int gsmc_set_replay(gsmc_handle h, const double* normals, size_t n_normals,
                             const double* uniforms, size_t n_uniforms);

/* initialize_particle_filter, both methods (src/inference/particle_filter.jl:79-91, 99-108).
 * obs: the observation choices of time step 1 in catalogue order. */
GSMC_API int gsmc_init(gsmc_handle h, const double* obs, size_t n_obs,
                       int proposal_id, const double* proposal_params, size_t n_proposal_params);

/* particle_filter_step!, both methods (src/inference/particle_filter.jl:139-154, 162-180;
 * SimpleExtendingTraceTranslator, src/inference/trace_translators.jl:783-802). Extends every
 * trace by one time step with the given observations. */
GSMC_API int gsmc_step(gsmc_handle h, const double* obs, size_t n_obs,
                       int proposal_id, const double* proposal_params, size_t n_proposal_params);

/* maybe_resample! (src/inference/particle_filter.jl:189-213). Resamples iff ess < ess_threshold. */
GSMC_API int gsmc_maybe_resample(gsmc_handle h, double ess_threshold, int* did_resample, double* ess_out);

/* log_ml_estimate (src/inference/particle_filter.jl:52-55) */
GSMC_API int gsmc_log_ml_estimate(gsmc_handle h, double* out);

/* get_log_weights (src/inference/particle_filter.jl:43-45): unnormalised log weights of this
 * rank's particles, converted to f64. n must be the local particle count. */
GSMC_API int gsmc_get_log_weights(gsmc_handle h, double* host_dst, size_t n);
/* device pointer to the same column (storage dtype), valid until the next call on h */
GSMC_API int gsmc_get_log_weights_device(gsmc_handle h, void** dev_ptr);

/* get_traces (src/inference/particle_filter.jl:31-34), structure-of-arrays view:
 * latent of time step t (1-based; 0 = current) for this rank's particles in their current
 * order, column-major [D][n_local], f64. t < current needs keep_history (walks ancestors). */
GSMC_API int gsmc_get_state(gsmc_handle h, int64_t t, double* host_dst, size_t n_values);
/* An UNOBSERVED step: gsmc_init / gsmc_step with obs == NULL and n_obs == 0 (default proposal only). As the reference
 * does for an unconstrained choice (src/static_ir/generate.jl:36-42, src/modeling_library/unfold/update.jl:54-78), the
 * latent is sampled as usual, the observation choice is sampled from the model given the new latent, and the weight is
 * unchanged. gsmc_get_observation returns the sampled observation choice of unobserved step t (1-based; 0 = current)
 * for this handle's particles in their current order (walks ancestors like gsmc_get_state); it fails for observed steps. */
GSMC_API int gsmc_get_observation(gsmc_handle h, int64_t t, double* host_dst, size_t n);
/* full trajectories of selected particles: out[s][t][d], t = 1..num_steps. idx holds GLOBAL particle indices
 * (= local indices on an unsharded filter). Rows owned by other ranks of a sharded filter are read through the
 * NVLink peer mappings; a call that asks for such rows is collective (every rank passes the same indices). */
GSMC_API int gsmc_get_trajectories(gsmc_handle h, const int64_t* idx, size_t n_idx, double* out, size_t n_values);
/* state.parents of the last resample (particle_filter.jl:200), global 0-based indices */
GSMC_API int gsmc_get_ancestors(gsmc_handle h, int64_t* host_dst, size_t n);

/* sample_unweighted_traces (src/inference/particle_filter.jl:62-70): num_samples categorical
 * draws from the normalised weights; returns GLOBAL particle indices (every rank of a sharded filter makes
 * the same call -- it is collective -- and receives the same indices). */
GSMC_API int gsmc_sample_unweighted(gsmc_handle h, uint64_t num_samples, int64_t* idx_out);

/* importance_sampling, both methods (src/inference/importance.jl:20-52). Returns a handle whose
 * log weights are the NORMALISED log weights (importance.jl:31,50) and whose state is the
 * sampled latents; *lml_out = log_total_weight - log(num_samples). */
GSMC_API int gsmc_importance_sampling(const gsmc_config* cfg, const double* params, size_t n_params,
                                      const double* obs, size_t n_obs,
                                      int proposal_id, const double* proposal_params, size_t n_proposal_params,
                                      double* lml_out, gsmc_handle* out);

/* The canonical driver loop of test/inference/particle_filter.jl:130-137 enqueued without a host
 * round trip per step: for t in 1..T-1 { maybe_resample!(ess_frac*N); particle_filter_step!(obs[t]) }
 * starting from an initialised filter. obs is [T_steps][n_obs]. Results via gsmc_log_ml_estimate /
 * gsmc_get_stats. With residual resampling (ten kernels per event), a run shape that repeats on a handle (same first
 * step, step count, observations, proposal and threshold, e.g. reset + init + run_steps in a loop) is captured into one
 * CUDA graph on its second occurrence and replayed afterwards; the resampling kernels of every step sit behind a
 * conditional node that the deciding kernel sets on the device, so steps that do not resample launch nothing
 * (GSMC_GRAPH=1: also for multinomial resampling, where it measured slightly slower; GSMC_NO_GRAPH=1: never). */
GSMC_API int gsmc_run_steps(gsmc_handle h, const double* obs, size_t n_steps, size_t n_obs,
                            int proposal_id, const double* proposal_params, size_t n_proposal_params,
                            double ess_threshold);

/* Checkpoint / resume (the reference's analogue is saving Julia objects with JLD,
 * examples/planning/filtering.jl:822-829). gsmc_save writes the filter behind h -- scalars, log weights, the state and
 * ancestor columns that exist -- to one file (one per rank of a sharded filter). gsmc_restore loads it into a handle
 * created with the same configuration and parameters (and attached to the same sharding); the run then continues
 * bit-identically to one that was never interrupted. */
GSMC_API int gsmc_save(gsmc_handle h, const char* path);
GSMC_API int gsmc_restore(gsmc_handle h, const char* path);

/* Beyond the catalogue: a state-space kernel written in a static IR whose nodes are
 * arithmetic (the counterpart of src/static_ir/dag.jl:1-46; the reference generates Julia code per node,
 * src/static_ir/generate.jl:68-116) is turned into CUDA source by the host side (gen_b200/staticir.py), compiled with
 * nvcc into a small shared object that instantiates the propagate kernel for it, and registered here. The returned
 * model id is used in gsmc_config.model_id like a catalogue id (default proposal, f64 storage). */
GSMC_API int gsmc_register_model_plugin(const char* path, int* model_id_out);

/* Release the slab pool (column slabs of destroyed filters are cached per device for reuse). */
GSMC_API int gsmc_trim(void);

GSMC_API int gsmc_local_count(gsmc_handle h, uint64_t* n_local, uint64_t* first_global);
GSMC_API int gsmc_state_dim(gsmc_handle h, int* dim);
GSMC_API int gsmc_synchronize(gsmc_handle h);
GSMC_API int gsmc_get_stats(gsmc_handle h, gsmc_stats* out);
/* per-kernel-class CUDA-event timing on the handle's stream (for bench.py's roofline) */
GSMC_API int gsmc_set_profiling(gsmc_handle h, int enabled);
/* CUDA-event stopwatch on the handle's stream */
GSMC_API int gsmc_timer_start(gsmc_handle h);
GSMC_API int gsmc_timer_stop(gsmc_handle h, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* GEN_B200_H */
