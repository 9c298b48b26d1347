/* A non-Python caller of the C ABI (include/gen_b200.h): plain C, compiled with gcc by tests/test_abi.py and linked
 * against libgensmc.so. Exercises struct layout and call order: create -> init -> (maybe_resample -> step)* ->
 * log_ml_estimate / get_log_weights / get_ancestors / get_state / stats -> save/restore -> destroy, plus the error
 * path. Prints the numbers the test compares with the ctypes mirror (same seed => identical bits). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gen_b200.h"

#define CHECK(call)                                                                     \
  do {                                                                                  \
    int rc_ = (call);                                                                   \
    if (rc_ != GSMC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, gsmc_last_error(NULL)); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  const uint64_t N = argc > 1 ? strtoull(argv[1], NULL, 10) : 10000;
  const int T = 12;
  const double params[7] = {0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0};   /* m0, s0, a, b, q, c, r */
  double ys[12];
  for (int t = 0; t < T; ++t) ys[t] = 0.3 * t - 1.0 + ((t * 7) % 5) * 0.21;

  gsmc_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (uint32_t)sizeof cfg;
  cfg.model_id = GSMC_MODEL_LGSSM;
  cfg.dtype = GSMC_F64;
  cfg.resample_scheme = GSMC_RESAMPLE_MULTINOMIAL;
  cfg.num_particles = N;
  cfg.seed = 42;
  cfg.device = -1;
  cfg.keep_history = 1;
  cfg.history_capacity = T;
  gsmc_handle h = NULL;
  CHECK(gsmc_create(&cfg, params, 7, &h));
  CHECK(gsmc_init(h, &ys[0], 1, GSMC_PROPOSAL_DEFAULT, NULL, 0));
  int n_res = 0;
  for (int t = 1; t < T; ++t) {
    int did = 0;
    double ess = 0.0;
    CHECK(gsmc_maybe_resample(h, 0.5 * (double)N, &did, &ess));
    n_res += did;
    CHECK(gsmc_step(h, &ys[t], 1, GSMC_PROPOSAL_DEFAULT, NULL, 0));
  }
  double lml = 0.0;
  CHECK(gsmc_log_ml_estimate(h, &lml));
  double* lw = (double*)malloc(N * sizeof(double));
  double* x = (double*)malloc(N * sizeof(double));
  int64_t* anc = (int64_t*)malloc(N * sizeof(int64_t));
  CHECK(gsmc_get_log_weights(h, lw, N));
  CHECK(gsmc_get_state(h, 0, x, N));
  CHECK(gsmc_get_ancestors(h, anc, N));
  double lw_sum = 0.0, x_sum = 0.0;
  long long anc_sum = 0;
  for (uint64_t i = 0; i < N; ++i) { lw_sum += lw[i]; x_sum += x[i]; anc_sum += (long long)anc[i]; }
  gsmc_stats st;
  CHECK(gsmc_get_stats(h, &st));
  if (st.num_steps != T || st.num_resamples != n_res) { fprintf(stderr, "stats mismatch\n"); return 1; }
  /* error path: a second init must fail with GSMC_E_BADARG and leave the handle usable */
  if (gsmc_init(h, &ys[0], 1, GSMC_PROPOSAL_DEFAULT, NULL, 0) != GSMC_E_BADARG) { fprintf(stderr, "double init not refused\n"); return 1; }
  /* checkpoint -> fresh handle -> same estimate */
  const char* path = argc > 2 ? argv[2] : "/tmp/gsmc_c_abi_smoke.ckpt";
  CHECK(gsmc_save(h, path));
  gsmc_handle h2 = NULL;
  CHECK(gsmc_create(&cfg, params, 7, &h2));
  CHECK(gsmc_restore(h2, path));
  double lml2 = 0.0;
  CHECK(gsmc_log_ml_estimate(h2, &lml2));
  if (lml2 != lml) { fprintf(stderr, "restored estimate differs\n"); return 1; }
  gsmc_destroy(h2);
  gsmc_destroy(h);
  printf("c_abi_smoke ok: version=%s N=%llu resamples=%d log_ml=%.17g lw_sum=%.17g x_sum=%.17g anc_sum=%lld\n",
         gsmc_version(), (unsigned long long)N, n_res, lml, lw_sum, x_sum, anc_sum);
  free(lw); free(x); free(anc);
  return 0;
}
