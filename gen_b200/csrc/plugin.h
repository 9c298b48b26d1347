// plugin.h -- interface between libgensmc.so and a MODEL PLUGIN: a shared object generated from a static-IR description
// of a kernel (gen_b200/staticir.py: IR -> CUDA source -> nvcc, the counterpart of the reference's static-IR code
// generation, src/static_ir/generate.jl:68-116, which emits Julia per-node code into generated functions). A plugin
// instantiates propagate_kernel<GeneratedModel, double, INIT, 0> of kernels.cuh for its model and launches it itself;
// libgensmc.so keeps owning the state, the resampling path and the C ABI.
#ifndef GSMC_PLUGIN_H
#define GSMC_PLUGIN_H
#include <stddef.h>
#include <stdint.h>

#define GSMC_PLUGIN_ABI 3

typedef struct gsmc_plugin_info {
  int abi;                       // GSMC_PLUGIN_ABI the plugin was generated for
  int D;                         // latent columns
  int n_params;                  // model parameters
  int nz_init, nz_step;          // normals per particle (init kernel, step kernel)
  int has_obs_sampler;           // the observation choice can be sampled (unobserved steps)
  size_t sizeof_prop_args, sizeof_model_args, sizeof_dev_scalars;   // layout check of the shared structs
  char name[64];
} gsmc_plugin_info;

typedef int (*gsmc_plugin_describe_fn)(gsmc_plugin_info* out);
// launches the init / step kernel on `stream`; returns a cudaError_t as int; *n_blocks = grid size (logsumexp partials)
typedef int (*gsmc_plugin_propagate_fn)(const void* prop_args, const void* model_args, int init, int64_t n_pad, int sm_count,
                                        void* stream, int pdl, int* n_blocks);
typedef int (*gsmc_plugin_sample_obs_fn)(const void* model_args, const double* state, double* obs_col, int64_t n, int64_t stride,
                                         uint64_t first_global, uint64_t seed, uint32_t t, void* stream);
#endif
