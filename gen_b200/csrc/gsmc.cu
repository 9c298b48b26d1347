// gsmc.cu -- host side of libgensmc.so: the C ABI of include/gen_b200.h.
//
// Owns the device-resident ParticleFilterState (src/inference/particle_filter.jl:18-24 of the
// reference) as structure-of-arrays column slabs and enqueues the kernels of kernels.cuh on one
// CUDA stream. There is no CPU fallback: every entry point either runs the CUDA path or fails.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <dlfcn.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gen_b200.h"
#include "kernels.cuh"
#include "plugin.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define CK(expr)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) return fail(GSMC_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define CKRC(expr)                \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != GSMC_OK) return rc__; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded at run time so that single-GPU use has no NCCL dependency
// ------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return GSMC_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(GSMC_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy)
    return fail(GSMC_E_NCCL, "libnccl is missing required symbols");
  g_nccl.lib = lib;
  return GSMC_OK;
}
#define NK(expr)                                                                                         \
  do {                                                                                                   \
    int r__ = (expr);                                                                                    \
    if (r__ != 0) return fail(GSMC_E_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
  } while (0)
enum { NCCL_UINT8 = 1 };

// ------------------------------------------------------------------------------------------------
// slab pool: the big column slabs of a destroyed filter are kept (per device, keyed by size) and
// handed to the next filter of the same shape, so that repeated initialize_particle_filter calls do not
// pay cudaMalloc/cudaFree of tens of GB each time. gsmc_trim() / GSMC_NO_POOL=1 release / disable it.
// ------------------------------------------------------------------------------------------------
#include <map>
#include <mutex>
struct PoolKey { int device; size_t bytes; bool operator<(const PoolKey& o) const { return device != o.device ? device < o.device : bytes < o.bytes; } };
static std::multimap<PoolKey, void*> g_pool;
static std::mutex g_pool_mu;
static bool pool_enabled() { static int on = getenv("GSMC_NO_POOL") ? 0 : 1; return on != 0; }
static cudaError_t pool_alloc(int device, void** p, size_t bytes) {
  if (pool_enabled()) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool.find(PoolKey{device, bytes});
    if (it != g_pool.end()) { *p = it->second; g_pool.erase(it); return cudaSuccess; }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && pool_enabled()) {          // out of memory: drop the cache and retry once
    cudaGetLastError();
    { std::lock_guard<std::mutex> lk(g_pool_mu); for (auto& kv : g_pool) if (kv.first.device == device) cudaFree(kv.second);
      for (auto it = g_pool.begin(); it != g_pool.end();) { if (it->first.device == device) it = g_pool.erase(it); else ++it; } }
    e = cudaMalloc(p, bytes);
  }
  return e;
}
static std::vector<DevScalars*> g_pinned_pool;
static cudaError_t pinned_alloc(DevScalars** p) {
  { std::lock_guard<std::mutex> lk(g_pool_mu); if (pool_enabled() && !g_pinned_pool.empty()) { *p = g_pinned_pool.back(); g_pinned_pool.pop_back(); return cudaSuccess; } }
  return cudaMallocHost(p, sizeof(DevScalars));
}
static void pinned_free(DevScalars* p) {
  if (!p) return;
  if (pool_enabled()) { std::lock_guard<std::mutex> lk(g_pool_mu); g_pinned_pool.push_back(p); return; }
  cudaFreeHost(p);
}
static void pool_free(int device, void* p, size_t bytes) {
  if (!p) return;
  if (pool_enabled()) { std::lock_guard<std::mutex> lk(g_pool_mu); g_pool.insert({PoolKey{device, bytes}, p}); return; }
  cudaFree(p);
}

// Peer mappings of IPC handles are cached for the life of the process (or until gsmc_trim): the slabs come
// from the pool above, so consecutive filters export the same allocations, and re-opening a multi-GB
// mapping costs tens of milliseconds per peer.
struct IpcKey { int device; char h[sizeof(cudaIpcMemHandle_t)]; bool operator<(const IpcKey& o) const { return device != o.device ? device < o.device : memcmp(h, o.h, sizeof h) < 0; } };
static std::map<IpcKey, void*> g_ipc_cache;
static cudaError_t ipc_open_cached(int device, const cudaIpcMemHandle_t& handle, void** p) {
  IpcKey k; k.device = device; memcpy(k.h, &handle, sizeof k.h);
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_ipc_cache.find(k);
  if (it != g_ipc_cache.end()) { *p = it->second; return cudaSuccess; }
  cudaError_t e = cudaIpcOpenMemHandle(p, handle, cudaIpcMemLazyEnablePeerAccess);
  if (e == cudaSuccess) g_ipc_cache[k] = *p;
  return e;
}

// ------------------------------------------------------------------------------------------------
// model plugins (generated from a static-IR description, see plugin.h): model ids GSMC_MODEL_PLUGIN_BASE + k
// ------------------------------------------------------------------------------------------------
struct ModelPlugin {
  void* lib = nullptr;
  gsmc_plugin_info info;
  gsmc_plugin_propagate_fn propagate = nullptr;
  gsmc_plugin_sample_obs_fn sample_obs = nullptr;
  std::string path;
};
static std::vector<ModelPlugin> g_plugins;
static const ModelPlugin* plugin_of(int model) {
  const int k = model - GSMC_MODEL_PLUGIN_BASE;
  return (k >= 0 && k < (int)g_plugins.size()) ? &g_plugins[k] : nullptr;
}

// ------------------------------------------------------------------------------------------------
// the filter object
// ------------------------------------------------------------------------------------------------
enum KernelClass { KC_PROPAGATE = 0, KC_PROPAGATE_GATHER, KC_FINALIZE, KC_SCAN, KC_SPACINGS, KC_SEARCH, KC_OTHER, KC_COUNT };

struct ProfEvent { cudaEvent_t a, b; int cls; };

struct gsmc_comm_s {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

struct gsmc_filter;
// Shard emulation: R logical ranks of one sharded filter in ONE process on ONE device, all on one stream. The scalar
// exchanges become direct reads of the peers' DevScalars (XMODE_LOCAL); the group calls below enqueue every rank's
// producer kernels before any rank's consumer kernels. Everything else (rank offsets, cross-shard CDF windows,
// boundary gathers through the "peer" pointers) is the code that runs on R GPUs.
struct gsmc_group_s {
  int nranks = 0, device = 0;
  cudaStream_t stream = nullptr;
  gsmc_filter* member[GSMC_MAX_RANKS] = {};
};

struct gsmc_filter {
  gsmc_config cfg;
  int model = 0, D = 0;
  bool f32 = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int rank = 0, nranks = 1;
  NcclComm comm = nullptr;
  int64_t N = 0, n = 0, n_pad = 0, first = 0;   // global count, local count, padded local, first global index
  int sm_count = 148;
  int n_tiles = 0;       // GSMC_TILE (2048) particle tiles of the scan / search kernels
  int n_partials = 0;    // blocks (= logsumexp partials) of the last propagate launch
  std::vector<double> params;
  double* d_params = nullptr;
  double* d_obs = nullptr;
  size_t d_obs_cap = 0;
  // column slabs
  int64_t cap = 2;            // state/ancestor columns (ring of 2, or history_capacity)
  int64_t flag_mod = 4;
  void* state_slab = nullptr; // Real[cap][D][n_pad]
  uint32_t* anc_slab = nullptr;  // uint32[cap][n_pad]
  void* lw = nullptr;         // Real[n_pad]
  uint64_t* cdf = nullptr;    // u64[n_pad] segment-local inclusive CDF of the integer weights, followed by seg_a
  uint64_t* cc = nullptr;     // residual: segment-local inclusive counts of deterministic copies
  uint64_t* seg_a = nullptr;  // u64[n_segs+1] segment totals / exclusive prefixes of the weights (inside the cdf allocation: peers read it)
  uint64_t* seg_b = nullptr;  // u64[n_segs+1] residual scheme: segment prefixes of the residual fractions (inside the cdf allocation too)
  uint64_t* seg_e = nullptr;  // u64[n_segs+1] exclusive prefixes of the group gaps
  uint64_t* raw0 = nullptr;   // u64[n_segs] raw segment totals written by the streaming passes (weights; residual: copies)
  uint64_t* raw1 = nullptr;   // u64[n_segs] raw segment totals (group gaps; residual: fractions)
  uint64_t* tile_e = nullptr; // u64[nt] gap prefix at which every tile opens (segment-local after the gap pass, global after partition)
  uint64_t* gap = nullptr;    // u64[n_pad / 256] Gamma gaps of this rank's groups of sorted draws (fixed point)
  int seg_tiles = 1, n_segs = 0;  // tiles per segment (= per block of the streaming pass), segments per rank
  uint32_t* win = nullptr;          // nt+1 window words of the sorted search
  LseTriple* partials = nullptr;
  DevScalars* ds = nullptr;
  DevScalars* h_ds = nullptr; // pinned mirror
  int* resampled = nullptr;   // device flags, index = step % flag_mod
  double* d_f64 = nullptr;    // staging for f64 conversions / outputs
  size_t d_f64_cap = 0;
  // peers
  const void* peer_slab[GSMC_MAX_RANKS] = {};
  const uint32_t* peer_anc[GSMC_MAX_RANKS] = {};
  const uint64_t* peer_cdf[GSMC_MAX_RANKS] = {};
  DevScalars* peer_ds[GSMC_MAX_RANKS] = {};
  gsmc_group_s* group = nullptr;  // shard emulation (all ranks on this device and stream)
  // gsmc_run_steps as a CUDA graph (see run_steps_graph): the repeated call of one run shape is captured once
  uint64_t graph_key = 0, graph_candidate = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_disabled = false;
  int64_t graph_steps = 0;
  cudaStream_t body_stream = nullptr;            // captures the bodies of the conditional nodes
  unsigned long long cond_finalize = 0, cond_next = 0;   // conditional handles the next finalize / propagate launch sets
  int64_t graph_launches = 0;
  uint32_t host_token = 0;      // token of the last decision published to the pinned host mirror
  bool fuse_next_decide = false;  // gsmc_run_steps: the propagate being launched also decides for the next step
  double fuse_thr = -1.0;
  bool use_nccl_scalars = false;  // GSMC_NCCL_SCALARS=1: exchange the per-step scalars with ncclAllGather instead
  size_t bytes_state = 0, bytes_anc = 0, bytes_lw = 0, bytes_cdf = 0;
  // unobserved steps: sampled observation choices, Real[cap][n_pad], allocated by the first unobserved step
  void* obs_slab = nullptr;
  std::vector<char> unobserved;   // per time step (index t-1): 1 = the observation choice was sampled
  // replay staging
  double* d_zrep = nullptr; size_t zrep_n = 0, zrep_cap = 0;
  double* d_urep = nullptr; size_t urep_n = 0, urep_cap = 0;
  // logical state
  int64_t T = 0;                 // time steps in the traces
  bool decided_since_step = false;
  bool pending = false;          // a resample has been decided but not yet applied by a propagate
  bool stats_fresh = false;
  bool is_importance = false;
  double is_lml = 0.0;           // importance-sampling handles: log_total_weight - log(num_samples) of the run
  int64_t last_resample_step = 0;
  uint32_t n_sample_calls = 0;
  // profiling
  bool profiling = false;
  std::vector<ProfEvent> prof_live, prof_free;
  double prof_ms[KC_COUNT] = {};
  int64_t prof_n[KC_COUNT] = {};
  int64_t launches = 0;
  cudaEvent_t timer_a = nullptr, timer_b = nullptr;
  cudaEvent_t decision_ev = nullptr;   // recorded after the D2H copy of the decision
};

static void drop_graph(gsmc_filter* f) {
  if (f->graph_exec) cudaGraphExecDestroy(f->graph_exec);
  if (f->graph) cudaGraphDestroy(f->graph);
  f->graph_exec = nullptr; f->graph = nullptr; f->graph_key = 0;
}
static size_t real_size(const gsmc_filter* f) { return f->f32 ? 4 : 8; }
static size_t partials_bytes(const gsmc_filter* f) { return (size_t)(f->n_pad / (2 * GSMC_BLOCK * 2)) * sizeof(LseTriple); }
static int xmode(const gsmc_filter* f) {
  if (f->nranks <= 1) return XMODE_NONE;
  if (f->group) return XMODE_LOCAL;
  return f->use_nccl_scalars ? XMODE_NONE : XMODE_LL;
}
static PeerScalars peer_scalars(const gsmc_filter* f) {
  PeerScalars peers;
  for (int r = 0; r < GSMC_MAX_RANKS; ++r) peers.ds[r] = f->peer_ds[r];
  return peers;
}
static char* state_col(const gsmc_filter* f, const void* slab, int64_t step) {
  return (char*)slab + (size_t)((step - 1) % f->cap) * f->D * f->n_pad * real_size(f);
}
static uint32_t* anc_col(const gsmc_filter* f, const uint32_t* slab, int64_t step) {
  return (uint32_t*)slab + (size_t)((step - 1) % f->cap) * f->n_pad;
}

struct ProfScope {
  gsmc_filter* f; ProfEvent ev; bool on;
  ProfScope(gsmc_filter* f_, int cls) : f(f_), on(f_->profiling) {
    f->launches += 1;
    if (!on) return;
    if (!f->prof_free.empty()) { ev = f->prof_free.back(); f->prof_free.pop_back(); }
    else { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
    ev.cls = cls;
    cudaEventRecord(ev.a, f->stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(ev.b, f->stream);
    f->prof_live.push_back(ev);
  }
};
static void harvest_profile(gsmc_filter* f) {
  for (ProfEvent& e : f->prof_live) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      f->prof_ms[e.cls] += ms;
      f->prof_n[e.cls] += 1;
    }
    f->prof_free.push_back(e);
  }
  f->prof_live.clear();
}

static int model_dim(int model) {
  switch (model) {
    case GSMC_MODEL_HMM: return HmmModel::D;
    case GSMC_MODEL_LGSSM: return LgssmModel::D;
    case GSMC_MODEL_SV: return SvModel::D;
    case GSMC_MODEL_BEARINGS: return BearingsModel::D;
    case GSMC_MODEL_REGRESSION: return RegressionModel::D;
    case GSMC_MODEL_NORMAL_NORMAL: return NormalNormalModel::D;
    case GSMC_MODEL_OUTLIER_REGRESSION: return OutlierRegressionModel::D;
    case GSMC_MODEL_UNIFORM_NORMAL: return UniformNormalModel::D;
    default: return plugin_of(model) ? plugin_of(model)->info.D : -1;
  }
}
static bool model_is_importance(int model) {
  return model == GSMC_MODEL_REGRESSION || model == GSMC_MODEL_NORMAL_NORMAL || model == GSMC_MODEL_OUTLIER_REGRESSION || model == GSMC_MODEL_UNIFORM_NORMAL;
}
static bool model_obs_on_device(int model) { return model == GSMC_MODEL_REGRESSION || model == GSMC_MODEL_OUTLIER_REGRESSION; }

static int check_params(int model, const double* p, size_t np) {
  switch (model) {
    case GSMC_MODEL_HMM: {
      if (np < 2) return fail(GSMC_E_BADARG, "HMM params: [K, V, prior, trans, emis]");
      const int K = (int)p[0], V = (int)p[1];
      if (K < 1 || K > GSMC_HMM_MAX_K || V < 1) return fail(GSMC_E_BADARG, "HMM needs 1 <= K <= %d", GSMC_HMM_MAX_K);
      if (np != (size_t)(2 + K + K * K + K * V)) return fail(GSMC_E_BADARG, "HMM params: expected %d values", 2 + K + K * K + K * V);
      return GSMC_OK;
    }
    case GSMC_MODEL_LGSSM: return np == 7 ? GSMC_OK : fail(GSMC_E_BADARG, "LGSSM params: [m0, s0, a, b, q, c, r]");
    case GSMC_MODEL_SV: return np == 3 ? GSMC_OK : fail(GSMC_E_BADARG, "SV params: [mu, phi, sigma]");
    case GSMC_MODEL_BEARINGS: return np == 10 ? GSMC_OK : fail(GSMC_E_BADARG, "bearings params: [m[4], sd[4], sigma_w, sigma_theta]");
    case GSMC_MODEL_REGRESSION:
      if (np < 4 || np != (size_t)(4 + (int)p[0])) return fail(GSMC_E_BADARG, "regression params: [n, sd_slope, sd_intercept, sd_noise, xs[n]]");
      return GSMC_OK;
    case GSMC_MODEL_NORMAL_NORMAL: return np == 3 ? GSMC_OK : fail(GSMC_E_BADARG, "normal-normal params: [mu0, sd0, sd_y]");
    case GSMC_MODEL_OUTLIER_REGRESSION:
      if (np < 4 || np != (size_t)(3 + (int)p[0]) || (int)p[0] < 1 || (int)p[0] > 32 * GSMC_OUTLIER_ZWORDS)
        return fail(GSMC_E_BADARG, "outlier regression params: [n, prob_outlier, prior_sd, xs[n]] with 1 <= n <= %d", 32 * GSMC_OUTLIER_ZWORDS);
      return GSMC_OK;
    case GSMC_MODEL_UNIFORM_NORMAL: return (np == 3 && p[1] > p[0]) ? GSMC_OK : fail(GSMC_E_BADARG, "uniform-normal params: [low, high, sd_y] with high > low");
    default:
      if (const ModelPlugin* pl = plugin_of(model))
        return (int)np == pl->info.n_params ? GSMC_OK : fail(GSMC_E_BADARG, "generated model '%s' takes %d parameters, got %zu", pl->info.name, pl->info.n_params, np);
      return fail(GSMC_E_UNSUPPORTED, "unknown model id %d", model);
  }
}
static int expected_obs(const gsmc_filter* f) {
  if (model_obs_on_device(f->model)) return (int)f->params[0];
  return 1;
}

// ------------------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------------------
static int alloc_buffers(gsmc_filter* f) {
  const size_t rs = real_size(f);
  f->n_pad = (f->n + GSMC_PAD - 1) / GSMC_PAD * GSMC_PAD;
  f->n_tiles = (int)(f->n_pad / GSMC_TILE);
  if (f->n > (int64_t)GSMC_ANC_INDEX_MASK) return fail(GSMC_E_BADARG, "at most 2^28-1 particles per GPU");
  f->cap = f->cfg.keep_history ? (f->cfg.history_capacity > 0 ? f->cfg.history_capacity : 128) : 2;
  if (f->cap < 2) f->cap = 2;
  f->flag_mod = f->cfg.keep_history ? f->cap + 2 : 4;
  f->bytes_state = (size_t)f->cap * f->D * f->n_pad * rs; f->bytes_anc = (size_t)f->cap * f->n_pad * sizeof(uint32_t);
  // segments: one per block of the streaming pass, all blocks resident at once -> one wave
  int occ_w = 0;
  const void* wk = f->f32 ? (const void*)weights_kernel<float, true, true> : (const void*)weights_kernel<double, true, true>;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, wk, GSMC_BLOCK, 0) != cudaSuccess || occ_w < 1) occ_w = 3;
  const int max_segs = f->sm_count * occ_w < GSMC_MAX_SEGS - 1 ? f->sm_count * occ_w : GSMC_MAX_SEGS - 1;   // n_segs + 1 <= 1024 threads
  f->seg_tiles = (f->n_tiles + max_segs - 1) / max_segs;
  f->n_segs = (f->n_tiles + f->seg_tiles - 1) / f->seg_tiles;
  const size_t seg_words = GSMC_MAX_SEGS + 16;
  // the CDF column is followed by two arrays of segment prefixes (weights / copies, residual fractions): peers read all three
  f->bytes_lw = f->n_pad * rs; f->bytes_cdf = (f->n_pad + 2 * seg_words) * sizeof(uint64_t);
  CK(pool_alloc(f->device, &f->state_slab, f->bytes_state));
  CK(pool_alloc(f->device, (void**)&f->anc_slab, f->bytes_anc));
  CK(pool_alloc(f->device, &f->lw, f->bytes_lw));
  CK(pool_alloc(f->device, (void**)&f->cdf, f->bytes_cdf));
  if (f->cfg.resample_scheme == GSMC_RESAMPLE_RESIDUAL) CK(pool_alloc(f->device, (void**)&f->cc, f->n_pad * sizeof(uint64_t)));
  f->seg_a = f->cdf + f->n_pad;
  f->seg_b = f->seg_a + seg_words;
  CK(pool_alloc(f->device, (void**)&f->seg_e, seg_words * sizeof(uint64_t)));
  CK(pool_alloc(f->device, (void**)&f->raw0, seg_words * sizeof(uint64_t)));
  CK(pool_alloc(f->device, (void**)&f->raw1, seg_words * sizeof(uint64_t)));
  CK(pool_alloc(f->device, (void**)&f->tile_e, (size_t)f->n_tiles * sizeof(uint64_t)));
  CK(pool_alloc(f->device, (void**)&f->gap, (size_t)(f->n_pad / GSMC_GROUP) * sizeof(uint64_t)));
  CK(pool_alloc(f->device, (void**)&f->win, (size_t)(f->n_tiles + 1) * sizeof(uint32_t)));
  // one logsumexp partial per propagate block; the grid is at most one block per propagate tile, and the smallest
  // propagate tile (wide states: 2 pairs per thread) is half a GSMC_TILE
  CK(pool_alloc(f->device, (void**)&f->partials, partials_bytes(f)));
  CK(pool_alloc(f->device, (void**)&f->ds, sizeof(DevScalars)));
  CK(pinned_alloc(&f->h_ds));
  memset(f->h_ds, 0, sizeof(DevScalars));          // pooled buffer: no stale decision token
  f->host_token = 0;
  CK(pool_alloc(f->device, (void**)&f->resampled, (size_t)f->flag_mod * sizeof(int)));
  CK(cudaMemsetAsync(f->ds, 0, sizeof(DevScalars), f->stream));
  CK(cudaMemsetAsync(f->resampled, 0, (size_t)f->flag_mod * sizeof(int), f->stream));
  // pad lanes of the log-weight column are read by vector loads: keep them finite and harmless
  CK(cudaMemsetAsync(f->lw, 0, f->n_pad * rs, f->stream));
  // state slabs are written before they are read, so they are not cleared; the pad words of the ancestor columns
  // are read (and used as indices) by the gathering propagate and never written by the search: zero them
  if (f->n_pad > f->n)
    CK(cudaMemset2DAsync(f->anc_slab + f->n, f->n_pad * sizeof(uint32_t), 0, (size_t)(f->n_pad - f->n) * sizeof(uint32_t), (size_t)f->cap, f->stream));
  f->peer_slab[f->rank] = f->state_slab;
  f->peer_anc[f->rank] = f->anc_slab;
  f->peer_cdf[f->rank] = f->cdf;
  f->peer_ds[f->rank] = f->ds;
  return GSMC_OK;
}
static void free_buffers(gsmc_filter* f) {
  for (int r = 0; r < f->nranks; ++r) {
    if (r == f->rank) continue;
    // peer mappings stay open in g_ipc_cache
    f->peer_slab[r] = nullptr; f->peer_anc[r] = nullptr; f->peer_cdf[r] = nullptr; f->peer_ds[r] = nullptr;
  }
  pool_free(f->device, f->state_slab, f->bytes_state); pool_free(f->device, f->anc_slab, f->bytes_anc);
  pool_free(f->device, f->lw, f->bytes_lw); pool_free(f->device, f->cdf, f->bytes_cdf); pool_free(f->device, f->cc, f->n_pad * sizeof(uint64_t));
  const size_t seg_words = GSMC_MAX_SEGS + 16;
  pool_free(f->device, f->seg_e, seg_words * sizeof(uint64_t));
  pool_free(f->device, f->raw0, seg_words * sizeof(uint64_t)); pool_free(f->device, f->raw1, seg_words * sizeof(uint64_t));
  pool_free(f->device, f->tile_e, (size_t)f->n_tiles * sizeof(uint64_t));
  pool_free(f->device, f->gap, (size_t)(f->n_pad / GSMC_GROUP) * sizeof(uint64_t)); pool_free(f->device, f->win, (size_t)(f->n_tiles + 1) * sizeof(uint32_t));
  pool_free(f->device, f->partials, partials_bytes(f)); pool_free(f->device, f->ds, sizeof(DevScalars));
  pool_free(f->device, f->resampled, (size_t)f->flag_mod * sizeof(int));
  pinned_free(f->h_ds);
  pool_free(f->device, f->obs_slab, (size_t)f->cap * f->n_pad * real_size(f)); f->obs_slab = nullptr;
  f->state_slab = nullptr; f->anc_slab = nullptr; f->lw = nullptr; f->cdf = nullptr; f->cc = nullptr;
  f->seg_a = f->seg_b = f->seg_e = f->tile_e = f->raw0 = f->raw1 = nullptr; f->gap = nullptr; f->win = nullptr; f->partials = nullptr; f->ds = nullptr; f->h_ds = nullptr;
  f->resampled = nullptr;
}
static int ensure_f64(gsmc_filter* f, size_t n) {
  if (n <= f->d_f64_cap) return GSMC_OK;
  cudaFree(f->d_f64);
  f->d_f64 = nullptr; f->d_f64_cap = 0;
  CK(cudaMalloc(&f->d_f64, n * sizeof(double)));
  f->d_f64_cap = n;
  return GSMC_OK;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
// Hot-path kernels are launched with programmatic stream serialization (see pdl_wait in kernels.cuh): the
// launch latency and the prologue of kernel k+1 overlap the tail of kernel k. GSMC_NO_PDL=1 turns it off.
// GSMC_PDL_MASK=<bits> selects the kernel classes that may start early (bit 0 propagate, 1 finalize, 2 weights,
// 3 partition, 4 search, 5 everything else). Default 0x3f: every class. Measured on B200 (cfg 3, ms per run): while
// search_sorted_kernel triggered its dependents early, all classes 20.47, all but propagate 19.84, none 20.70; with the
// trigger of the search kernel left to block exit, all classes 19.52 (profiles/r2_pdl_masks.txt).
enum { PDL_PROPAGATE = 0, PDL_FINALIZE, PDL_WEIGHTS, PDL_PARTITION, PDL_SEARCH, PDL_OTHER };
static unsigned pdl_mask() {
  static unsigned mask = getenv("GSMC_NO_PDL") ? 0u : (getenv("GSMC_PDL_MASK") ? (unsigned)strtoul(getenv("GSMC_PDL_MASK"), nullptr, 0) : 0x3fu);
  return mask;
}
// While a run is being captured into a graph, the kernel that follows a conditional node (and the first kernel of a
// conditional body) must not carry the programmatic-dependency attribute: its predecessor is not a kernel node.
static thread_local bool g_pdl_suppress_next = false;
template <int CLS = PDL_OTHER, typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (((pdl_mask() >> CLS) & 1u) && !g_pdl_suppress_next) ? 1 : 0;
  g_pdl_suppress_next = false;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// everything of PropArgs that does not depend on the model: columns, flags, RNG keys, replay buffers (checked against nz / nu)
template <typename Real>
static int fill_prop_args(gsmc_filter* f, PropArgs<Real>& g, bool INIT, int nz, int nu, bool use_anc) {
  memset(&g, 0, sizeof g);
  const int64_t new_step = f->T + 1;
  for (int r = 0; r < f->nranks; ++r)
    g.cur[r] = INIT ? nullptr : (const Real*)state_col(f, f->peer_slab[r], f->T);
  g.nxt = (Real*)state_col(f, f->state_slab, new_step);
  g.lw = (Real*)f->lw;
  g.anc = anc_col(f, f->anc_slab, new_step);
  g.resampled_flag = f->resampled + (new_step % f->flag_mod);
  g.partials = f->partials;
  g.ds = f->ds; g.nranks = f->nranks;
  g.fuse_decide = (f->fuse_next_decide && (f->nranks == 1 || xmode(f) == XMODE_LL)) ? 1 : 0;
  g.peers = peer_scalars(f);
  g.cond_handle = g.fuse_decide ? f->cond_next : 0;
  g.fuse_threshold = f->fuse_thr; g.n_global = (double)f->N;
  g.next_flag = f->resampled + ((new_step + 1) % f->flag_mod);
  g.n = f->n; g.stride = f->n_pad; g.first_global = (uint64_t)f->first; g.seed = f->cfg.seed; g.keys = make_philox_keys(f->cfg.seed);
  g.t = (uint32_t)new_step;
  g.use_anc = use_anc ? 1 : 0;
  g.rank = f->rank;
  g.zrep = f->zrep_n ? f->d_zrep : nullptr;
  g.urep = f->urep_n ? f->d_urep : nullptr;
  if (g.zrep && f->zrep_n != (size_t)(f->n * nz)) return fail(GSMC_E_BADARG, "replay normals: expected %lld values", (long long)(f->n * nz));
  if (g.urep && nu && f->urep_n != (size_t)(f->n * nu)) return fail(GSMC_E_BADARG, "replay uniforms: expected %lld values", (long long)(f->n * nu));
  if (!nu) g.urep = nullptr;
  if (!nz) g.zrep = nullptr;
  return GSMC_OK;
}
template <class Model, typename Real, bool INIT, int PROP>
static int launch_propagate_t(gsmc_filter* f, ModelArgs& a, bool use_anc) {
  Model::prepare(a, INIT, PROP);
  PropArgs<Real> g;
  // models that draw a run-time number of uniforms per particle (DrawCtx) take p[0] of them
  const int nz = Model::nz(INIT, PROP), nu = model_ctx_uniforms<Model>::value ? (int)f->params[0] : Model::nu(INIT, PROP);
  CKRC(fill_prop_args<Real>(f, g, INIT, nz, nu, use_anc));
  {
    ProfScope ps(f, (use_anc && f->pending) ? KC_PROPAGATE_GATHER : KC_PROPAGATE);
    // persistent grid: as many blocks as are resident at once (occupancy of this instantiation), at most one per tile
    static int occ_dev[64] = {};               // per device: occupancy is a property of (kernel, device)
    int& occ = occ_dev[f->device & 63];
    if (occ == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)propagate_kernel<Model, Real, INIT, PROP>, GSMC_BLOCK,
                                                        Model::SMEM_DOUBLES * sizeof(double)) != cudaSuccess || occ < 1) occ = 2;
    }
    static_assert(PropTile<Model>::TILE >= 2 * GSMC_BLOCK * 2, "partials_bytes() assumes at least 2 pairs per thread");
    g.n_tiles = (int)(f->n_pad / PropTile<Model>::TILE);
    f->n_partials = g.n_tiles < f->sm_count * occ ? g.n_tiles : f->sm_count * occ;
    CK(launch_pdl<PDL_PROPAGATE>(propagate_kernel<Model, Real, INIT, PROP>, f->n_partials, GSMC_BLOCK, Model::SMEM_DOUBLES * sizeof(double), f->stream, g, a));
  }
  CK(cudaGetLastError());
  return GSMC_OK;
}
template <class Model, typename Real>
static int launch_propagate_m(gsmc_filter* f, ModelArgs& a, bool init, int prop, bool use_anc) {
  if (!Model::has_proposal(prop)) return fail(GSMC_E_UNSUPPORTED, "model %d has no proposal %d in the catalogue", f->model, prop);
  if (init) return prop ? launch_propagate_t<Model, Real, true, 1>(f, a, use_anc) : launch_propagate_t<Model, Real, true, 0>(f, a, use_anc);
  return prop ? launch_propagate_t<Model, Real, false, 1>(f, a, use_anc) : launch_propagate_t<Model, Real, false, 0>(f, a, use_anc);
}
// a generated model: the plugin launches its own instantiation of propagate_kernel
static int launch_propagate_plugin(gsmc_filter* f, const ModelPlugin* pl, ModelArgs& a, bool init, int prop, bool use_anc) {
  if (prop != GSMC_PROPOSAL_DEFAULT) return fail(GSMC_E_UNSUPPORTED, "generated model '%s' has no custom proposal", pl->info.name);
  if (f->f32) return fail(GSMC_E_UNSUPPORTED, "generated models are compiled for f64 storage");
  PropArgs<double> g;
  CKRC(fill_prop_args<double>(f, g, init, init ? pl->info.nz_init : pl->info.nz_step, 0, use_anc));
  int n_blocks = 0;
  {
    ProfScope ps(f, (use_anc && f->pending) ? KC_PROPAGATE_GATHER : KC_PROPAGATE);
    const bool pdl = ((pdl_mask() >> PDL_PROPAGATE) & 1u) && !g_pdl_suppress_next;
    g_pdl_suppress_next = false;
    const int e = pl->propagate(&g, &a, init ? 1 : 0, f->n_pad, f->sm_count, (void*)f->stream, pdl ? 1 : 0, &n_blocks);
    if (e != 0) return fail(GSMC_E_CUDA, "generated model '%s': kernel launch failed: %s", pl->info.name, cudaGetErrorString((cudaError_t)e));
  }
  f->n_partials = n_blocks;
  return GSMC_OK;
}
template <typename Real>
static int launch_propagate_r(gsmc_filter* f, ModelArgs& a, bool init, int prop, bool use_anc) {
  if (const ModelPlugin* pl = plugin_of(f->model)) return launch_propagate_plugin(f, pl, a, init, prop, use_anc);
  switch (f->model) {
    case GSMC_MODEL_HMM: return launch_propagate_m<HmmModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_LGSSM: return launch_propagate_m<LgssmModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_SV: return launch_propagate_m<SvModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_BEARINGS: return launch_propagate_m<BearingsModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_REGRESSION: return launch_propagate_m<RegressionModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_NORMAL_NORMAL: return launch_propagate_m<NormalNormalModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_OUTLIER_REGRESSION:
      if (f->f32) return fail(GSMC_E_UNSUPPORTED, "the outlier-regression model packs its flags into exact f64 integers: dtype must be f64");
      return launch_propagate_m<OutlierRegressionModel, Real>(f, a, init, prop, use_anc);
    case GSMC_MODEL_UNIFORM_NORMAL: return launch_propagate_m<UniformNormalModel, Real>(f, a, init, prop, use_anc);
  }
  return fail(GSMC_E_UNSUPPORTED, "unknown model");
}

static int fill_model_args(gsmc_filter* f, ModelArgs& a, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp) {
  memset(&a, 0, sizeof a);
  const int need = expected_obs(f);
  static const double dummy_obs[1] = {1.0};
  if (!obs && n_obs == 0) {
    // Unobserved step (the reference samples the unconstrained observation choice, static_ir/generate.jl:36-42)
    if (model_is_importance(f->model)) return fail(GSMC_E_BADARG, "importance sampling needs the observations");
    if (prop != GSMC_PROPOSAL_DEFAULT) return fail(GSMC_E_BADARG, "the catalogue's custom proposals condition on the observation: an unobserved step takes the default proposal");
    a.unobserved = 1;
    obs = dummy_obs; n_obs = 1;
  }
  if (!obs || (int)n_obs != need) return fail(GSMC_E_BADARG, "model %d needs %d observation value(s) per step, got %zu", f->model, need, n_obs);
  if (npp > 8) return fail(GSMC_E_BADARG, "at most 8 proposal parameters");
  if (prop != GSMC_PROPOSAL_DEFAULT && prop != GSMC_PROPOSAL_CUSTOM) return fail(GSMC_E_BADARG, "bad proposal id %d", prop);
  a.n_p = (int)f->params.size(); a.n_obs = (int)n_obs;
  for (size_t i = 0; i < f->params.size() && i < GSMC_MAX_INLINE_PARAMS; ++i) a.p[i] = f->params[i];
  for (size_t i = 0; i < n_obs && i < GSMC_MAX_INLINE_OBS; ++i) a.obs[i] = obs[i];
  for (size_t i = 0; i < npp; ++i) a.pp[i] = pp[i];
  a.p_dev = f->d_params;
  a.obs_dev = nullptr;
  if (model_obs_on_device(f->model)) {
    if (f->model == GSMC_MODEL_REGRESSION && prop == GSMC_PROPOSAL_CUSTOM && npp != 4) return fail(GSMC_E_BADARG, "regression proposal params: [mu_slope, sd_slope, mu_intercept, sd_intercept]");
    if (n_obs > f->d_obs_cap) {
      cudaFree(f->d_obs); f->d_obs = nullptr; f->d_obs_cap = 0;
      CK(cudaMalloc(&f->d_obs, n_obs * sizeof(double)));
      f->d_obs_cap = n_obs;
    }
    CK(cudaMemcpyAsync(f->d_obs, obs, n_obs * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    CK(cudaStreamSynchronize(f->stream));      // obs is a borrowed host pointer
    a.obs_dev = f->d_obs;
  }
  if (f->model == GSMC_MODEL_NORMAL_NORMAL && prop == GSMC_PROPOSAL_CUSTOM && npp != 2)
    return fail(GSMC_E_BADARG, "normal-normal proposal params: [mu_q, sd_q]");
  if (f->model == GSMC_MODEL_UNIFORM_NORMAL && prop == GSMC_PROPOSAL_CUSTOM && (npp != 2 || !(pp[1] > pp[0])))
    return fail(GSMC_E_BADARG, "uniform-normal proposal params: [low_q, high_q] with high_q > low_q");
  if (f->model == GSMC_MODEL_HMM && !a.unobserved) {
    const int V = (int)f->params[1];
    if (!(obs[0] >= 1 && obs[0] <= V) || obs[0] != (double)(int)obs[0]) return fail(GSMC_E_BADARG, "HMM observation must be an integer in 1..%d", V);
  }
  return GSMC_OK;
}

// the sampled observation choice of the step that launch_propagate has just produced (column of step f->T + 1)
template <typename Real>
static int launch_sample_obs(gsmc_filter* f, const ModelArgs& a) {
  const size_t rs = real_size(f);
  if (!f->obs_slab) {
    CK(pool_alloc(f->device, &f->obs_slab, (size_t)f->cap * f->n_pad * rs));
    CK(cudaMemsetAsync(f->obs_slab, 0, (size_t)f->cap * f->n_pad * rs, f->stream));
  }
  const int64_t step = f->T + 1;
  const Real* state = (const Real*)state_col(f, f->state_slab, step);
  Real* col = (Real*)((char*)f->obs_slab + (size_t)((step - 1) % f->cap) * f->n_pad * rs);
  const int grid = (int)((f->n + GSMC_BLOCK - 1) / GSMC_BLOCK);
  ProfScope ps(f, KC_OTHER);
  switch (f->model) {
    case GSMC_MODEL_HMM: sample_obs_kernel<HmmModel, Real><<<grid, GSMC_BLOCK, 0, f->stream>>>(a, state, col, f->n, f->n_pad, (uint64_t)f->first, f->cfg.seed, (uint32_t)step); break;
    case GSMC_MODEL_LGSSM: sample_obs_kernel<LgssmModel, Real><<<grid, GSMC_BLOCK, 0, f->stream>>>(a, state, col, f->n, f->n_pad, (uint64_t)f->first, f->cfg.seed, (uint32_t)step); break;
    case GSMC_MODEL_SV: sample_obs_kernel<SvModel, Real><<<grid, GSMC_BLOCK, 0, f->stream>>>(a, state, col, f->n, f->n_pad, (uint64_t)f->first, f->cfg.seed, (uint32_t)step); break;
    case GSMC_MODEL_BEARINGS: sample_obs_kernel<BearingsModel, Real><<<grid, GSMC_BLOCK, 0, f->stream>>>(a, state, col, f->n, f->n_pad, (uint64_t)f->first, f->cfg.seed, (uint32_t)step); break;
    default: {
      const ModelPlugin* pl = plugin_of(f->model);
      if (!pl || !pl->sample_obs || !pl->info.has_obs_sampler || f->f32) return fail(GSMC_E_UNSUPPORTED, "model %d has no sampler for its observation choice", f->model);
      const int e = pl->sample_obs(&a, (const double*)state, (double*)col, f->n, f->n_pad, (uint64_t)f->first, f->cfg.seed, (uint32_t)step, (void*)f->stream);
      if (e != 0) return fail(GSMC_E_CUDA, "generated model '%s': observation sampler failed: %s", pl->info.name, cudaGetErrorString((cudaError_t)e));
    }
  }
  CK(cudaGetLastError());
  return GSMC_OK;
}

static int launch_propagate(gsmc_filter* f, bool init, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp, bool use_anc) {
  ModelArgs a;
  CKRC(fill_model_args(f, a, obs, n_obs, prop, pp, npp));
  if (f->cfg.keep_history && f->T + 1 > f->cap) return fail(GSMC_E_BADARG, "history_capacity (%lld steps) exceeded", (long long)f->cap);
  int rc = f->f32 ? launch_propagate_r<float>(f, a, init, prop, use_anc) : launch_propagate_r<double>(f, a, init, prop, use_anc);
  f->zrep_n = 0; f->urep_n = 0;
  if (rc == GSMC_OK) {
    if ((int64_t)f->unobserved.size() < f->T + 1) f->unobserved.resize(f->T + 1, 0);
    f->unobserved[f->T] = (char)a.unobserved;
    if (a.unobserved) rc = f->f32 ? launch_sample_obs<float>(f, a) : launch_sample_obs<double>(f, a);
  }
  return rc;
}

// finalize (+ decision when ess_threshold >= 0); leaves the statistics in f->ds
static int launch_finalize(gsmc_filter* f, double ess_threshold, bool to_host = false) {
  int* flag = ess_threshold >= 0.0 ? f->resampled + ((f->T + 1) % f->flag_mod) : nullptr;
  const int xm = xmode(f);
  const bool nccl = f->nranks > 1 && xm == XMODE_NONE;
  // to_host: the deciding thread also writes the decision into the pinned host mirror, tagged with a fresh token
  DevScalars* host = to_host ? f->h_ds : nullptr;
  if (to_host) { f->host_token += 1; if (f->host_token == 0) f->host_token = 1; }
  {
    ProfScope ps(f, KC_FINALIZE);
    CK(launch_pdl<PDL_FINALIZE>(finalize_kernel, 1, 32, 0, f->stream, f->ds, f->rank, f->nranks, ess_threshold, (double)f->N, flag, (int64_t)(f->T + 1),
                                peer_scalars(f), xm, nccl ? (DevScalars*)nullptr : host, f->host_token, f->cond_finalize));
  }
  CK(cudaGetLastError());
  if (nccl) {
    NK(g_nccl.AllGather((const char*)f->ds->triples + f->rank * sizeof(LseTriple), f->ds->triples, sizeof(LseTriple), NCCL_UINT8, f->comm, f->stream));
    ProfScope ps(f, KC_FINALIZE);
    decide_kernel<<<1, 32, 0, f->stream>>>(f->ds, f->nranks, ess_threshold, (double)f->N, flag, (int64_t)(f->T + 1), host, f->host_token);
    CK(cudaGetLastError());
  }
  return GSMC_OK;
}
// Host side of publish_decision: spin on the token in the pinned mirror; falls back to a stream synchronisation +
// copy if the token does not show up (e.g. a platform without device-mapped pinned memory).
static int wait_decision(gsmc_filter* f) {
  volatile unsigned int* tok = &f->h_ds->host_token;
  const auto t0 = std::chrono::steady_clock::now();
  for (long spins = 0;; ++spins) {
    if (*tok == f->host_token) { std::atomic_thread_fence(std::memory_order_acquire); return GSMC_OK; }
    if ((spins & 0x3ff) == 0x3ff) {
      if (cudaStreamQuery(f->stream) == cudaSuccess && *tok != f->host_token) break;          // all work done, no token: fall back
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) break;
    }
  }
  CK(cudaMemcpyAsync(f->h_ds, f->ds, offsetof(DevScalars, mbox), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}
// All ranks have finished every kernel that reads this rank's slabs (needed before they are reused or freed).
static int peer_barrier(gsmc_filter* f) {
  if (f->nranks <= 1 || !f->ds || !f->peer_ds[(f->rank + 1) % f->nranks]) return GSMC_OK;
  if (f->group) return GSMC_OK;                     // one stream: stream order is the barrier
  { ProfScope ps(f, KC_OTHER); peer_barrier_kernel<<<1, 32, 0, f->stream>>>(peer_scalars(f), f->ds, f->rank, f->nranks); }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

// sticky device-side error -> status code (1: total weight zero / not finite at a resample; 2: a peer never answered)
static int device_error(gsmc_filter* f, const char* when) {
  const int e = f->h_ds->error;
  if (!e) return GSMC_OK;
  cudaMemsetAsync(&f->ds->error, 0, sizeof(int), f->stream);
  if (e == 2) return fail(GSMC_E_PEER, "a peer GPU did not answer a scalar exchange within the time limit %s; the sharded filter is no longer consistent", when);
  return fail(GSMC_E_DEGENERATE, "total weight is zero or not finite %s (log_total = %g)", when, f->h_ds->log_total);
}

static int fetch_scalars(gsmc_filter* f) {
  CK(cudaMemcpyAsync(f->h_ds, f->ds, offsetof(DevScalars, mbox), cudaMemcpyDeviceToHost, f->stream));   // the scalars, not the mailboxes
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

static CdfView make_cdf_view(const gsmc_filter* f, bool residual_fractions) {
  CdfView v;
  memset(&v, 0, sizeof v);
  const size_t seg_words = GSMC_MAX_SEGS + 16;
  // residual scheme: the searched CDF is the one of the residual fractions, whose segment prefixes are the second array
  for (int r = 0; r < f->nranks; ++r) { v.seg[r] = f->peer_cdf[r]; v.sp[r] = f->peer_cdf[r] + f->n_pad + (residual_fractions ? seg_words : 0); }
  v.n_per = f->n;
  v.n_pad = (int)f->n_pad;
  v.seg_len = f->seg_tiles * GSMC_TILE;
  v.n_segs = f->n_segs;
  v.nranks = f->nranks;
  return v;
}
// the ancestor column of the step being produced on every rank (for the remote stores of the sharded residual scheme)
static AncOut make_anc_out(const gsmc_filter* f, int64_t step) {
  AncOut a;
  memset(&a, 0, sizeof a);
  for (int r = 0; r < f->nranks; ++r) a.col[r] = anc_col(f, f->peer_anc[r], step);
  a.n_per = f->n; a.nranks = f->nranks;
  return a;
}
static double weight_scale(const gsmc_filter* f) {
  int lg = 0;
  while (((uint64_t)1 << lg) < (uint64_t)f->N) ++lg;
  const int k = 62 - lg;
  return gm_pow2(k > 52 ? 52 : k);
}
// One-block scan of the raw segment totals in0/in1 into the prefix arrays out0/out1 + exchange of this rank's
// totals + the event's totals (see kernels.cuh). phases: bit 0 = the scan (fused exchange included), bit 1 = what
// follows it when the exchange is not fused (ncclAllGather, or the direct peer reads of the shard emulation).
// Every exchange of per-rank scalars separates two phases; the multinomial scheme has one (PH_LOCAL | PH_GLOBAL), the
// residual scheme three (weights totals; copy and fraction totals; gap totals): PH_LOCAL, PH_R1, PH_R2, PH_GLOBAL.
enum { PH_LOCAL = 1, PH_R1 = 2, PH_R2 = 4, PH_GLOBAL = 8, PH_ALL = 15 };
static int launch_scan(gsmc_filter* f, int cls, const uint64_t* in0, const uint64_t* in1, uint64_t* out0, uint64_t* out1, int what, int conditional,
                       bool do_local = true, bool do_global = true) {
  const bool multi = f->nranks > 1;
  const int xm = xmode(f);
  const bool fused = multi && xm == XMODE_LL;
  if (do_local) {
    ProfScope ps(f, cls);
    CK(launch_pdl(scan_segments_kernel, 1, 1024, 0, f->stream, in0, in1, out0, out1, f->n_segs, f->ds, what, f->cfg.seed, (uint64_t)f->N, conditional,
                  peer_scalars(f), f->rank, f->nranks, fused ? 1 : 0));
    CK(cudaGetLastError());
  }
  if (do_global && multi && !fused) {
    if (xm == XMODE_LOCAL) {
      ProfScope ps(f, KC_OTHER);
      peer_copy_kernel<<<1, 32, 0, f->stream>>>(peer_scalars(f), f->ds, f->rank, f->nranks,
          ((what & SCAN_Q) ? PEER_COPY_CDF : 0) | ((what & SCAN_E) ? PEER_COPY_GAP : 0) | ((what & SCAN_RESID) ? PEER_COPY_DET : 0), conditional);
    } else {
      if (what & SCAN_Q) NK(g_nccl.AllGather((const char*)(f->ds->cdf_rank_total + f->rank), f->ds->cdf_rank_total, sizeof(uint64_t), NCCL_UINT8, f->comm, f->stream));
      if (what & SCAN_RESID) NK(g_nccl.AllGather((const char*)(f->ds->frac_rank_total + f->rank), f->ds->frac_rank_total, sizeof(uint64_t), NCCL_UINT8, f->comm, f->stream));
      if (what & SCAN_E) NK(g_nccl.AllGather((const char*)(f->ds->gap_rank_total + f->rank), f->ds->gap_rank_total, sizeof(uint64_t), NCCL_UINT8, f->comm, f->stream));
      if (what & SCAN_RESID) NK(g_nccl.AllGather((const char*)(f->ds->det_rank_total + f->rank), f->ds->det_rank_total, sizeof(uint64_t), NCCL_UINT8, f->comm, f->stream));
    }
    ProfScope ps(f, KC_OTHER);
    totals_kernel<<<1, 1024, 0, f->stream>>>(out0, out1, f->n_segs, f->ds, f->nranks, f->rank, f->cfg.seed, (uint64_t)f->N, what, conditional);
    CK(cudaGetLastError());
  }
  return GSMC_OK;
}

// maybe_resample! (particle_filter.jl:199-200) on the device: integer CDF, sorted uniforms, ancestors.
//   multinomial, Philox draws:         weights + group-gaps pass -> partition (scan and the ranks' exchange fused) -> search   (3 launches)
//   exported uniforms (replay): weights pass -> scan -> iid search
//   residual: + the copy counts / residual fractions pass and the deterministic copies
// phases: see PH_* above. A filter on its own runs all of them in one go; the shard emulation runs each phase on every
// rank before the next one (a phase may read what the previous phase produced on ANY rank).
template <typename Real>
static int launch_resample_t(gsmc_filter* f, int conditional, bool replay_iid, int phases) {
  const Real* lw = (const Real*)f->lw;
  const double scale = weight_scale(f);
  const int nt = f->n_tiles, ns = f->n_segs, st = f->seg_tiles;
  const bool residual = f->cfg.resample_scheme == GSMC_RESAMPLE_RESIDUAL;
  const bool multi = f->nranks > 1;
  uint32_t* anc = anc_col(f, f->anc_slab, f->T + 1);
  if (residual && multi && replay_iid) return fail(GSMC_E_UNSUPPORTED, "exported uniforms with residual resampling on a sharded filter");
  // draws are indexed globally; rank r generates (and, multinomial scheme, owns the output slots of) draws [r n, (r+1) n)
  const uint64_t k_first = (uint64_t)f->first;
  const bool fuse_spacings = !residual && !replay_iid;
  // the partition kernel scans the segment totals itself and, on a sharded filter, exchanges the ranks' totals
  // over the peer mailboxes; with GSMC_NCCL_SCALARS=1 (and in the shard emulation) the separate scan + gather path is used
  const bool fuse_scan = fuse_spacings && (!multi || xmode(f) == XMODE_LL);
  const bool L = phases & PH_LOCAL, G = phases & PH_GLOBAL;
  if (L) {
    // 1. integer weights -> segment-local CDF + segment totals (and, fused, the group gaps of the N sorted draws)
    if (fuse_spacings) {
      ProfScope ps(f, KC_SCAN);
      CK(launch_pdl<PDL_WEIGHTS>(weights_kernel<Real, true, true>, ns, GSMC_BLOCK, 0, f->stream,
          lw, f->n, scale, f->ds, f->cdf, f->raw0, f->cfg.seed, k_first, (uint64_t)f->N, f->gap, f->tile_e, f->raw1, nt, st, conditional));
    } else {
      ProfScope ps(f, KC_SCAN);
      CK(launch_pdl<PDL_WEIGHTS>(weights_kernel<Real, true, false>, ns, GSMC_BLOCK, 0, f->stream,
          lw, f->n, scale, f->ds, f->cdf, f->raw0, (uint64_t)0, (uint64_t)0, (uint64_t)0, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint64_t*)nullptr, nt, st, conditional));
    }
    CK(cudaGetLastError());
  }
  // 2. segment prefixes and the totals of the event
  if (fuse_spacings) { if (!fuse_scan) CKRC(launch_scan(f, KC_SCAN, f->raw0, f->raw1, f->seg_a, f->seg_e, SCAN_Q | SCAN_SET_DRAWS | SCAN_E, conditional, L, G)); }
  else if (!residual) CKRC(launch_scan(f, KC_SCAN, f->raw0, nullptr, f->seg_a, nullptr, SCAN_Q | SCAN_SET_DRAWS, conditional, L, G));
  else CKRC(launch_scan(f, KC_SCAN, f->raw0, nullptr, f->seg_a, nullptr, SCAN_Q, conditional, L, (phases & PH_R1) != 0));
  if (residual) {
    if (phases & PH_R1) {
      { ProfScope ps(f, KC_OTHER); CK(launch_pdl(resid_scale_kernel, 1, 32, 0, f->stream, f->ds, (double)f->N)); }
      { ProfScope ps(f, KC_SCAN);
        CK(launch_pdl(resid_cdf_kernel<Real>, ns, GSMC_BLOCK, 0, f->stream, lw, f->n, scale, f->ds, f->cc, f->raw0, f->cdf, f->raw1, nt, st, conditional)); }
    }
    // copy and fraction totals of all ranks -> n_det, M, the total of the fractions
    CKRC(launch_scan(f, KC_SCAN, f->raw0, f->raw1, f->seg_a, f->seg_b, SCAN_RESID, conditional, (phases & PH_R1) != 0, (phases & PH_R2) != 0));
    if (phases & PH_R2) {
      { ProfScope ps(f, KC_SEARCH);
        const int want = (int)((f->n_pad + GSMC_BLOCK - 1) / GSMC_BLOCK), wave = 8 * f->sm_count;
        CK(launch_pdl(det_copies_kernel, want < wave ? want : wave, GSMC_BLOCK, 0, f->stream,
            f->cc, f->seg_a, ns, st * GSMC_TILE, (int)f->n_pad, f->n, f->N, f->rank, f->ds, make_anc_out(f, f->T + 1), conditional)); }
      CK(cudaGetLastError());
      if (!replay_iid) {
        // the number of draws M is only known now: group gaps of this rank's share [r n, (r+1) n) of the M sorted draws
        ProfScope ps(f, KC_SPACINGS);
        CK(launch_pdl<PDL_WEIGHTS>(weights_kernel<Real, false, true>, ns, GSMC_BLOCK, 0, f->stream,
            lw, f->n, scale, f->ds, (uint64_t*)nullptr, (uint64_t*)nullptr, f->cfg.seed, k_first, (uint64_t)0, f->gap, f->tile_e, f->raw1, nt, st, conditional));
      }
    }
    if (!replay_iid) CKRC(launch_scan(f, KC_SPACINGS, f->raw1, nullptr, f->seg_e, nullptr, SCAN_E, conditional, (phases & PH_R2) != 0, G));
  }
  if (!G) return GSMC_OK;
  const CdfView v = make_cdf_view(f, residual);
  if (replay_iid) {
    // one exported uniform per output slot (per multinomial draw in the residual scheme)
    if (!residual && f->urep_n != (size_t)f->n) return fail(GSMC_E_BADARG, "replay uniforms for maybe_resample: expected %lld values", (long long)f->n);
    ProfScope ps(f, KC_SEARCH);
    search_iid_kernel<<<(int)((f->n + GSMC_BLOCK - 1) / GSMC_BLOCK), GSMC_BLOCK, 0, f->stream>>>(
        v, f->ds, f->d_urep, 0, 0, 0, f->n, residual ? 1 : 0, anc, nullptr, conditional);
    f->urep_n = 0;
  } else {
    // 3. ancestors
    uint64_t* sp_q = residual ? f->seg_b : f->seg_a;
    { ProfScope ps(f, KC_SEARCH);
      const int need = (nt + 1 + 31) / 32;
      const int grid = need < 2 * f->sm_count ? need : 2 * f->sm_count;      // one wave of 1024-thread blocks
      if (fuse_scan) CK(launch_pdl<PDL_PARTITION>(partition_kernel<true>, grid, 1024, 0, f->stream, v, k_first, f->rank, f->ds, f->raw0, f->raw1, sp_q, f->seg_e,
                                   f->tile_e, st, nt, f->win, (uint64_t)f->N, conditional, peer_scalars(f)));
      else CK(launch_pdl<PDL_PARTITION>(partition_kernel<false>, grid, 1024, 0, f->stream, v, k_first, f->rank, f->ds, (const uint64_t*)nullptr, (const uint64_t*)nullptr, sp_q, f->seg_e,
                         f->tile_e, st, nt, f->win, (uint64_t)f->N, conditional, peer_scalars(f))); }
    { static bool attr_set[64] = {};           // function attributes are per device
      if (!attr_set[f->device & 63]) { CK(cudaFuncSetAttribute((const void*)search_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GSMC_SEARCH_SMEM)); attr_set[f->device & 63] = true; }
      static int occ_dev[64] = {};
      int& occ = occ_dev[f->device & 63];
      if (occ == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)search_sorted_kernel, GSMC_BLOCK, GSMC_SEARCH_SMEM) != cudaSuccess || occ < 1)) occ = 2;
      const int grid = nt < f->sm_count * occ ? nt : f->sm_count * occ;
      const uint32_t magic = st > 1 ? (uint32_t)(0x100000000ULL / (uint64_t)st) + 1u : 0u;
      ProfScope ps(f, KC_SEARCH);
      CK(launch_pdl<PDL_SEARCH>(search_sorted_kernel, grid, GSMC_BLOCK, GSMC_SEARCH_SMEM, f->stream,
          v, k_first, f->ds, f->tile_e, f->gap, (uint32_t)st, magic, make_philox_keys(f->cfg.seed), f->win, anc, f->n, nt, residual ? 1 : 0, conditional, f->rank,
          make_anc_out(f, f->T + 1))); }
    // sharded residual scheme: deterministic copies and draws were stored into the ranks that own the slots -- nobody
    // gathers through its ancestor column before every rank's stores have landed
    if (residual && multi && !f->group) {
      ProfScope ps(f, KC_OTHER);
      peer_fence_kernel<<<1, 32, 0, f->stream>>>(peer_scalars(f), f->ds, f->rank, f->nranks, conditional);
    }
  }
  CK(cudaGetLastError());
  return GSMC_OK;
}
static int launch_resample(gsmc_filter* f, int conditional, bool replay_iid, int phases = PH_ALL) {
  return f->f32 ? launch_resample_t<float>(f, conditional, replay_iid, phases) : launch_resample_t<double>(f, conditional, replay_iid, phases);
}

template <typename Real>
static HistView<Real> make_hist_view(const gsmc_filter* f) {
  HistView<Real> h;
  memset(&h, 0, sizeof h);
  for (int r = 0; r < f->nranks; ++r) { h.slab[r] = (const Real*)f->peer_slab[r]; h.anc_slab[r] = f->peer_anc[r]; }
  h.resampled = f->resampled; h.stride = f->n_pad; h.D = f->D; h.cap = f->cap; h.flag_mod = f->flag_mod;
  return h;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

GSMC_API const char* gsmc_version(void) { return "gen_b200 0.1.0 (sm_100a)"; }
GSMC_API const char* gsmc_last_error(gsmc_handle) { return g_last_error.c_str(); }

GSMC_API int gsmc_create(const gsmc_config* cfg, const double* params, size_t n_params, gsmc_handle* out) {
  if (!cfg || !out || !params) return fail(GSMC_E_BADARG, "null argument");
  if (cfg->struct_size != sizeof(gsmc_config)) return fail(GSMC_E_BADARG, "gsmc_config.struct_size mismatch (%u vs %zu)", cfg->struct_size, sizeof(gsmc_config));
  *out = nullptr;
  const int D = model_dim(cfg->model_id);
  if (D < 0) return fail(GSMC_E_UNSUPPORTED, "unknown model id %d", cfg->model_id);
  CKRC(check_params(cfg->model_id, params, n_params));
  if (cfg->num_particles < 1) return fail(GSMC_E_BADARG, "num_particles must be >= 1");
  if (cfg->dtype != GSMC_F64 && cfg->dtype != GSMC_F32) return fail(GSMC_E_BADARG, "bad dtype");
  if (cfg->resample_scheme != GSMC_RESAMPLE_MULTINOMIAL && cfg->resample_scheme != GSMC_RESAMPLE_RESIDUAL) return fail(GSMC_E_BADARG, "bad resample_scheme");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1) return fail(GSMC_E_CUDA, "no CUDA device available (%s); libgensmc has no CPU path", cudaGetErrorString(e));
  gsmc_filter* f = new gsmc_filter();
  f->cfg = *cfg;
  f->model = cfg->model_id; f->D = D; f->f32 = cfg->dtype == GSMC_F32;
  f->is_importance = model_is_importance(cfg->model_id);
  if (cfg->device >= 0) f->device = cfg->device; else cudaGetDevice(&f->device);
  if (cudaSetDevice(f->device) != cudaSuccess) { delete f; return fail(GSMC_E_CUDA, "cannot select device %d", cfg->device); }
  if (cfg->stream) f->stream = (cudaStream_t)cfg->stream;
  else { if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) { delete f; return fail(GSMC_E_CUDA, "stream creation failed"); } f->own_stream = true; }
  cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, f->device);
  if (f->sm_count < 1) f->sm_count = 148;
  f->params.assign(params, params + n_params);
  f->N = (int64_t)cfg->num_particles; f->n = f->N; f->first = 0;
  if (pool_alloc(f->device, (void**)&f->d_params, n_params * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(f->d_params, params, n_params * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    gsmc_destroy(f);
    return fail(GSMC_E_CUDA, "parameter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaEventCreate(&f->timer_a); cudaEventCreate(&f->timer_b);
  cudaEventCreateWithFlags(&f->decision_ev, cudaEventDisableTiming);
  *out = f;
  return GSMC_OK;
}

GSMC_API void gsmc_destroy(gsmc_handle f) {
  if (!f) return;
  cudaSetDevice(f->device);
  peer_barrier(f);
  if (f->stream) cudaStreamSynchronize(f->stream);
  harvest_profile(f);
  for (ProfEvent& e : f->prof_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  drop_graph(f);
  if (f->body_stream) cudaStreamDestroy(f->body_stream);
  free_buffers(f);
  pool_free(f->device, f->d_params, f->params.size() * sizeof(double)); cudaFree(f->d_obs); cudaFree(f->d_zrep); cudaFree(f->d_urep); cudaFree(f->d_f64);
  if (f->timer_a) cudaEventDestroy(f->timer_a);
  if (f->timer_b) cudaEventDestroy(f->timer_b);
  if (f->decision_ev) cudaEventDestroy(f->decision_ev);
  if (f->own_stream && f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

GSMC_API int gsmc_reset(gsmc_handle f) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  if (f->is_importance && f->T > 0) return fail(GSMC_E_BADARG, "importance-sampling handles cannot be reset");
  CK(cudaSetDevice(f->device));
  f->T = 0; f->decided_since_step = false; f->pending = false; f->stats_fresh = false;
  f->last_resample_step = 0; f->n_sample_calls = 0; f->zrep_n = 0; f->urep_n = 0;
  f->unobserved.clear();
  CKRC(peer_barrier(f));
  if (f->ds) {
    CK(cudaMemsetAsync(f->ds, 0, offsetof(DevScalars, xseq), f->stream));   // the exchange sequence number and the mailboxes keep their tags
    CK(cudaMemsetAsync(f->resampled, 0, (size_t)f->flag_mod * sizeof(int), f->stream));
  }
  return GSMC_OK;
}

GSMC_API int gsmc_comm_unique_id(void* id_out, size_t nbytes) {
  if (!id_out || nbytes < sizeof(NcclId)) return fail(GSMC_E_BADARG, "id buffer must hold 128 bytes");
  CKRC(load_nccl());
  NK(g_nccl.GetUniqueId((NcclId*)id_out));
  return GSMC_OK;
}

GSMC_API int gsmc_comm_create(const void* unique_id, size_t nbytes, int rank, int nranks, int device, gsmc_comm* out) {
  if (!unique_id || nbytes < sizeof(NcclId) || !out) return fail(GSMC_E_BADARG, "bad arguments");
  if (nranks < 1 || nranks > GSMC_MAX_RANKS || rank < 0 || rank >= nranks) return fail(GSMC_E_BADARG, "1 <= nranks <= %d", GSMC_MAX_RANKS);
  CKRC(load_nccl());
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  gsmc_comm_s* c = new gsmc_comm_s();
  c->rank = rank; c->nranks = nranks; c->device = device;
  NcclId id;
  memcpy(&id, unique_id, sizeof id);
  int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) { delete c; return fail(GSMC_E_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); }
  *out = c;
  return GSMC_OK;
}

GSMC_API void gsmc_comm_destroy(gsmc_comm c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

GSMC_API int gsmc_comm_attach(gsmc_handle f, gsmc_comm c) {
  if (!f || !c) return fail(GSMC_E_BADARG, "bad arguments");
  if (f->T != 0 || f->state_slab) return fail(GSMC_E_BADARG, "attach must precede gsmc_init");
  if (c->device != f->device) return fail(GSMC_E_BADARG, "communicator lives on device %d, filter on device %d", c->device, f->device);
  const int rank = c->rank, nranks = c->nranks;
  if (f->N % ((int64_t)nranks * GSMC_PAD) != 0) return fail(GSMC_E_BADARG, "num_particles must be a multiple of %d * nranks", GSMC_PAD);
  CK(cudaSetDevice(f->device));
  f->rank = rank; f->nranks = nranks; f->comm = c->comm;
  f->n = f->N / nranks; f->first = f->n * rank;
  if (nranks == 1) return GSMC_OK;
  // allocate now and exchange IPC handles of the slabs peers read (state, ancestors, CDF)
  CKRC(alloc_buffers(f));
  struct Handles { cudaIpcMemHandle_t slab, anc, cdf, ds; };
  f->use_nccl_scalars = getenv("GSMC_NCCL_SCALARS") != nullptr;
  Handles mine;
  CK(cudaIpcGetMemHandle(&mine.slab, f->state_slab));
  CK(cudaIpcGetMemHandle(&mine.anc, f->anc_slab));
  CK(cudaIpcGetMemHandle(&mine.cdf, f->cdf));
  CK(cudaIpcGetMemHandle(&mine.ds, f->ds));
  Handles* d_all = nullptr;
  CK(pool_alloc(f->device, (void**)&d_all, sizeof(Handles) * GSMC_MAX_RANKS));
  CK(cudaMemcpyAsync(d_all + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, f->stream));
  NK(g_nccl.AllGather(d_all + rank, d_all, sizeof(Handles), NCCL_UINT8, f->comm, f->stream));
  std::vector<Handles> all(nranks);
  CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * nranks, cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  pool_free(f->device, d_all, sizeof(Handles) * GSMC_MAX_RANKS);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) continue;
    void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr;
    CK(ipc_open_cached(f->device, all[r].slab, &p0));
    CK(ipc_open_cached(f->device, all[r].anc, &p1));
    CK(ipc_open_cached(f->device, all[r].cdf, &p2));
    CK(ipc_open_cached(f->device, all[r].ds, &p3));
    f->peer_slab[r] = p0; f->peer_anc[r] = (const uint32_t*)p1; f->peer_cdf[r] = (const uint64_t*)p2; f->peer_ds[r] = (DevScalars*)p3;
  }
  return GSMC_OK;
}

GSMC_API int gsmc_set_replay(gsmc_handle f, const double* normals, size_t n_normals, const double* uniforms, size_t n_uniforms) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  CK(cudaSetDevice(f->device));
  f->zrep_n = 0; f->urep_n = 0;
  if (normals && n_normals) {
    if (n_normals > f->zrep_cap) { cudaFree(f->d_zrep); f->d_zrep = nullptr; f->zrep_cap = 0; CK(cudaMalloc(&f->d_zrep, n_normals * sizeof(double))); f->zrep_cap = n_normals; }
    CK(cudaMemcpyAsync(f->d_zrep, normals, n_normals * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    f->zrep_n = n_normals;
  }
  if (uniforms && n_uniforms) {
    if (n_uniforms > f->urep_cap) { cudaFree(f->d_urep); f->d_urep = nullptr; f->urep_cap = 0; CK(cudaMalloc(&f->d_urep, n_uniforms * sizeof(double))); f->urep_cap = n_uniforms; }
    CK(cudaMemcpyAsync(f->d_urep, uniforms, n_uniforms * sizeof(double), cudaMemcpyHostToDevice, f->stream));
    f->urep_n = n_uniforms;
  }
  CK(cudaStreamSynchronize(f->stream));   // host pointers are borrowed for this call only
  return GSMC_OK;
}

GSMC_API int gsmc_init(gsmc_handle f, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  if (f->T != 0) return fail(GSMC_E_BADARG, "filter is already initialised");
  CK(cudaSetDevice(f->device));
  if (!f->state_slab) CKRC(alloc_buffers(f));
  CKRC(launch_propagate(f, true, obs, n_obs, prop, pp, npp, false));
  f->T = 1;
  f->decided_since_step = false; f->pending = false; f->stats_fresh = false;
  return GSMC_OK;
}

GSMC_API int gsmc_step(gsmc_handle f, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (f->is_importance) return fail(GSMC_E_BADARG, "importance-sampling models have a single step");
  CK(cudaSetDevice(f->device));
  CKRC(launch_propagate(f, false, obs, n_obs, prop, pp, npp, f->decided_since_step));
  f->T += 1;
  f->decided_since_step = false; f->pending = false; f->stats_fresh = false;
  return GSMC_OK;
}

GSMC_API int gsmc_maybe_resample(gsmc_handle f, double ess_threshold, int* did_resample, double* ess_out) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  if (f->group) return fail(GSMC_E_BADARG, "member of an emulated shard group: call gsmc_group_maybe_resample");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (!(ess_threshold >= 0.0)) return fail(GSMC_E_BADARG, "ess_threshold must be >= 0");
  CK(cudaSetDevice(f->device));
  if (f->pending) {
    // log weights are all zero after a resample: ESS = N exactly (particle_filter.jl:193)
    if ((double)f->N < ess_threshold) return fail(GSMC_E_UNSUPPORTED, "two resampling events without a step in between");
    if (did_resample) *did_resample = 0;
    if (ess_out) *ess_out = (double)f->N;
    return GSMC_OK;
  }
  const bool replay = f->urep_n > 0;
  // The decision is taken on the device and written by the deciding thread into the pinned host mirror. Unless
  // exported uniforms are being replayed, the resampling kernels are enqueued right away in their conditional form
  // (they exit at once when no resample was decided), so the GPU never idles while the host reads the Bool.
  CKRC(launch_finalize(f, ess_threshold, true));
  const size_t ev0 = f->prof_live.size();
  if (!replay) CKRC(launch_resample(f, 1, false));
  CKRC(wait_decision(f));
  // profiling: the conditional launches of a step that did not resample exited at once; book them as "other" so that
  // the scan / search classes time resampling events only
  if (f->profiling && !f->h_ds->do_resample)
    for (size_t k = ev0; k < f->prof_live.size(); ++k)
      if (f->prof_live[k].cls == KC_SCAN || f->prof_live[k].cls == KC_SEARCH || f->prof_live[k].cls == KC_SPACINGS) f->prof_live[k].cls = KC_OTHER;
  f->stats_fresh = true;
  f->decided_since_step = true;
  if (f->h_ds->error) {
    f->decided_since_step = false;
    return device_error(f, "in maybe_resample");
  }
  const int did = f->h_ds->do_resample;
  if (did) {
    if (replay) CKRC(launch_resample(f, 0, true));
    f->pending = true;
    f->last_resample_step = f->T + 1;
  }
  f->urep_n = 0;
  if (did_resample) *did_resample = did;
  if (ess_out) *ess_out = f->h_ds->ess;
  return GSMC_OK;
}

static int refresh_stats(gsmc_filter* f) {
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (!f->stats_fresh && !f->pending) {
    CKRC(launch_finalize(f, -1.0));
    CKRC(fetch_scalars(f));
    f->stats_fresh = true;
  } else if (!f->stats_fresh) {
    CKRC(fetch_scalars(f));
  }
  return GSMC_OK;
}

GSMC_API int gsmc_log_ml_estimate(gsmc_handle f, double* out) {
  if (!f || !out) return fail(GSMC_E_BADARG, "null argument");
  CK(cudaSetDevice(f->device));
  if (f->is_importance) { *out = f->is_lml; return GSMC_OK; }      // importance.jl:30,49 (the handle holds normalised weights)
  CKRC(refresh_stats(f));
  // particle_filter.jl:52-55; after a resample the log weights are zero: logsumexp = log N
  const double logn = gm_log((double)f->N);
  *out = f->pending ? f->h_ds->log_ml_est + (logn - logn) : f->h_ds->log_ml_est + f->h_ds->log_total - logn;
  return GSMC_OK;
}

GSMC_API int gsmc_get_log_weights(gsmc_handle f, double* host_dst, size_t n) {
  if (!f || !host_dst) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (n != (size_t)f->n) return fail(GSMC_E_BADARG, "expected a buffer of %lld values", (long long)f->n);
  CK(cudaSetDevice(f->device));
  if (f->pending) { for (size_t i = 0; i < n; ++i) host_dst[i] = 0.; return GSMC_OK; }   // particle_filter.jl:204
  if (!f->f32) {
    CK(cudaMemcpyAsync(host_dst, f->lw, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  } else {
    CKRC(ensure_f64(f, n));
    { ProfScope ps(f, KC_OTHER); column_to_f64_kernel<float><<<(int)((n + GSMC_BLOCK - 1) / GSMC_BLOCK), GSMC_BLOCK, 0, f->stream>>>((const float*)f->lw, f->d_f64, (int64_t)n); }
    CK(cudaMemcpyAsync(host_dst, f->d_f64, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  }
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

GSMC_API int gsmc_get_log_weights_device(gsmc_handle f, void** dev_ptr) {
  if (!f || !dev_ptr) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (f->pending) CK(cudaMemsetAsync(f->lw, 0, f->n_pad * real_size(f), f->stream));
  *dev_ptr = f->lw;
  return GSMC_OK;
}

GSMC_API int gsmc_get_state(gsmc_handle f, int64_t t, double* host_dst, size_t n_values) {
  if (!f || !host_dst) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (t == 0) t = f->T;
  if (t < 1 || t > f->T) return fail(GSMC_E_BADARG, "time step %lld out of range 1..%lld", (long long)t, (long long)f->T);
  if (t != f->T && !f->cfg.keep_history) return fail(GSMC_E_BADARG, "earlier time steps need keep_history=1");
  if (n_values != (size_t)(f->D * f->n)) return fail(GSMC_E_BADARG, "expected a buffer of %lld values", (long long)(f->D * f->n));
  CK(cudaSetDevice(f->device));
  CKRC(ensure_f64(f, n_values));
  const int grid = (int)((f->n + GSMC_BLOCK - 1) / GSMC_BLOCK);
  {
    ProfScope ps(f, KC_OTHER);
    if (f->f32) get_state_kernel<float><<<grid, GSMC_BLOCK, 0, f->stream>>>(make_hist_view<float>(f), f->rank, f->n, t, f->T, f->pending ? 1 : 0, f->d_f64);
    else get_state_kernel<double><<<grid, GSMC_BLOCK, 0, f->stream>>>(make_hist_view<double>(f), f->rank, f->n, t, f->T, f->pending ? 1 : 0, f->d_f64);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host_dst, f->d_f64, n_values * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

GSMC_API int gsmc_get_observation(gsmc_handle f, int64_t t, double* host_dst, size_t n) {
  if (!f || !host_dst) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (t == 0) t = f->T;
  if (t < 1 || t > f->T) return fail(GSMC_E_BADARG, "time step %lld out of range 1..%lld", (long long)t, (long long)f->T);
  if ((int64_t)f->unobserved.size() < t || !f->unobserved[t - 1] || !f->obs_slab)
    return fail(GSMC_E_BADARG, "time step %lld was observed: its observation is the constraint the caller passed", (long long)t);
  if (t != f->T && !f->cfg.keep_history) return fail(GSMC_E_BADARG, "earlier time steps need keep_history=1");
  if (f->nranks > 1) return fail(GSMC_E_UNSUPPORTED, "sampled observation choices are read back on unsharded filters only");
  if (n != (size_t)f->n) return fail(GSMC_E_BADARG, "expected a buffer of %lld values", (long long)f->n);
  CK(cudaSetDevice(f->device));
  CKRC(ensure_f64(f, n));
  const int grid = (int)((f->n + GSMC_BLOCK - 1) / GSMC_BLOCK);
  {
    // the ancestor walk of get_state over the one-column observation slab
    ProfScope ps(f, KC_OTHER);
    if (f->f32) { HistView<float> h = make_hist_view<float>(f); h.slab[0] = (const float*)f->obs_slab; h.D = 1;
                  get_state_kernel<float><<<grid, GSMC_BLOCK, 0, f->stream>>>(h, f->rank, f->n, t, f->T, f->pending ? 1 : 0, f->d_f64); }
    else { HistView<double> h = make_hist_view<double>(f); h.slab[0] = (const double*)f->obs_slab; h.D = 1;
           get_state_kernel<double><<<grid, GSMC_BLOCK, 0, f->stream>>>(h, f->rank, f->n, t, f->T, f->pending ? 1 : 0, f->d_f64); }
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host_dst, f->d_f64, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

GSMC_API int gsmc_get_trajectories(gsmc_handle f, const int64_t* idx, size_t n_idx, double* out, size_t n_values) {
  if (!f || !idx || !out) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (!f->cfg.keep_history && f->T > 1) return fail(GSMC_E_BADARG, "trajectories need keep_history=1");
  if (n_values != n_idx * (size_t)f->T * f->D) return fail(GSMC_E_BADARG, "expected a buffer of %zu values", n_idx * (size_t)f->T * f->D);
  bool remote = false;
  for (size_t s = 0; s < n_idx; ++s) {
    if (idx[s] < 0 || idx[s] >= f->N) return fail(GSMC_E_BADARG, "particle index %lld out of range", (long long)idx[s]);
    if (idx[s] < f->first || idx[s] >= f->first + f->n) remote = true;
  }
  if (n_idx == 0) return GSMC_OK;
  CK(cudaSetDevice(f->device));
  CKRC(ensure_f64(f, n_values + n_idx));
  int64_t* d_idx = (int64_t*)(f->d_f64 + n_values);
  CK(cudaMemcpyAsync(d_idx, idx, n_idx * sizeof(int64_t), cudaMemcpyHostToDevice, f->stream));
  const int grid = (int)((n_idx + GSMC_BLOCK - 1) / GSMC_BLOCK);
  {
    ProfScope ps(f, KC_OTHER);
    if (f->f32) trajectories_kernel<float><<<grid, GSMC_BLOCK, 0, f->stream>>>(make_hist_view<float>(f), f->n, d_idx, (int64_t)n_idx, f->T, f->pending ? 1 : 0, f->d_f64);
    else trajectories_kernel<double><<<grid, GSMC_BLOCK, 0, f->stream>>>(make_hist_view<double>(f), f->n, d_idx, (int64_t)n_idx, f->T, f->pending ? 1 : 0, f->d_f64);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, f->d_f64, n_values * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  // rows of other ranks were read: the call is then collective (every rank asks for the same set, as after
  // gsmc_sample_unweighted) and nobody moves on -- and overwrites rows -- while a peer may still be reading them
  if (remote) CKRC(peer_barrier(f));
  return GSMC_OK;
}

GSMC_API int gsmc_get_ancestors(gsmc_handle f, int64_t* host_dst, size_t n) {
  if (!f || !host_dst) return fail(GSMC_E_BADARG, "null argument");
  if (n != (size_t)f->n) return fail(GSMC_E_BADARG, "expected a buffer of %lld values", (long long)f->n);
  CK(cudaSetDevice(f->device));
  if (f->last_resample_step == 0) {         // parents = collect(1:num_particles), particle_filter.jl:90,107
    for (size_t i = 0; i < n; ++i) host_dst[i] = f->first + (int64_t)i;
    return GSMC_OK;
  }
  if (f->cfg.keep_history ? false : (f->last_resample_step < f->T)) return fail(GSMC_E_BADARG, "ancestor column has been recycled (keep_history=0)");
  CKRC(ensure_f64(f, n));
  { ProfScope ps(f, KC_OTHER);
    anc_to_global_kernel<<<(int)((n + GSMC_BLOCK - 1) / GSMC_BLOCK), GSMC_BLOCK, 0, f->stream>>>(anc_col(f, f->anc_slab, f->last_resample_step), (int64_t)n, f->n, (int64_t*)f->d_f64); }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host_dst, f->d_f64, n * sizeof(int64_t), cudaMemcpyDeviceToHost, f->stream));
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}

static int sample_unweighted_impl(gsmc_filter* f, uint64_t num_samples, int64_t* idx_out, int phases) {
  if (!f || (!idx_out && num_samples)) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (num_samples == 0 && f->nranks == 1) return GSMC_OK;
  if (num_samples == 0) return fail(GSMC_E_BADARG, "sample_unweighted is collective on a sharded filter: every rank passes the same num_samples > 0");
  // sampling indexes the CURRENT particle order; with a pending resample that order only exists once the next step
  // has applied the ancestor column. Nothing is touched before this check.
  if (f->pending) return fail(GSMC_E_UNSUPPORTED, "sample_unweighted between maybe_resample and the next step");
  if (f->urep_n && f->urep_n != num_samples) return fail(GSMC_E_BADARG, "replay uniforms for sample_unweighted: expected %llu values", (unsigned long long)num_samples);
  CK(cudaSetDevice(f->device));
  CKRC(ensure_f64(f, num_samples));
  int64_t* d_idx = (int64_t*)f->d_f64;
  const int grid = (int)((num_samples + GSMC_BLOCK - 1) / GSMC_BLOCK);
  if (phases & PH_LOCAL) {
    // statistics (max) of the current weights, then the integer CDF, unconditionally
    CKRC(launch_finalize(f, -1.0));
    const double scale = weight_scale(f);
    { ProfScope ps(f, KC_SCAN);
      if (f->f32) weights_kernel<float, true, false><<<f->n_segs, GSMC_BLOCK, 0, f->stream>>>(
          (const float*)f->lw, f->n, scale, f->ds, f->cdf, f->raw0, 0, 0, 0, nullptr, nullptr, nullptr, f->n_tiles, f->seg_tiles, 0);
      else weights_kernel<double, true, false><<<f->n_segs, GSMC_BLOCK, 0, f->stream>>>(
          (const double*)f->lw, f->n, scale, f->ds, f->cdf, f->raw0, 0, 0, 0, nullptr, nullptr, nullptr, f->n_tiles, f->seg_tiles, 0); }
    CK(cudaGetLastError());
  }
  CKRC(launch_scan(f, KC_SCAN, f->raw0, nullptr, f->seg_a, nullptr, SCAN_Q, 0, (phases & PH_LOCAL) != 0, (phases & PH_GLOBAL) != 0));
  if (!(phases & PH_GLOBAL)) return GSMC_OK;
  const double* urep = f->urep_n ? f->d_urep : nullptr;
  { ProfScope ps(f, KC_SEARCH);
    search_iid_kernel<<<grid, GSMC_BLOCK, 0, f->stream>>>(make_cdf_view(f, false), f->ds, urep, f->cfg.seed, f->n_sample_calls, GSMC_STREAM_SAMPLE,
                                                         (int64_t)num_samples, 0, nullptr, d_idx, 0); }
  CK(cudaGetLastError());
  f->urep_n = 0;
  f->n_sample_calls += 1;
  CK(cudaMemcpyAsync(idx_out, d_idx, num_samples * sizeof(int64_t), cudaMemcpyDeviceToHost, f->stream));
  CKRC(fetch_scalars(f));
  if (!(f->h_ds->cdf_total > 0)) return fail(GSMC_E_DEGENERATE, "total weight is zero or not finite");
  return GSMC_OK;
}
GSMC_API int gsmc_sample_unweighted(gsmc_handle f, uint64_t num_samples, int64_t* idx_out) {
  if (f && f->group) return fail(GSMC_E_BADARG, "member of an emulated shard group: call gsmc_group_sample_unweighted");
  return sample_unweighted_impl(f, num_samples, idx_out, PH_ALL);
}

GSMC_API int gsmc_importance_sampling(const gsmc_config* cfg, const double* params, size_t n_params, const double* obs, size_t n_obs,
                                      int prop, const double* pp, size_t npp, double* lml_out, gsmc_handle* out) {
  if (!cfg || !out || !lml_out) return fail(GSMC_E_BADARG, "null argument");
  if (!model_is_importance(cfg->model_id)) return fail(GSMC_E_UNSUPPORTED, "model %d is a state-space family: compose init/step instead", cfg->model_id);
  gsmc_handle f = nullptr;
  CKRC(gsmc_create(cfg, params, n_params, &f));
  int rc = gsmc_init(f, obs, n_obs, prop, pp, npp);                       // importance.jl:25-28 / 41-47
  if (rc == GSMC_OK) rc = launch_finalize(f, -1.0);                       // log_total_weight = logsumexp(log_weights)
  if (rc == GSMC_OK) {
    ProfScope ps(f, KC_OTHER);
    const int grid = (int)((f->n + GSMC_BLOCK - 1) / GSMC_BLOCK);
    if (f->f32) normalize_lw_kernel<float><<<grid, GSMC_BLOCK, 0, f->stream>>>((float*)f->lw, f->n, f->ds, f->rank);
    else normalize_lw_kernel<double><<<grid, GSMC_BLOCK, 0, f->stream>>>((double*)f->lw, f->n, f->ds, f->rank);
  }
  if (rc == GSMC_OK) rc = fetch_scalars(f);
  if (rc != GSMC_OK) { std::string keep = g_last_error; gsmc_destroy(f); g_last_error = keep; return rc; }
  *lml_out = f->h_ds->log_total - gm_log((double)f->N);                   // importance.jl:30,49
  f->is_lml = *lml_out;
  f->stats_fresh = false;                   // the triple now describes the normalised weights: recomputed on demand
  *out = f;
  return GSMC_OK;
}

// The loop of gsmc_run_steps, enqueued on f->stream. handles (captured run only): one conditional handle per step; the
// resampling kernels of step s are then captured into the body of an IF node instead of being launched in their
// early-exit form.
static int enqueue_steps(gsmc_filter* f, const double* obs, size_t n_steps, size_t n_obs, int prop, const double* pp, size_t npp,
                         double ess_threshold, const std::vector<cudaGraphConditionalHandle>* handles, std::vector<cudaGraph_t>* bodies) {
  bool decided = false;
  for (size_t s = 0; s < n_steps; ++s) {
    if (!decided) {
      f->cond_finalize = handles ? (*handles)[s] : 0;
      const int rc = launch_finalize(f, ess_threshold);
      f->cond_finalize = 0;
      if (rc != GSMC_OK) return rc;
    }
    if (handles) {
      // IF node on this step's handle, depending on everything captured so far; its body = the resampling kernels
      cudaStreamCaptureStatus status;
      cudaGraph_t graph = nullptr;
      const cudaGraphNode_t* deps = nullptr;
      size_t n_deps = 0;
      CK(cudaStreamGetCaptureInfo_v2(f->stream, &status, nullptr, &graph, &deps, &n_deps));
      cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = (*handles)[s];
      np.conditional.type = cudaGraphCondTypeIf;
      np.conditional.size = 1;
      cudaGraphNode_t node;
      CK(cudaGraphAddNode(&node, graph, deps, n_deps, &np));
      cudaGraph_t body = np.conditional.phGraph_out[0];
      cudaStream_t main_stream = f->stream;
      CK(cudaStreamBeginCaptureToGraph(f->body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
      f->stream = f->body_stream;
      g_pdl_suppress_next = true;
      int rc = launch_resample(f, 0, false);
      f->stream = main_stream;
      cudaGraph_t out = nullptr;
      const cudaError_t e = cudaStreamEndCapture(f->body_stream, &out);
      if (rc != GSMC_OK) return rc;
      CK(e);
      bodies->push_back(body);
      CK(cudaStreamUpdateCaptureDependencies(f->stream, &node, 1, cudaStreamSetCaptureDependencies));
      g_pdl_suppress_next = true;                        // the propagate below follows the conditional node
    } else {
      CKRC(launch_resample(f, 1, false));
    }
    f->fuse_next_decide = (f->nranks == 1 || xmode(f) == XMODE_LL) && !f->profiling && s + 1 < n_steps;
    f->fuse_thr = ess_threshold;
    f->cond_next = (handles && s + 1 < n_steps) ? (*handles)[s + 1] : 0;
    const int rc = launch_propagate(f, false, obs + s * n_obs, n_obs, prop, pp, npp, true);
    g_pdl_suppress_next = false;
    decided = f->fuse_next_decide;
    f->fuse_next_decide = false;
    f->cond_next = 0;
    if (rc != GSMC_OK) return rc;
    f->T += 1;
  }
  return GSMC_OK;
}

static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 0x100000001b3ULL; }
  return h;
}
// A run shape that repeats (same first step, number of steps, observations, proposal and threshold on this handle --
// e.g. reset + init + run_steps in a loop) is captured into ONE CUDA graph on its second occurrence and replayed from
// then on: every step's resampling kernels sit in the body of a conditional (IF) node whose handle the deciding
// kernel sets on the device, so the steps that do not resample launch nothing, and kernel-to-kernel gaps shrink to
// graph edges. Returns 1 if the steps were enqueued through the graph, 0 if the caller should enqueue them itself.
static int run_steps_graph(gsmc_filter* f, const double* obs, size_t n_steps, size_t n_obs, int prop, const double* pp, size_t npp,
                           double ess_threshold, int* used) {
  *used = 0;
  // Measured on B200 (profiles/r2_graph.txt): a conditional node costs about as much as the three early-exit launches of
  // the multinomial scheme (cfg 3: 22.9 ms with the graph, 22.6 without), but far less than the ten of the residual
  // scheme (cfg 4: 98 ms against 130 ms). Default: residual only; GSMC_GRAPH=1 captures every scheme, GSMC_NO_GRAPH=1 none.
  static int env_off = getenv("GSMC_NO_GRAPH") ? 1 : 0;
  static int env_all = getenv("GSMC_GRAPH") ? 1 : 0;
  const bool residual = f->cfg.resample_scheme == GSMC_RESAMPLE_RESIDUAL;
  if (env_off || (!residual && !env_all) || f->graph_disabled || f->profiling || f->group || (f->nranks > 1 && xmode(f) != XMODE_LL) || n_steps < 2 ||
      model_obs_on_device(f->model) || false) return GSMC_OK;
  uint64_t key = 0xcbf29ce484222325ULL;
  const int64_t t0 = f->T;
  key = fnv1a(key, &t0, sizeof t0); key = fnv1a(key, &n_steps, sizeof n_steps); key = fnv1a(key, &n_obs, sizeof n_obs);
  key = fnv1a(key, &prop, sizeof prop); key = fnv1a(key, pp, npp * sizeof(double)); key = fnv1a(key, &ess_threshold, sizeof ess_threshold);
  key = fnv1a(key, obs, n_steps * n_obs * sizeof(double));
  if (key == 0) key = 1;
  if (f->graph_exec && f->graph_key == key) {
    CK(cudaGraphLaunch(f->graph_exec, f->stream));
    f->T += (int64_t)n_steps;
    f->graph_launches += 1; f->launches += f->graph_steps;
    *used = 1;
    return GSMC_OK;
  }
  if (f->graph_candidate != key) { f->graph_candidate = key; return GSMC_OK; }       // first occurrence: plain launches
  // second occurrence: capture
  drop_graph(f);
  if (!f->body_stream && cudaStreamCreateWithFlags(&f->body_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); f->graph_disabled = true; return GSMC_OK; }
  const int64_t launches0 = f->launches;
  std::vector<cudaGraphConditionalHandle> handles(n_steps);
  std::vector<cudaGraph_t> bodies;
  cudaGraph_t graph = nullptr;
  bool ok = cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    cudaStreamCaptureStatus status;
    cudaGraph_t g0 = nullptr;
    ok = cudaStreamGetCaptureInfo_v2(f->stream, &status, nullptr, &g0, nullptr, nullptr) == cudaSuccess && g0;
    for (size_t s = 0; ok && s < n_steps; ++s) ok = cudaGraphConditionalHandleCreate(&handles[s], g0, 0, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) ok = enqueue_steps(f, obs, n_steps, n_obs, prop, pp, npp, ess_threshold, &handles, &bodies) == GSMC_OK;
    g_pdl_suppress_next = false;
    cudaStreamCaptureStatus st2 = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(f->body_stream, &st2);
    if (st2 != cudaStreamCaptureStatusNone) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(f->body_stream, &junk); }
    const cudaError_t e = cudaStreamEndCapture(f->stream, &graph);
    ok = ok && e == cudaSuccess && graph;
  }
  f->T = t0;                                              // nothing has run yet
  f->zrep_n = 0; f->urep_n = 0;
  if (ok) ok = cudaGraphInstantiate(&f->graph_exec, graph, 0) == cudaSuccess;
  if (!ok) {
    const cudaError_t e = cudaGetLastError();
    if (getenv("GSMC_GRAPH_DEBUG")) fprintf(stderr, "libgensmc: graph capture of gsmc_run_steps failed (%s); using stream launches\n", cudaGetErrorString(e));
    if (graph) cudaGraphDestroy(graph);
    f->graph_exec = nullptr; f->graph_disabled = true; f->launches = launches0;
    return GSMC_OK;
  }
  f->graph = graph; f->graph_key = key; f->graph_steps = f->launches - launches0; f->launches = launches0;
  CK(cudaGraphLaunch(f->graph_exec, f->stream));
  f->T += (int64_t)n_steps;
  f->graph_launches += 1; f->launches += f->graph_steps;
  *used = 1;
  return GSMC_OK;
}

GSMC_API int gsmc_run_steps(gsmc_handle f, const double* obs, size_t n_steps, size_t n_obs, int prop, const double* pp, size_t npp,
                            double ess_threshold) {
  if (!f || !obs) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (f->is_importance) return fail(GSMC_E_BADARG, "importance-sampling models have a single step");
  if (f->pending) return fail(GSMC_E_BADARG, "call gsmc_step after gsmc_maybe_resample before gsmc_run_steps");
  if (f->zrep_n || f->urep_n) return fail(GSMC_E_BADARG, "replay draws are consumed by the per-call API only");
  if (!(ess_threshold >= 0.0)) return fail(GSMC_E_BADARG, "ess_threshold must be >= 0");
  if (f->cfg.keep_history && f->T + (int64_t)n_steps > f->cap)
    return fail(GSMC_E_BADARG, "history_capacity (%lld steps) would be exceeded by %zu more steps after step %lld", (long long)f->cap, n_steps, (long long)f->T);
  if (n_obs != (size_t)expected_obs(f)) return fail(GSMC_E_BADARG, "model %d needs %d observation value(s) per step, got %zu", f->model, expected_obs(f), n_obs);
  if (f->group) return fail(GSMC_E_UNSUPPORTED, "gsmc_run_steps is not available on an emulated shard group");
  CK(cudaSetDevice(f->device));
  // The threshold of every decision is known here, so the last block of each propagate also takes the decision of the
  // next step (on a sharded filter after exchanging the ranks' triples over the LL mailboxes): only the first step of
  // the call needs a finalize launch.
  {
    // validate the arguments of every step before anything is enqueued or captured
    ModelArgs a;
    for (size_t s = 0; s < n_steps; ++s) if (!model_obs_on_device(f->model)) CKRC(fill_model_args(f, a, obs + s * n_obs, n_obs, prop, pp, npp));
  }
  int used = 0;
  CKRC(run_steps_graph(f, obs, n_steps, n_obs, prop, pp, npp, ess_threshold, &used));
  if (!used) CKRC(enqueue_steps(f, obs, n_steps, n_obs, prop, pp, npp, ess_threshold, nullptr, nullptr));
  f->decided_since_step = false; f->pending = false; f->stats_fresh = false;
  // surface a degenerate-weight error recorded on the device
  CKRC(fetch_scalars(f));
  f->last_resample_step = f->h_ds->last_resample_step;          // tracked on the device while the host was not looking
  if (f->h_ds->error) return device_error(f, "during gsmc_run_steps");
  return GSMC_OK;
}

// ------------------------------------------------------------------------------------------------
// shard emulation: R ranks of one sharded filter on one device (see gsmc_group_s)
// ------------------------------------------------------------------------------------------------
GSMC_API int gsmc_group_create(int nranks, int device, gsmc_group* out) {
  if (!out) return fail(GSMC_E_BADARG, "null argument");
  if (nranks < 2 || nranks > GSMC_MAX_RANKS) return fail(GSMC_E_BADARG, "2 <= nranks <= %d", GSMC_MAX_RANKS);
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  gsmc_group_s* g = new gsmc_group_s();
  g->nranks = nranks; g->device = device;
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { delete g; return fail(GSMC_E_CUDA, "stream creation failed"); }
  *out = g;
  return GSMC_OK;
}
GSMC_API void gsmc_group_destroy(gsmc_group g) {
  if (!g) return;
  cudaSetDevice(g->device);
  for (int r = 0; r < g->nranks; ++r) if (g->member[r]) gsmc_destroy(g->member[r]);      // (group stays set: no peer barrier kernel on one stream)
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}
GSMC_API int gsmc_group_attach(gsmc_group g, int rank, gsmc_handle f) {
  if (!g || !f) return fail(GSMC_E_BADARG, "null argument");
  if (rank < 0 || rank >= g->nranks || g->member[rank]) return fail(GSMC_E_BADARG, "rank %d out of range or already attached", rank);
  if (f->T != 0 || f->state_slab) return fail(GSMC_E_BADARG, "attach must precede gsmc_init");
  if (f->device != g->device) return fail(GSMC_E_BADARG, "group lives on device %d, filter on device %d", g->device, f->device);
  if (f->N % ((int64_t)g->nranks * GSMC_PAD) != 0) return fail(GSMC_E_BADARG, "num_particles must be a multiple of %d * nranks", GSMC_PAD);
  CK(cudaSetDevice(f->device));
  if (f->own_stream && f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
  f->stream = g->stream; f->own_stream = false;
  f->rank = rank; f->nranks = g->nranks; f->group = g;
  f->n = f->N / g->nranks; f->first = f->n * rank;
  CKRC(alloc_buffers(f));
  g->member[rank] = f;
  for (int a = 0; a < g->nranks; ++a) for (int b = 0; b < g->nranks; ++b) {
    gsmc_filter *x = g->member[a], *y = g->member[b];
    if (!x || !y) continue;
    x->peer_slab[b] = y->state_slab; x->peer_anc[b] = y->anc_slab; x->peer_cdf[b] = y->cdf; x->peer_ds[b] = y->ds;
  }
  return GSMC_OK;
}
static int group_ready(gsmc_group g) {
  if (!g) return fail(GSMC_E_BADARG, "null group");
  for (int r = 0; r < g->nranks; ++r) if (!g->member[r]) return fail(GSMC_E_BADARG, "rank %d of the group has no filter attached", r);
  for (int r = 1; r < g->nranks; ++r)
    if (g->member[r]->T != g->member[0]->T || g->member[r]->pending != g->member[0]->pending) return fail(GSMC_E_BADARG, "the ranks of the group are out of step");
  return GSMC_OK;
}
GSMC_API int gsmc_group_init(gsmc_group g, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp) {
  CKRC(group_ready(g));
  for (int r = 0; r < g->nranks; ++r) CKRC(gsmc_init(g->member[r], obs, n_obs, prop, pp, npp));
  return GSMC_OK;
}
GSMC_API int gsmc_group_step(gsmc_group g, const double* obs, size_t n_obs, int prop, const double* pp, size_t npp) {
  CKRC(group_ready(g));
  for (int r = 0; r < g->nranks; ++r) CKRC(gsmc_step(g->member[r], obs, n_obs, prop, pp, npp));
  return GSMC_OK;
}
GSMC_API int gsmc_group_maybe_resample(gsmc_group g, double ess_threshold, int* did_resample, double* ess_out) {
  CKRC(group_ready(g));
  gsmc_filter* f0 = g->member[0];
  if (f0->T < 1) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (!(ess_threshold >= 0.0)) return fail(GSMC_E_BADARG, "ess_threshold must be >= 0");
  CK(cudaSetDevice(g->device));
  if (f0->pending) {
    if ((double)f0->N < ess_threshold) return fail(GSMC_E_UNSUPPORTED, "two resampling events without a step in between");
    if (did_resample) *did_resample = 0;
    if (ess_out) *ess_out = (double)f0->N;
    return GSMC_OK;
  }
  const bool replay = f0->urep_n > 0;
  const int R = g->nranks;
  for (int r = 0; r < R; ++r) CKRC(launch_finalize(g->member[r], ess_threshold, true));
  if (!replay) {
    for (int ph = PH_LOCAL; ph <= PH_GLOBAL; ph <<= 1)
      for (int r = 0; r < R; ++r) CKRC(launch_resample(g->member[r], 1, false, ph));
  }
  int did = 0;
  for (int r = 0; r < R; ++r) {
    gsmc_filter* f = g->member[r];
    CKRC(wait_decision(f));
    f->stats_fresh = true; f->decided_since_step = true;
    if (f->h_ds->error) { f->decided_since_step = false; return device_error(f, "in maybe_resample"); }
    if (r == 0) did = f->h_ds->do_resample;
    else if (f->h_ds->do_resample != did) return fail(GSMC_E_PEER, "the ranks of the group took different decisions");
  }
  if (did) {
    if (replay) {
      for (int ph = PH_LOCAL; ph <= PH_GLOBAL; ph <<= 1)
        for (int r = 0; r < R; ++r) CKRC(launch_resample(g->member[r], 0, true, ph));
    }
    for (int r = 0; r < R; ++r) { g->member[r]->pending = true; g->member[r]->last_resample_step = g->member[r]->T + 1; }
  }
  for (int r = 0; r < R; ++r) g->member[r]->urep_n = 0;
  if (did_resample) *did_resample = did;
  if (ess_out) *ess_out = f0->h_ds->ess;
  return GSMC_OK;
}
GSMC_API int gsmc_group_sample_unweighted(gsmc_group g, uint64_t num_samples, int64_t* idx_out) {
  CKRC(group_ready(g));
  std::vector<int64_t> scratch(num_samples);
  for (int r = 0; r < g->nranks; ++r) CKRC(sample_unweighted_impl(g->member[r], num_samples, scratch.data(), PH_LOCAL));
  for (int r = 0; r < g->nranks; ++r) CKRC(sample_unweighted_impl(g->member[r], num_samples, r == 0 ? idx_out : scratch.data(), PH_GLOBAL));
  return GSMC_OK;
}

// ------------------------------------------------------------------------------------------------
// checkpoint / resume (the reference's analogue is saving Julia objects with JLD,
// examples/planning/filtering.jl:822-829). One file per handle (per rank of a sharded filter): header, model
// parameters, device scalars, the resample flags, the log weights and the state / ancestor columns that exist.
// ------------------------------------------------------------------------------------------------
struct CkptHeader {
  char magic[8];                 // "GSMCCKP1"
  uint32_t header_bytes, dev_scalars_bytes;
  gsmc_config cfg;               // stream pointer zeroed
  int32_t rank, nranks;
  int64_t N, n, n_pad, cap, flag_mod, T, last_resample_step;
  int32_t D, pending, decided_since_step, is_importance;
  uint32_t n_sample_calls, n_params;
  double is_lml;
  int64_t n_cols;                // state / ancestor columns stored (steps T-n_cols+1 .. T)
};
static int write_dev(FILE* fp, gsmc_filter* f, const void* dev, size_t bytes, std::vector<char>& host) {
  const size_t chunk = (size_t)64 << 20;
  if (host.size() < (bytes < chunk ? bytes : chunk)) host.resize(bytes < chunk ? bytes : chunk);
  for (size_t off = 0; off < bytes; off += chunk) {
    const size_t len = bytes - off < chunk ? bytes - off : chunk;
    CK(cudaMemcpyAsync(host.data(), (const char*)dev + off, len, cudaMemcpyDeviceToHost, f->stream));
    CK(cudaStreamSynchronize(f->stream));
    if (fwrite(host.data(), 1, len, fp) != len) return fail(GSMC_E_BADARG, "short write to the checkpoint file");
  }
  return GSMC_OK;
}
static int read_dev(FILE* fp, gsmc_filter* f, void* dev, size_t bytes, std::vector<char>& host) {
  const size_t chunk = (size_t)64 << 20;
  if (host.size() < (bytes < chunk ? bytes : chunk)) host.resize(bytes < chunk ? bytes : chunk);
  for (size_t off = 0; off < bytes; off += chunk) {
    const size_t len = bytes - off < chunk ? bytes - off : chunk;
    if (fread(host.data(), 1, len, fp) != len) return fail(GSMC_E_BADARG, "checkpoint file is truncated");
    CK(cudaMemcpyAsync((char*)dev + off, host.data(), len, cudaMemcpyHostToDevice, f->stream));
    CK(cudaStreamSynchronize(f->stream));
  }
  return GSMC_OK;
}
GSMC_API int gsmc_save(gsmc_handle f, const char* path) {
  if (!f || !path) return fail(GSMC_E_BADARG, "null argument");
  if (f->T < 1 || !f->state_slab) return fail(GSMC_E_BADARG, "filter is not initialised");
  if (f->obs_slab) return fail(GSMC_E_UNSUPPORTED, "checkpoints of filters with unobserved steps (sampled observation choices) are not supported");
  CK(cudaSetDevice(f->device));
  CK(cudaStreamSynchronize(f->stream));
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail(GSMC_E_BADARG, "cannot open %s for writing", path);
  CkptHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "GSMCCKP1", 8);
  h.header_bytes = (uint32_t)sizeof h; h.dev_scalars_bytes = (uint32_t)offsetof(DevScalars, mbox);
  h.cfg = f->cfg; h.cfg.stream = nullptr;
  h.rank = f->rank; h.nranks = f->nranks; h.N = f->N; h.n = f->n; h.n_pad = f->n_pad; h.cap = f->cap; h.flag_mod = f->flag_mod;
  h.T = f->T; h.last_resample_step = f->last_resample_step; h.D = f->D;
  h.pending = f->pending; h.decided_since_step = f->decided_since_step; h.is_importance = f->is_importance;
  h.n_sample_calls = f->n_sample_calls; h.n_params = (uint32_t)f->params.size(); h.is_lml = f->is_lml;
  // keep_history: every step so far; otherwise the current column and, with a resample pending, the column it fills
  h.n_cols = f->cfg.keep_history ? f->T : 1;
  std::vector<char> host;
  int rc = GSMC_OK;
  if (fwrite(&h, sizeof h, 1, fp) != 1 || fwrite(f->params.data(), sizeof(double), f->params.size(), fp) != f->params.size())
    rc = fail(GSMC_E_BADARG, "short write to the checkpoint file");
  if (rc == GSMC_OK) rc = write_dev(fp, f, f->ds, offsetof(DevScalars, mbox), host);
  if (rc == GSMC_OK) rc = write_dev(fp, f, f->resampled, (size_t)f->flag_mod * sizeof(int), host);
  if (rc == GSMC_OK) rc = write_dev(fp, f, f->lw, f->n_pad * real_size(f), host);
  for (int64_t t = f->T - h.n_cols + 1; t <= f->T && rc == GSMC_OK; ++t) {
    rc = write_dev(fp, f, state_col(f, f->state_slab, t), (size_t)f->D * f->n_pad * real_size(f), host);
    if (rc == GSMC_OK) rc = write_dev(fp, f, anc_col(f, f->anc_slab, t), (size_t)f->n_pad * sizeof(uint32_t), host);
  }
  if (rc == GSMC_OK && f->pending) rc = write_dev(fp, f, anc_col(f, f->anc_slab, f->T + 1), (size_t)f->n_pad * sizeof(uint32_t), host);
  if (fclose(fp) != 0 && rc == GSMC_OK) rc = fail(GSMC_E_BADARG, "closing %s failed", path);
  return rc;
}
GSMC_API int gsmc_restore(gsmc_handle f, const char* path) {
  if (!f || !path) return fail(GSMC_E_BADARG, "null argument");
  CK(cudaSetDevice(f->device));
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(GSMC_E_BADARG, "cannot open %s", path);
  CkptHeader h;
  std::vector<double> params;
  int rc = GSMC_OK;
  if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "GSMCCKP1", 8) != 0 || h.header_bytes != sizeof h || h.dev_scalars_bytes != offsetof(DevScalars, mbox))
    rc = fail(GSMC_E_BADARG, "%s is not a checkpoint of this library version", path);
  if (rc == GSMC_OK) {
    params.resize(h.n_params);
    if (fread(params.data(), sizeof(double), params.size(), fp) != params.size()) rc = fail(GSMC_E_BADARG, "checkpoint file is truncated");
  }
  if (rc == GSMC_OK && (h.cfg.model_id != f->cfg.model_id || h.cfg.dtype != f->cfg.dtype || h.cfg.resample_scheme != f->cfg.resample_scheme ||
                        h.cfg.num_particles != f->cfg.num_particles || h.cfg.seed != f->cfg.seed || h.cfg.keep_history != f->cfg.keep_history ||
                        h.rank != f->rank || h.nranks != f->nranks || params != f->params))
    rc = fail(GSMC_E_BADARG, "the checkpoint was written by a filter with another configuration (model, dtype, particles, seed, history, parameters or sharding)");
  if (rc == GSMC_OK && !f->state_slab) rc = alloc_buffers(f);
  if (rc == GSMC_OK && (h.cap > f->cap || h.n_pad != f->n_pad || h.flag_mod != f->flag_mod))
    rc = fail(GSMC_E_BADARG, "the checkpoint needs history_capacity %lld, the handle has %lld", (long long)h.cap, (long long)f->cap);
  std::vector<char> host;
  if (rc == GSMC_OK) rc = peer_barrier(f);
  if (rc == GSMC_OK) rc = read_dev(fp, f, f->ds, offsetof(DevScalars, xseq), host);       // the exchange sequence number and mailboxes stay
  if (rc == GSMC_OK && fseek(fp, (long)(offsetof(DevScalars, mbox) - offsetof(DevScalars, xseq)), SEEK_CUR) != 0) rc = fail(GSMC_E_BADARG, "checkpoint file is truncated");
  if (rc == GSMC_OK) rc = read_dev(fp, f, f->resampled, (size_t)f->flag_mod * sizeof(int), host);
  if (rc == GSMC_OK) rc = read_dev(fp, f, f->lw, f->n_pad * real_size(f), host);
  if (rc == GSMC_OK) {
    f->T = h.T;                                         // the column helpers index by step
    for (int64_t t = h.T - h.n_cols + 1; t <= h.T && rc == GSMC_OK; ++t) {
      rc = read_dev(fp, f, state_col(f, f->state_slab, t), (size_t)f->D * f->n_pad * real_size(f), host);
      if (rc == GSMC_OK) rc = read_dev(fp, f, anc_col(f, f->anc_slab, t), (size_t)f->n_pad * sizeof(uint32_t), host);
    }
    if (rc == GSMC_OK && h.pending) rc = read_dev(fp, f, anc_col(f, f->anc_slab, h.T + 1), (size_t)f->n_pad * sizeof(uint32_t), host);
  }
  fclose(fp);
  if (rc != GSMC_OK) return rc;
  f->last_resample_step = h.last_resample_step; f->pending = h.pending != 0; f->decided_since_step = h.decided_since_step != 0;
  f->is_importance = h.is_importance != 0; f->n_sample_calls = h.n_sample_calls; f->is_lml = h.is_lml;
  f->stats_fresh = false; f->zrep_n = 0; f->urep_n = 0;
  return GSMC_OK;
}

GSMC_API int gsmc_register_model_plugin(const char* path, int* model_id_out) {
  if (!path || !model_id_out) return fail(GSMC_E_BADARG, "null argument");
  for (size_t k = 0; k < g_plugins.size(); ++k) if (g_plugins[k].path == path) { *model_id_out = GSMC_MODEL_PLUGIN_BASE + (int)k; return GSMC_OK; }
  void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return fail(GSMC_E_BADARG, "cannot load model plugin %s: %s", path, dlerror());
  ModelPlugin pl;
  pl.lib = lib; pl.path = path;
  gsmc_plugin_describe_fn describe = (gsmc_plugin_describe_fn)dlsym(lib, "gsmc_plugin_describe");
  pl.propagate = (gsmc_plugin_propagate_fn)dlsym(lib, "gsmc_plugin_propagate");
  pl.sample_obs = (gsmc_plugin_sample_obs_fn)dlsym(lib, "gsmc_plugin_sample_obs");
  memset(&pl.info, 0, sizeof pl.info);
  if (!describe || !pl.propagate || describe(&pl.info) != 0) { dlclose(lib); return fail(GSMC_E_BADARG, "%s is not a libgensmc model plugin", path); }
  if (pl.info.abi != GSMC_PLUGIN_ABI || pl.info.sizeof_prop_args != sizeof(PropArgs<double>) || pl.info.sizeof_model_args != sizeof(ModelArgs) ||
      pl.info.sizeof_dev_scalars != sizeof(DevScalars)) {
    dlclose(lib);
    return fail(GSMC_E_BADARG, "model plugin %s was generated for another version of libgensmc (regenerate it)", path);
  }
  if (pl.info.D < 1 || pl.info.D > 16 || pl.info.n_params < 0 || pl.info.n_params > GSMC_MAX_INLINE_PARAMS) { dlclose(lib); return fail(GSMC_E_BADARG, "model plugin %s: unsupported shape", path); }
  g_plugins.push_back(pl);
  *model_id_out = GSMC_MODEL_PLUGIN_BASE + (int)g_plugins.size() - 1;
  return GSMC_OK;
}

GSMC_API int gsmc_trim(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  int dev = 0; cudaGetDevice(&dev);
  for (auto& kv : g_pool) { cudaSetDevice(kv.first.device); cudaFree(kv.second); }
  g_pool.clear();
  for (DevScalars* p : g_pinned_pool) cudaFreeHost(p);
  g_pinned_pool.clear();
  for (auto& kv : g_ipc_cache) { cudaSetDevice(kv.first.device); cudaIpcCloseMemHandle(kv.second); }
  g_ipc_cache.clear();
  cudaSetDevice(dev);
  return GSMC_OK;
}

GSMC_API int gsmc_local_count(gsmc_handle f, uint64_t* n_local, uint64_t* first_global) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  if (n_local) *n_local = (uint64_t)f->n;
  if (first_global) *first_global = (uint64_t)f->first;
  return GSMC_OK;
}
GSMC_API int gsmc_state_dim(gsmc_handle f, int* dim) {
  if (!f || !dim) return fail(GSMC_E_BADARG, "null argument");
  *dim = f->D;
  return GSMC_OK;
}
GSMC_API int gsmc_synchronize(gsmc_handle f) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  CK(cudaSetDevice(f->device));
  CK(cudaStreamSynchronize(f->stream));
  return GSMC_OK;
}
GSMC_API int gsmc_get_stats(gsmc_handle f, gsmc_stats* out) {
  if (!f || !out) return fail(GSMC_E_BADARG, "null argument");
  CK(cudaSetDevice(f->device));
  memset(out, 0, sizeof *out);
  if (f->ds) CKRC(fetch_scalars(f));
  harvest_profile(f);
  if (f->h_ds) {
    out->last_ess = f->h_ds->ess; out->last_log_total = f->h_ds->log_total; out->log_ml_est = f->h_ds->log_ml_est;
    out->num_resamples = f->h_ds->n_resamples;
  }
  out->num_steps = f->T;
  out->kernel_launches = f->launches;
  out->ms_propagate = f->prof_ms[KC_PROPAGATE]; out->n_propagate = f->prof_n[KC_PROPAGATE];
  out->ms_finalize = f->prof_ms[KC_FINALIZE]; out->n_finalize = f->prof_n[KC_FINALIZE];
  out->ms_scan = f->prof_ms[KC_SCAN]; out->n_scan = f->prof_n[KC_SCAN];
  out->ms_spacings = f->prof_ms[KC_SPACINGS]; out->n_spacings = f->prof_n[KC_SPACINGS];
  out->ms_search = f->prof_ms[KC_SEARCH]; out->n_search = f->prof_n[KC_SEARCH];
  out->ms_other = f->prof_ms[KC_OTHER]; out->n_other = f->prof_n[KC_OTHER];
  out->ms_propagate_gather = f->prof_ms[KC_PROPAGATE_GATHER]; out->n_propagate_gather = f->prof_n[KC_PROPAGATE_GATHER];
  out->graph_replays = f->graph_launches;
  return GSMC_OK;
}
GSMC_API int gsmc_set_profiling(gsmc_handle f, int enabled) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  harvest_profile(f);
  f->profiling = enabled != 0;
  if (enabled) { for (int c = 0; c < KC_COUNT; ++c) { f->prof_ms[c] = 0; f->prof_n[c] = 0; } }
  return GSMC_OK;
}
GSMC_API int gsmc_timer_start(gsmc_handle f) {
  if (!f) return fail(GSMC_E_BADARG, "null handle");
  CK(cudaSetDevice(f->device));
  CK(cudaEventRecord(f->timer_a, f->stream));
  return GSMC_OK;
}
GSMC_API int gsmc_timer_stop(gsmc_handle f, double* elapsed_ms) {
  if (!f || !elapsed_ms) return fail(GSMC_E_BADARG, "null argument");
  CK(cudaSetDevice(f->device));
  CK(cudaEventRecord(f->timer_b, f->stream));
  CK(cudaEventSynchronize(f->timer_b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, f->timer_a, f->timer_b));
  *elapsed_ms = ms;
  return GSMC_OK;
}

}  // extern "C"
