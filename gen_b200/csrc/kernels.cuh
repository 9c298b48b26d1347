// kernels.cuh -- hand-written sm_100a kernels of the SMC / importance-sampling hot path.
//
// All kernels are streaming passes over structure-of-arrays particle columns (no tensor cores: nothing here is a
// dense contraction). In propagate one thread owns QUADS of neighbouring particles: every column access is a 32-byte
// vector access (LDG/STG.256), a warp touches 1 KB contiguous per column, and one Philox call feeds the four particles of
// a quad; the resampling kernels work on pairs / octets with 16- and 32-byte accesses. The hot
// kernels are persistent (one resident wave of blocks looping over 2048-particle tiles) and are launched with
// programmatic dependent launch (pdl_wait / pdl_trigger below).
//
//   propagate_kernel   particle_filter.jl:84-88,103-105 (init), :143-146,165-172 (step) fused with the ancestor gather
//                      of :202-205 and with logsumexp/ESS: every thread carries a running (max, sum e, sum e^2) over its
//                      particles, the block combines them once, the last block to finish reduces the block partials into
//                      this rank's triple and (gsmc_run_steps) takes the next maybe_resample! decision itself
//   finalize_kernel    inference.jl:3-6 + particle_filter.jl:3-12 (normalize_weights, ESS): merge of the ranks' triples
//                      (fused NVLink mailbox exchange), the `ess < ess_threshold` decision and
//                      `log_ml_est += log_total - log N` of :194,201; the Bool goes to the host through a pinned mirror
//   weights_kernel     `weights = exp.(lnw)` + the CDF behind Categorical(weights/sum(weights)), :199-200, in 64-bit fixed
//                      point so the prefix sum is associative (order/shard independent), as a two-level CDF; fused: the
//                      Gamma gaps of the groups of sorted draws (the N iid draws of :200 are generated as grouped order
//                      statistics, so that search + gather stream through memory)
//   partition_kernel   scan of the segment totals (+ exchange of the ranks' totals) and the CDF window of every tile
//   search_sorted_kernel  parents[i] (:200): TMA-staged CDF window, per-group bracket, integer keys, binary searches over
//                      4-byte keys in shared memory
//   search_iid_kernel  replay / sample_unweighted_traces (:62-70); resid_* / det_copies: the residual scheme
#ifndef GSMC_KERNELS_CUH
#define GSMC_KERNELS_CUH

#include <stdint.h>
#include <cuda_runtime.h>
#include "gsmc_math.h"
#include "gsmc_rng.cuh"
#include "gsmc_fixed.h"
#include "models.cuh"

#define GSMC_BLOCK 256
#ifndef GSMC_PROP_QUADS
#define GSMC_PROP_QUADS 2
#endif
#ifndef GSMC_PROP_OCC
#define GSMC_PROP_OCC 3
#endif
#ifndef GSMC_PROP_OCC_WIDE
#define GSMC_PROP_OCC_WIDE 3
#endif
#define GSMC_TILE 2048            // particles (or thresholds) per block iteration of the scan / search kernels
#define GSMC_TILE_SHIFT 11
#define GSMC_PAD 2048             // local columns are padded to this many particles
#define GSMC_MAX_RANKS 8
#define GSMC_MAX_SEGS 1024          // segments (blocks of the streaming resample pass) per rank
#define GSMC_ANC_RANK_SHIFT 28    // ancestor word = (owner rank << 28) | local index
#define GSMC_ANC_INDEX_MASK 0x0fffffffu

// Programmatic dependent launch: kernels launched with the programmatic-stream-serialization attribute may start
// (block scheduling, shared-memory table staging) while the previous kernel of the stream drains; pdl_wait()
// blocks until that kernel has completed and its writes are visible, so it precedes the first access of any
// buffer another kernel writes or reads. pdl_trigger() lets the NEXT kernel's blocks be scheduled early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct LseTriple { double m, s1, s2; };   // max, sum exp(lw-m), sum exp(2(lw-m))

// Device-resident scalars: the filter never needs a host round trip to take a decision.
struct DevScalars {
  double max_lw, log_total, ess;
  double log_ml_est;
  int do_resample;          // decision of the last finalize(decide)
  int error;                // sticky: 1 = total weight zero / not finite at a resample
  uint32_t n_resamples;     // resampling events so far
  uint32_t rho;             // Philox event index of the resample being executed
  unsigned int blocks_done; // propagate: blocks that have published their logsumexp partial (the last one reduces them)
  unsigned int host_token;  // pinned host mirror only: token of the last decision the device has written there
  uint64_t cdf_total;       // C_N over all ranks
  uint64_t gap_total;       // S_tot = head + all group gaps of the event (fixed point, scale 2^20)
  uint64_t gap_head;        // A_0: the Exp(1) gap below the first order statistic of the event
  uint64_t n_det;           // residual scheme: number of deterministic copies
  uint64_t n_draws;         // M: number of multinomial draws of this event
  double resid_scale;       // residual scheme: N * 2^32 / C_N
  double thr_ratio;         // thresholds: T = min(trunc(x * thr_ratio), C_N - 1); thr_ratio = (double)C_N / (double)S_tot
  int64_t last_resample_step;  // time step whose ancestor column the last resampling event filled (0 = none yet)
  LseTriple triples[GSMC_MAX_RANKS];
  uint64_t cdf_rank_total[GSMC_MAX_RANKS];      // per-rank integer weight totals (allgathered)
  uint64_t gap_rank_total[GSMC_MAX_RANKS];      // per-rank totals of the group gaps (allgathered)
  uint64_t det_rank_total[GSMC_MAX_RANKS];      // residual scheme: per-rank totals of the deterministic copies (allgathered)
  uint64_t frac_rank_total[GSMC_MAX_RANKS];     // residual scheme: per-rank totals of the residual fractions (allgathered; they
                                                // replace cdf_rank_total once every rank's are known)
  // Sequence number of the peer exchanges, kept ON THE DEVICE (identical on every rank: all ranks run the same
  // exchanges in the same order), so that a captured / replayed launch sequence never reuses a tag. The exchange of the
  // logsumexp triples (and every other one-block exchange) advances it by 2 and uses the new value; the exchange of the
  // resampling totals of the same step, read by every block of partition_kernel, uses that value + 1 without
  // touching it. gsmc_reset leaves it (and the mailboxes) alone.
  uint32_t xseq, xseq_pad;
  // LL mailboxes for the fused peer-memory exchange: [sequence mod 4][source rank][word]; a word is
  // (payload 32 bit) | (sequence number << 32), written by the source rank with one 8-byte store.
  unsigned long long mbox[4][GSMC_MAX_RANKS][8];
};

// Peer-memory exchange of a few scalars between the ranks (one kernel per GPU, each GPU's kernel only
// waits for stores issued by the OTHER GPUs' kernels). Every rank writes its payload, 32 bits per
// 8-byte word tagged with the sequence number, straight into every peer's mailbox over NVLink
// (no fence needed: data and tag arrive in one atomic store, as in NCCL's LL protocol); a reader spins
// on the tag. Four mailboxes are used in turn (sequence number mod 4): every step has one exchange that nobody
// skips (the logsumexp triples), so a rank is never more than two exchanges ahead of a peer that has not read
// its words yet. The spin is bounded (~2 s) so a lost peer cannot hang the GPU.
struct PeerScalars { DevScalars* ds[GSMC_MAX_RANKS]; };
// Output column of a resampling event on every rank (peer-mapped): the sharded residual scheme stores ancestors into the
// rank that owns the output slot.
struct AncOut { uint32_t* col[GSMC_MAX_RANKS]; int64_t n_per; int nranks; };
__device__ __forceinline__ void anc_store(const AncOut& a, uint64_t slot, uint32_t word) {
  if (a.nranks == 1) { a.col[0][slot] = word; return; }
  const uint64_t r = slot / (uint64_t)a.n_per;
  a.col[r][slot - r * (uint64_t)a.n_per] = word;
}
__device__ __forceinline__ void ll_send(DevScalars* peer, int my_rank, uint32_t seq, const uint32_t* words, int n) {
  volatile unsigned long long* box = peer->mbox[seq & 3][my_rank];
  for (int w = 0; w < n; ++w) box[w] = (unsigned long long)words[w] | ((unsigned long long)seq << 32);
}
__device__ __forceinline__ bool ll_recv(DevScalars* self, int src, uint32_t seq, uint32_t* words, int n) {
  volatile unsigned long long* box = self->mbox[seq & 3][src];
  const long long t0 = clock64();
  for (int w = 0; w < n; ++w) {
    unsigned long long v;
    while ((uint32_t)((v = box[w]) >> 32) != seq) {
      if (clock64() - t0 > 4000000000LL) return false;
    }
    words[w] = (uint32_t)v;
  }
  return true;
}
// all ranks: send `n64` u64 values to everybody, receive everybody's into out[r*n64 ..]; threads 0..R-1 of one block
__device__ __forceinline__ void ll_allgather_u64(const PeerScalars& peers, DevScalars* self, int my_rank, int nranks, uint32_t seq,
                                                 const uint64_t* mine, int n64, uint64_t* out) {
  const int r = threadIdx.x;
  if (r < nranks) {
    uint32_t w[8];
    for (int j = 0; j < n64; ++j) { w[2 * j] = (uint32_t)mine[j]; w[2 * j + 1] = (uint32_t)(mine[j] >> 32); }
    ll_send(peers.ds[r], my_rank, seq, w, 2 * n64);
    uint32_t g[8];
    if (!ll_recv(self, r, seq, g, 2 * n64)) { self->error = 2; return; }
    for (int j = 0; j < n64; ++j) out[r * n64 + j] = (uint64_t)g[2 * j] | ((uint64_t)g[2 * j + 1] << 32);
  }
}
// How the ranks of a sharded filter exchange their per-step scalars.
//   XMODE_NONE   single rank, or the host runs ncclAllGather between the kernels (GSMC_NCCL_SCALARS=1)
//   XMODE_LL     fused: LL mailboxes in peer memory (the default on NVLink)
//   XMODE_LOCAL  shard emulation: all ranks live on ONE device and ONE stream (gsmc_group_*); a rank reads its peers'
//                scalars directly, the host having enqueued every rank's producer before any rank's consumer
enum { XMODE_NONE = 0, XMODE_LL = 1, XMODE_LOCAL = 2 };


template <typename Real> struct Vec2T;
template <> struct Vec2T<double> { typedef double2 type; };
template <> struct Vec2T<float> { typedef float2 type; };

// ------------------------------------------------------------------------------------------------
// block-level helpers (256 threads = 8 warps)
// ------------------------------------------------------------------------------------------------
// max(a, b) that keeps a when b is NaN: one compare + select (fmax() costs ~8 instructions in fp64: NaN quieting and
// signed-zero handling that the log-weight maxima do not need; a never is NaN here)
__device__ __forceinline__ double max_nn(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max_nn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
  return (uint64_t)__shfl_up_sync(0xffffffffu, (unsigned long long)v, d);
}
// inclusive scan of one u64 per thread over the block; returns inclusive value, *total = block sum.
// sm must hold GSMC_BLOCK/32 + 1 u64.
__device__ __forceinline__ uint64_t block_scan_u64(uint64_t v, uint64_t* sm, uint64_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t y = shfl_up_u64(x, d);
    if (lane >= d) x += y;
  }
  __syncthreads();                  // protect sm reuse across calls
  if (lane == 31) sm[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < GSMC_BLOCK / 32 ? sm[lane] : 0;
#pragma unroll
    for (int d = 1; d < GSMC_BLOCK / 32; d <<= 1) {
      const uint64_t y = shfl_up_u64(w, d);
      if (lane >= d) w += y;
    }
    if (lane < GSMC_BLOCK / 32) sm[lane] = w;    // inclusive warp totals
  }
  __syncthreads();
  const uint64_t warp_off = warp ? sm[warp - 1] : 0;
  *total = sm[GSMC_BLOCK / 32 - 1];
  return x + warp_off;
}
// Shared-memory copies of the lookup tables of gsmc_math.h (lanes index them divergently), staged with
// coalesced loads from their global-memory copies.
struct __align__(16) SmemTabs {
  double exp2[64];      // 2^(j/64)
  float sincosf[256];   // (sin, cos)(pi j/64), fp32 Box-Muller
  float logf[128];      // gm_nlog_u32f
};
__device__ __forceinline__ void load_tabs(SmemTabs& t, bool normals) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) t.exp2[i] = gm_exp2tab_g[i];
  if (normals) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) t.logf[i] = gm_logtabf_g[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) t.sincosf[i] = gm_sincostabf_g[i];
  }
}
// 32-byte vector accesses (sm_100: LDG/STG.256) of four consecutive column entries, widened to / narrowed from fp64
__device__ __forceinline__ void load4(const double* p, double* o) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void load4(const float* p, double* o) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = (double)v.x; o[1] = (double)v.y; o[2] = (double)v.z; o[3] = (double)v.w;
}
__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void store4(float* p, double a, double b, double c, double d) {
  *reinterpret_cast<float4*>(p) = make_float4((float)a, (float)b, (float)c, (float)d);
}

// ------------------------------------------------------------------------------------------------
// propagate: init / step, fused with the ancestor gather and the logsumexp/ESS block partials
// ------------------------------------------------------------------------------------------------
template <typename Real>
struct PropArgs {
  const Real* cur[GSMC_MAX_RANKS];   // previous state column block of every rank (peer-mapped), [D][stride]
  Real* nxt;                         // new state column block [D][stride]
  Real* lw;                          // log weights (in place)
  const uint32_t* anc;               // ancestor words of the pending resample
  const int* resampled_flag;         // device flag: was a resample decided for this step?
  LseTriple* partials;               // one per block
  DevScalars* ds;                    // the last block to finish leaves this rank's (max, s1, s2) in ds->triples[rank]
  int nranks;
  int n_tiles;                       // tiles of PropTile<Model>::TILE particles (the grid is persistent)
  int fuse_decide;                   // threshold known in advance (gsmc_run_steps): the last block also takes the
  double fuse_threshold, n_global;   //   maybe_resample! decision of the NEXT step (no finalize launch in between);
  int* next_flag;                    //   resampled[] slot of the next step
  PeerScalars peers;                 //   sharded filter: it first exchanges the ranks' triples over the LL mailboxes
  unsigned long long cond_handle;    //   captured run (CUDA graph): conditional handle of the next step's resampling block, 0 = none
  int64_t n;                         // local particle count
  int64_t stride;                    // column stride (padded n)
  uint64_t first_global;             // global index of local particle 0
  uint64_t seed;
  PhiloxKeys keys;                   // round keys of `seed`
  uint32_t t;                        // 1-based time index of the step being produced
  int use_anc;                       // 0: read cur directly; 1: consult *resampled_flag
  int rank;                          // this rank: cur[rank] is the local column
  const double* zrep;                // replay normals [n][nz] or NULL
  const double* urep;                // replay uniforms [n][nu] or NULL
};

// quads of neighbouring particles per thread: 2 (2048-particle tile) for 1-2 column models, 1 for wider states. One
// Philox call yields the four normals of four consecutive elements of the step's virtual normal array.
template <class Model> struct PropTile {
  static constexpr int QUADS = Model::D <= 2 ? GSMC_PROP_QUADS : 1;
  static constexpr int TILE = 4 * GSMC_BLOCK * QUADS;
  static constexpr int OCC = Model::D <= 2 ? GSMC_PROP_OCC : GSMC_PROP_OCC_WIDE;   // resident blocks per SM asked of the compiler
};

// Merge of the ranks' (max, s1, s2) triples by ONE WARP (all 32 lanes call it): lane r < R evaluates its own rescaling
// factor exp(m_r - M) against the maximum over the ranks -- one exp per lane in parallel instead of a chain of 2 (R - 1)
// on one thread, which sat on the critical path of every step of a sharded filter -- and the products are added in rank
// order on every lane. All ranks run the same code on the same eight triples: identical bits everywhere. One rank: the
// triple itself (exp(0) = 1 exactly). NaN sums propagate (0 * NaN), empty triples
// (max = -inf, sums 0) drop out.
__device__ __forceinline__ LseTriple merge_ranks_warp(const DevScalars* ds, int nranks) {
  const int lane = threadIdx.x & 31;
  LseTriple t; t.m = -gm_inf(); t.s1 = 0.0; t.s2 = 0.0;
  if (lane < nranks) t = ds->triples[lane];
  double M = t.m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_xor_sync(0xffffffffu, M, o); M = y > M ? y : M; }
  const double e = (t.m > -gm_inf()) ? gm_exp(t.m - M) : 0.0;
  const double p1 = t.s1 * e, p2 = t.s2 * (e * e);
  LseTriple r; r.m = M; r.s1 = 0.0; r.s2 = 0.0;
  for (int q = 0; q < nranks; ++q) { r.s1 += __shfl_sync(0xffffffffu, p1, q); r.s2 += __shfl_sync(0xffffffffu, p2, q); }
  return r;
}

// The maybe_resample! decision from the merged triple t (one thread).
// ess_threshold < 0: statistics only. resampled_flag_out: flag slot of the NEXT step.
__device__ __forceinline__ void combine_and_decide(DevScalars* ds, LseTriple t, double ess_threshold,
                                                   double n_global, int* resampled_flag_out, int64_t next_step) {
  const bool empty = !(t.m > -gm_inf()) && t.s1 == t.s1;
  const double log_total = empty ? -gm_inf() : t.m + gm_log(t.s1);       // inference.jl:3-6
  // particle_filter.jl:3-12 literally: lnw = lw - log_total; ess = exp(-logsumexp(2 lnw)) with
  // logsumexp(2 lnw) = 2 lnw_max + log(sum exp(2 (lw - max))). Written this way (not s1^2/s2) the
  // all-weights-equal case rounds exactly like the reference formula, where `ess < N` is a tie.
  const double ess = empty ? gm_nan() : gm_exp(-(2.0 * (t.m - log_total) + gm_log(t.s2)));
  ds->max_lw = t.m; ds->log_total = log_total; ds->ess = ess;
  if (ess_threshold < 0.0) return;
  int doit = ess < ess_threshold;                                        // particle_filter.jl:194
  if (doit && !(log_total > -gm_inf() && log_total < gm_inf())) { ds->error = 1; doit = 0; }
  ds->do_resample = doit;
  if (resampled_flag_out) *resampled_flag_out = doit;
  if (doit) {
    ds->log_ml_est += log_total - gm_log(n_global);                      // particle_filter.jl:201
    ds->rho = ds->n_resamples;
    ds->n_resamples += 1;
    ds->last_resample_step = next_step;                                  // its ancestor column is filled by this event
  }
}

// merge of two (max, s1, s2) triples; NaN sums propagate, empty triples (max = -inf) drop out
__device__ __forceinline__ LseTriple lse_merge_t(LseTriple a, LseTriple b, const double* etab) {
  if (!(b.m > -gm_inf()) && b.s1 == b.s1) return a;
  if (!(a.m > -gm_inf()) && a.s1 == a.s1) return b;
  // One of the two rescaling factors is exp(0) = 1 exactly: only the other one is evaluated (same bits as evaluating both).
  LseTriple r;
  const bool a_big = a.m >= b.m;
  r.m = a_big ? a.m : b.m;
  const double e = gm_exp_nonpos_t(a_big ? b.m - a.m : a.m - b.m, etab);
  const double ea = a_big ? 1.0 : e, eb = a_big ? e : 1.0;
  r.s1 = a.s1 * ea + b.s1 * eb;
  r.s2 = a.s2 * (ea * ea) + b.s2 * (eb * eb);
  return r;
}

// This rank's (max, s1, s2) from the block partials: the global max first, then the rescaled sums with one exp
// per partial, in a fixed summation order (deterministic). Called by every thread of ONE block (any size up to
// 1024); the result is valid in thread 0. sm: 96 doubles. The partials were written by other blocks: L2 loads.
__device__ __forceinline__ LseTriple reduce_partials(const LseTriple* partials, int nblk, double* sm, const double* etab) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
  double m = -gm_inf();
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) m = max_nn(m, __ldcg(&partials[b].m));
  m = warp_max(m);
  __syncthreads();
  if (lane == 0) sm[warp] = m;
  __syncthreads();
  double M = sm[0];
  for (int w = 1; w < nw; ++w) M = max_nn(M, sm[w]);
  double a1 = 0.0, a2 = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const double pm = __ldcg(&partials[b].m), p1 = __ldcg(&partials[b].s1), p2 = __ldcg(&partials[b].s2);
    const double e = (M > -gm_inf()) ? gm_exp_nonpos_t(pm - M, etab) : 1.0;
    a1 += p1 * e;
    a2 += p2 * (e * e);
  }
  a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) { sm[32 + warp] = a1; sm[64 + warp] = a2; }
  __syncthreads();
  LseTriple tr; tr.m = M; tr.s1 = 0.0; tr.s2 = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < nw; ++w) { tr.s1 += sm[32 + w]; tr.s2 += sm[64 + w]; }
  }
  return tr;
}

// models that draw a run-time number of uniforms per particle declare `static constexpr bool CTX_UNIFORMS = true`
template <class M, class = void> struct model_ctx_uniforms { static constexpr bool value = false; };
template <class M> struct model_ctx_uniforms<M, decltype((void)M::CTX_UNIFORMS)> { static constexpr bool value = M::CTX_UNIFORMS; };

template <class Model, typename Real, bool INIT, int PROP>
__global__ void __launch_bounds__(GSMC_BLOCK, PropTile<Model>::OCC) propagate_kernel(const PropArgs<Real> g, const ModelArgs a) {
  typedef typename Vec2T<Real>::type Real2;
  constexpr int D = Model::D;
  constexpr int QUADS = PropTile<Model>::QUADS, NP = 4 * QUADS, TILE = PropTile<Model>::TILE;
  constexpr int NZ = Model::nz(INIT, PROP), NU = Model::nu(INIT, PROP);
  constexpr int NZA = NZ > 0 ? NZ : 1, NUA = NU > 0 ? NU : 1;
  extern __shared__ double dyn_sm[];
  __shared__ double red[96];
  __shared__ SmemTabs tabs;
  __shared__ int s_last;
  __shared__ uint32_t s_seq;
  __shared__ uint64_t s_mine[3];
  load_tabs(tabs, NZ > 0);
  if (Model::SMEM_DOUBLES > 0) Model::template prologue<INIT, PROP>(a, dyn_sm);
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const bool gather = !INIT && g.use_anc && (*g.resampled_flag != 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Persistent blocks: block b handles tiles b, b + gridDim.x, ...; every thread carries a running (max, s1, s2)
  // over its own particles, so one block reduction, one partial, one fence and one atomic per BLOCK (not per tile) are left.
  LseTriple run; run.m = -gm_inf(); run.s1 = 0.0; run.s2 = 0.0;
  double tm = -gm_inf(), t1 = 0.0, t2 = 0.0;
  const bool unobserved = a.unobserved != 0;
  for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
  // particle j of this thread: quad q = j >> 2, position tile0 + q * 4 * GSMC_BLOCK + (j & 3)
  const int64_t tile0 = (int64_t)tile * TILE + 4 * threadIdx.x;
  // Only the last tile can hold pad lanes (the columns are padded to the tile): they are loaded, computed and
  // stored like real particles (harmless garbage; the pad words of the ancestor columns are zero) and only
  // masked out of the logsumexp partial. `lim` = real particles of this tile, block-uniform.
  const int64_t lim64 = g.n - (int64_t)tile * TILE;
  const int lim = lim64 < TILE ? (int)lim64 : TILE;

  // Stage A: previous state and log weights of all quads (all loads in flight together).
  double prev[NP][D], lwv[NP];
  if (!INIT) {
    if (gather) {
      uint4 aw[QUADS];
#pragma unroll
      for (int q = 0; q < QUADS; ++q) aw[q] = *reinterpret_cast<const uint4*>(g.anc + tile0 + (int64_t)q * 4 * GSMC_BLOCK);
#pragma unroll
      for (int q = 0; q < QUADS; ++q) {
        const uint32_t w[4] = {aw[q].x, aw[q].y, aw[q].z, aw[q].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real* c = (g.nranks == 1) ? g.cur[0] + w[k] : g.cur[w[k] >> GSMC_ANC_RANK_SHIFT] + (w[k] & GSMC_ANC_INDEX_MASK);
#pragma unroll
          for (int d = 0; d < D; ++d) prev[4 * q + k][d] = (double)__ldg(c + d * g.stride);
        }
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) lwv[j] = 0.0;      // log_weights[i] = 0. after a resample (:204)
    } else {
      const Real* cur = g.cur[g.rank];
#pragma unroll
      for (int q = 0; q < QUADS; ++q) {
        const int64_t i = tile0 + (int64_t)q * 4 * GSMC_BLOCK;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double x[4];
          load4(cur + d * g.stride + i, x);
#pragma unroll
          for (int k = 0; k < 4; ++k) prev[4 * q + k][d] = x[k];
        }
        load4(g.lw + i, &lwv[4 * q]);
      }
    }
  }

  // Stage B: draws. Quad q needs elements [(first_global+i)*NZ, +4NZ) of the step's virtual normal
  // array = NZ Philox calls of four normals each.
  double zz[QUADS][4 * NZA], uu[2 * QUADS][2 * NUA];
  if (NZ > 0) {
    if (g.zrep) {
#pragma unroll
      for (int q = 0; q < QUADS; ++q) {
        const int64_t i = tile0 + (int64_t)q * 4 * GSMC_BLOCK;
#pragma unroll
        for (int j = 0; j < 4 * NZ; ++j) zz[q][j] = (i * NZ + j < g.n * NZ) ? g.zrep[i * NZ + j] : 0.0;
      }
    } else {
      uint64_t calls[QUADS * NZA];
#pragma unroll
      for (int q = 0; q < QUADS; ++q) {
        const uint64_t c0 = ((g.first_global + (uint64_t)(tile0 + (int64_t)q * 4 * GSMC_BLOCK)) * (uint64_t)NZ) >> 2;
#pragma unroll
        for (int m = 0; m < NZ; ++m) calls[q * NZ + m] = c0 + m;
      }
      normal_quads_v<QUADS * NZA>(g.keys, calls, g.t, tabs.logf, tabs.sincosf, &zz[0][0]);
    }
  }
  if (NU > 0) {
#pragma unroll
    for (int u = 0; u < 2 * QUADS; ++u) {             // pair u: particles 2u, 2u+1 of this thread
      const int64_t i = tile0 + (int64_t)(u >> 1) * 4 * GSMC_BLOCK + 2 * (u & 1);
      if (g.urep) {
#pragma unroll
        for (int j = 0; j < 2 * NU; ++j) uu[u][j] = (i * NU + j < g.n * NU) ? g.urep[i * NU + j] : 0.0;
      } else {
        const uint64_t c0 = ((g.first_global + (uint64_t)i) * (uint64_t)NU) >> 1;
#pragma unroll
        for (int m = 0; m < NU; ++m) uniform_pair(g.seed, c0 + m, g.t, GSMC_STREAM_UNIFORM, &uu[u][2 * m], &uu[u][2 * m + 1]);
      }
    }
  }

  // Stage C: the model, per particle. log_weights[i] = weight (init) / += increment (step).
  // Stage D: vector stores of the new state and log weights.
#pragma unroll
  for (int q = 0; q < QUADS; ++q) {
    const int64_t i = tile0 + (int64_t)q * 4 * GSMC_BLOCK;
    double out[4][D], w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = 4 * q + k;
      if constexpr (model_ctx_uniforms<Model>::value) {
        // run-time number of uniforms per particle: the model draws them itself from the same virtual array
        const int nu_rt = (int)a.p[0];
        DrawCtx c;
        c.seed = g.seed; c.t = g.t;
        c.global_index = g.first_global + (uint64_t)(i + k);
        // pad lanes (i >= n) have no replayed values: they draw from Philox like everybody else (results are discarded)
        c.urep = (g.urep && i + k < g.n) ? g.urep + (i + k) * nu_rt : nullptr;
        w[k] = Model::template particle_ctx<INIT, PROP>(a, dyn_sm, prev[j], &zz[q][k * NZ], c, out[k]);
      } else {
        w[k] = Model::template particle<INIT, PROP>(a, dyn_sm, prev[j], &zz[q][k * NZ], &uu[j >> 1][(k & 1) * NU], out[k]);
      }
      if (unobserved) w[k] = 0.0;                     // no constrained choice at this step (static_ir/generate.jl:36-42)
    }
    Real r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = (Real)(INIT ? w[k] : lwv[4 * q + k] + w[k]);
#pragma unroll
    for (int d = 0; d < D; ++d) store4(g.nxt + d * g.stride + i, out[0][d], out[1][d], out[2][d], out[3][d]);
    store4(g.lw + i, (double)r[0], (double)r[1], (double)r[2], (double)r[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) lwv[4 * q + k] = (double)r[k];
  }
  if (lim < TILE) {                                            // last tile only: pad lanes drop out of the reduction
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int o = 4 * (int)threadIdx.x + (j >> 2) * 4 * GSMC_BLOCK + (j & 3);
      if (o >= lim) lwv[j] = -gm_inf();
    }
  }

  // Stage E: this THREAD's running (max, sum exp(lw-max), sum exp(2(lw-max))) over all its particles of all tiles: no
  // shuffle, barrier or shared-memory traffic per tile; the block combines its threads' triples once, after the tile
  // loop. NaN log weights poison s1/s2 on purpose: max_nn drops NaNs (the accumulator never is NaN), so the max needs
  // no NaN test, and a NaN log weight turns into a NaN exp below. Two chains of compares rather than one.
  double ma = max_nn(-gm_inf(), lwv[0]), mb = max_nn(-gm_inf(), lwv[NP / 2]);
#pragma unroll
  for (int j = 1; j < NP / 2; ++j) { ma = max_nn(ma, lwv[j]); mb = max_nn(mb, lwv[NP / 2 + j]); }
  const double mn = max_nn(max_nn(tm, ma), mb);
  if (mn > -gm_inf()) {
    double x[NP + 1], e[NP + 1];
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = lwv[j] - mn;             // -inf for pad lanes -> exp = 0
    x[NP] = tm - mn;                                             // rescales the sums so far (-inf the first time: they are 0)
    gm_exp_nonpos_v<NP + 1>(x, e, tabs.exp2);
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int j = 0; j < NP; ++j) { a1 += e[j]; a2 += e[j] * e[j]; }
    t1 = t1 * e[NP] + a1;
    t2 = t2 * (e[NP] * e[NP]) + a2;
    tm = mn;
  } else {
    // -inf / NaN log weights only so far: no exp is evaluated, so look for the NaNs explicitly
    bool any_nan = false;
#pragma unroll
    for (int j = 0; j < NP; ++j) any_nan = any_nan || (lwv[j] != lwv[j]);
    if (any_nan) { t1 = gm_nan(); t2 = gm_nan(); }
  }
  }  // tiles
  // The block's triple from its threads' triples: block max, one exp per thread, two block sums.
  {
    const double m = warp_max(tm);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    // every group of 8 lanes loads the 8 warp maxima and folds them with a 3-step butterfly
    double bm = red[lane & (GSMC_BLOCK / 32 - 1)];
#pragma unroll
    for (int o = GSMC_BLOCK / 64; o > 0; o >>= 1) bm = max_nn(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    double s1 = t1, s2 = t2;                                     // (all -inf: 0, or NaN when a NaN log weight was seen)
    if (bm > -gm_inf()) {
      const double e = gm_exp_nonpos_t(tm - bm, tabs.exp2);
      s1 = t1 * e; s2 = t2 * (e * e);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { red[32 + warp] = s1; red[64 + warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double b1 = 0.0, b2 = 0.0;
#pragma unroll
      for (int w = 0; w < GSMC_BLOCK / 32; ++w) { b1 += red[32 + w]; b2 += red[64 + w]; }
      run.m = bm; run.s1 = b1; run.s2 = b2;
    }
  }
  if (threadIdx.x == 0) {
    g.partials[blockIdx.x] = run;
    // The last block to publish its partial reduces all of them (replaces a separate one-block launch).
    __threadfence();
    s_last = (atomicAdd(&g.ds->blocks_done, 1u) + 1u == gridDim.x) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const LseTriple tr = reduce_partials(g.partials, (int)gridDim.x, red, tabs.exp2);
    if (threadIdx.x == 0) {
      g.ds->triples[g.rank] = tr; g.ds->blocks_done = 0;
      if (g.fuse_decide && g.nranks == 1) {
        combine_and_decide(g.ds, tr, g.fuse_threshold, g.n_global, g.next_flag, (int64_t)g.t + 1);
        if (g.cond_handle) cudaGraphSetConditional(g.cond_handle, g.ds->do_resample ? 1u : 0u);
      }
    }
    if (g.fuse_decide && g.nranks > 1) {
      // Sharded filter: this block sends the rank's triple to every peer as soon as it is known and merges all of them
      // in rank order (the logsumexp "allreduce"), so every rank takes the same decision without a finalize launch.
      if (threadIdx.x == 0) {
        s_mine[0] = gm_to_bits(tr.m); s_mine[1] = gm_to_bits(tr.s1); s_mine[2] = gm_to_bits(tr.s2);
        s_seq = (g.ds->xseq += 2);
      }
      __syncthreads();
      ll_allgather_u64(g.peers, g.ds, g.rank, g.nranks, s_seq, s_mine, 3, reinterpret_cast<uint64_t*>(g.ds->triples));
      __syncthreads();
      if (threadIdx.x < 32) {
        const LseTriple all = merge_ranks_warp(g.ds, g.nranks);
        if (threadIdx.x == 0) {
          combine_and_decide(g.ds, all, g.fuse_threshold, g.n_global, g.next_flag, (int64_t)g.t + 1);
          if (g.cond_handle) cudaGraphSetConditional(g.cond_handle, g.ds->do_resample ? 1u : 0u);
        }
      }
    }
  }
}

// Unobserved step: the observation choice is sampled from the model given the particle's NEW latent
// (static_ir/generate.jl:36-42: an unconstrained choice is random(dist, args...), weight unchanged). One draw per
// particle from Philox stream GSMC_STREAM_OBS (element = global particle index). A rare path: one thread per particle.
template <class Model, typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK) sample_obs_kernel(const ModelArgs a, const Real* state, Real* obs_col, int64_t n, int64_t stride,
                                                                uint64_t first_global, uint64_t seed, uint32_t t) {
  const int64_t i = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  if (i >= n) return;
  double lat[Model::D];
#pragma unroll
  for (int d = 0; d < Model::D; ++d) lat[d] = (double)state[d * stride + i];
  const uint64_t e = first_global + (uint64_t)i;
  double draw;
  if (Model::OBS_DRAW_UNIFORM) {
    double u0, u1;
    uniform_pair(seed, e >> 1, t, GSMC_STREAM_OBS, &u0, &u1);
    draw = (e & 1) ? u1 : u0;
  } else {
    double z0, z1;
    normal_pair(seed, e >> 1, t, gm_logtab64_g, &z0, &z1, GSMC_STREAM_OBS);
    draw = (e & 1) ? z1 : z0;
  }
  obs_col[i] = (Real)Model::sample_obs(a, lat, draw);
}

// ------------------------------------------------------------------------------------------------
// finalize: merge the ranks' triples and decide.
// ------------------------------------------------------------------------------------------------
// maybe_resample! returns a Bool to the host: instead of a D2H copy + event, the deciding thread stores the few scalars
// the host reads straight into the pinned (device-mapped) host mirror and then, after a system fence, the call's
// token; the host spins on the token (a couple of microseconds instead of a DMA + event round trip).
__device__ __forceinline__ void publish_decision(const DevScalars* ds, DevScalars* host, uint32_t token) {
  if (!host) return;
  host->max_lw = ds->max_lw; host->log_total = ds->log_total; host->ess = ds->ess; host->log_ml_est = ds->log_ml_est;
  host->do_resample = ds->do_resample; host->error = ds->error; host->n_resamples = ds->n_resamples; host->rho = ds->rho;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned int*>(&host->host_token) = token;
}

// The logsumexp/ESS statistics and the maybe_resample! decision from this rank's triple (left in ds->triples[rank]
// by the last block of the propagate kernel). Multi-rank: every rank first gathers all triples over NVLink peer
// stores (the logsumexp "allreduce") and merges them in rank order, so all ranks take the same decision
// without a separate collective.
__global__ void __launch_bounds__(32) finalize_kernel(DevScalars* ds, int rank, int nranks, double ess_threshold,
                                                      double n_global, int* resampled_flag_out, int64_t next_step,
                                                      PeerScalars peers, int xmode, DevScalars* host, uint32_t token,
                                                      unsigned long long cond_handle) {
  __shared__ uint64_t mine[3];
  pdl_wait();
  pdl_trigger();
  if (nranks > 1 && xmode == XMODE_LL) {
    uint32_t seq = 0;
    if (threadIdx.x == 0) {
      const LseTriple tr = ds->triples[rank];
      mine[0] = gm_to_bits(tr.m); mine[1] = gm_to_bits(tr.s1); mine[2] = gm_to_bits(tr.s2);
      seq = (ds->xseq += 2);
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    __syncwarp();
    ll_allgather_u64(peers, ds, rank, nranks, seq, mine, 3, reinterpret_cast<uint64_t*>(ds->triples));
    __syncwarp();
    const LseTriple all = merge_ranks_warp(ds, nranks);
    if (threadIdx.x == 0) { combine_and_decide(ds, all, ess_threshold, n_global, resampled_flag_out, next_step); publish_decision(ds, host, token); if (cond_handle) cudaGraphSetConditional(cond_handle, ds->do_resample ? 1u : 0u); }
  } else if (nranks > 1 && xmode == XMODE_LOCAL) {
    if ((int)threadIdx.x < nranks && (int)threadIdx.x != rank) ds->triples[threadIdx.x] = peers.ds[threadIdx.x]->triples[threadIdx.x];
    __syncwarp();
    const LseTriple all = merge_ranks_warp(ds, nranks);
    if (threadIdx.x == 0) { combine_and_decide(ds, all, ess_threshold, n_global, resampled_flag_out, next_step); publish_decision(ds, host, token); if (cond_handle) cudaGraphSetConditional(cond_handle, ds->do_resample ? 1u : 0u); }
  } else if (nranks == 1) {
    if (threadIdx.x == 0) { combine_and_decide(ds, ds->triples[0], ess_threshold, n_global, resampled_flag_out, next_step); publish_decision(ds, host, token); if (cond_handle) cudaGraphSetConditional(cond_handle, ds->do_resample ? 1u : 0u); }
  }
}
// cross-GPU barrier (peer-memory exchange of one word): nobody passes until every rank has arrived
__global__ void peer_barrier_kernel(PeerScalars peers, DevScalars* ds, int rank, int nranks) {
  __shared__ uint64_t mine[1];
  __shared__ uint64_t got[GSMC_MAX_RANKS];
  __shared__ uint32_t s_seq;
  if (threadIdx.x == 0) { s_seq = (ds->xseq += 2); mine[0] = s_seq; }
  __syncthreads();
  ll_allgather_u64(peers, ds, rank, nranks, s_seq, mine, 1, got);
}
// the same inside a stream of conditional kernels: the stores every rank issued before it (system-wide fence) are visible
// to every rank after it; skipped by everybody when no resample was decided (the sequence number then does not advance)
__global__ void peer_fence_kernel(PeerScalars peers, DevScalars* ds, int rank, int nranks, int conditional) {
  __shared__ uint64_t mine[1];
  __shared__ uint64_t got[GSMC_MAX_RANKS];
  __shared__ uint32_t s_seq;
  pdl_wait();
  if (conditional && !ds->do_resample) return;
  if (threadIdx.x == 0) { __threadfence_system(); s_seq = (ds->xseq += 2); mine[0] = s_seq; }
  __syncthreads();
  ll_allgather_u64(peers, ds, rank, nranks, s_seq, mine, 1, got);
  __threadfence_system();
}
// shard emulation (XMODE_LOCAL): copy the peers' own entries of the per-rank scalars into this rank's arrays
enum { PEER_COPY_TRIPLES = 1, PEER_COPY_CDF = 2, PEER_COPY_GAP = 4, PEER_COPY_DET = 8 };
__global__ void peer_copy_kernel(PeerScalars peers, DevScalars* ds, int rank, int nranks, int what, int conditional) {
  if (conditional && !ds->do_resample) return;
  const int r = threadIdx.x;
  if (r >= nranks || r == rank) return;
  if (what & PEER_COPY_TRIPLES) ds->triples[r] = peers.ds[r]->triples[r];
  if (what & PEER_COPY_CDF) ds->cdf_rank_total[r] = peers.ds[r]->cdf_rank_total[r];
  if (what & PEER_COPY_GAP) ds->gap_rank_total[r] = peers.ds[r]->gap_rank_total[r];
  if (what & PEER_COPY_DET) { ds->det_rank_total[r] = peers.ds[r]->det_rank_total[r]; ds->frac_rank_total[r] = peers.ds[r]->frac_rank_total[r]; }
}

// multi-rank: runs after the allgather of ds->triples
__global__ void decide_kernel(DevScalars* ds, int nranks, double ess_threshold, double n_global, int* resampled_flag_out,
                              int64_t next_step, DevScalars* host, uint32_t token) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const LseTriple all = merge_ranks_warp(ds, nranks);
  if (threadIdx.x == 0) { combine_and_decide(ds, all, ess_threshold, n_global, resampled_flag_out, next_step); publish_decision(ds, host, token); }
}

// ------------------------------------------------------------------------------------------------
// integer weights q_i = floor(exp(lw_i - max) * 2^k) and their CDF, stored in TWO LEVELS:
//   cl[i]  = sum of q over the particles of i's SEGMENT up to and including i   (segment-local inclusive CDF)
//   sp[s]  = sum of q over the segments before s (exclusive segment prefix), sp[n_segs] = this rank's total
// so that C_i = (rank offset) + sp[segment(i)] + cl[i]. A segment is a run of seg_tiles consecutive
// 2048-particle tiles owned by ONE block of the streaming pass, which carries the running sum in a
// register: no inter-block dependency, no look-back, and only n_segs (a few hundred) totals are left to
// scan. The global CDF is never materialised (lw is read once, exp evaluated once). The Gamma gaps of the
// groups of sorted draws (one per 256 output slots) and their segment-local tile prefixes are generated by
// the same pass.
// ------------------------------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ void q_from_lw(typename Vec2T<Real>::type a, typename Vec2T<Real>::type b, int64_t i, int64_t n,
                                          double mx, double scale, const double* etab, uint64_t q[4]) {
  const double x[4] = {(double)a.x - mx, (double)a.y - mx, (double)b.x - mx, (double)b.y - mx};
  double e[4];
  gm_exp_nonpos_v<4>(x, e, etab);
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = (i + j < n) ? (uint64_t)floor(e[j] * scale) : 0;
}
template <typename Real>
__device__ __forceinline__ void load_q4(const Real* lw, int64_t i, int64_t n, double mx, double scale, const double* etab, uint64_t q[4]) {
  // i is a multiple of 4 and the columns are padded to a tile, so the vector loads stay inside the allocation
  typedef typename Vec2T<Real>::type Real2;
  const Real2 a = *reinterpret_cast<const Real2*>(lw + i);
  const Real2 b = *reinterpret_cast<const Real2*>(lw + i + 2);
  q_from_lw<Real>(a, b, i, n, mx, scale, etab, q);
}

// Block-wide inclusive scan of v, one barrier per call.
// sm: [2 buffers][GSMC_BLOCK/32] u64, `buf` alternates between consecutive calls.
__device__ __forceinline__ uint64_t block_scan_incl(uint64_t v, uint64_t* sm, int buf, uint64_t* v_total) {
  constexpr int NW = GSMC_BLOCK / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint64_t y = shfl_up_u64(x, d); if (lane >= d) x += y; }
  uint64_t* sv = sm + buf * NW;
  if (lane == 31) sv[warp] = x;
  __syncthreads();
  // Every warp combines the NW warp totals with shuffles: lane k < NW holds total k, a 3-step inclusive scan gives the
  // prefixes (instead of every thread reading and adding all NW totals itself).
  uint64_t p = lane < NW ? sv[lane] : 0;
#pragma unroll
  for (int d = 1; d < NW; d <<= 1) { const uint64_t y = shfl_up_u64(p, d); if (lane >= d) p += y; }
  *v_total = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)p, NW - 1);
  const uint64_t off = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)p, warp ? warp - 1 : 0);
  return x + (warp ? off : 0);
}

// The streaming pass of a resampling event; block s owns segment s = tiles [s*seg_tiles, (s+1)*seg_tiles).
// WEIGHTS: lw -> q -> segment-local inclusive CDF cl + segment total seg_q[s].
// GAPS: the Gamma gaps of the groups (GSMC_GROUP output slots each) of this rank's draws [k_first, k_first + nt*TILE),
// one thread per group up front (a few hundred instructions each, 8 groups per tile), the segment-local exclusive
// prefix tile_e[tile] of every tile and the segment total seg_e[s]; block 0 also draws the event's head gap.
// m_draws_arg: number of draws M when the host knows it (multinomial: N), 0 = read ds->n_draws.
#ifndef GSMC_WK_OCC
#define GSMC_WK_OCC 4
#endif
#define GSMC_WPT (GSMC_TILE / GSMC_BLOCK)     // elements per thread and tile of the streaming pass: 8
#define GSMC_GPT (GSMC_TILE / GSMC_GROUP)     // groups per tile: 8
template <typename Real, bool WEIGHTS, bool GAPS>
__global__ void __launch_bounds__(GSMC_BLOCK, GSMC_WK_OCC) weights_kernel(const Real* lw, int64_t n, double scale, DevScalars* ds,
                                                             uint64_t* cl, uint64_t* seg_q, uint64_t seed, uint64_t k_first,
                                                             uint64_t m_draws_arg, uint64_t* gap, uint64_t* tile_e, uint64_t* seg_e,
                                                             int nt, int seg_tiles, int conditional) {
  typedef typename Vec2T<Real>::type Real2;
  constexpr int W = GSMC_WPT;
  __shared__ uint64_t sm[2 * (GSMC_BLOCK / 32)];
  __shared__ double etab[64];
  if (WEIGHTS && threadIdx.x < 64) etab[threadIdx.x] = gm_exp2tab_g[threadIdx.x];
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (conditional && !ds->do_resample) return;
  const double mx = ds->max_lw;
  const int t0 = blockIdx.x * seg_tiles, t1 = min(t0 + seg_tiles, nt);
  Real2 lv[W / 2];
  if (WEIGHTS) {                                          // first tile's log weights are in flight during the gap draws
    const int64_t i0 = (int64_t)t0 * GSMC_TILE + W * threadIdx.x;
#pragma unroll
    for (int j = 0; j < W / 2; ++j) lv[j] = *reinterpret_cast<const Real2*>(lw + i0 + 2 * j);
  }
  if (GAPS) {
    const uint32_t rho = ds->rho;
    const uint64_t m_draws = m_draws_arg ? m_draws_arg : ds->n_draws;
    for (int gl = t0 * GSMC_GPT + (int)threadIdx.x; gl < t1 * GSMC_GPT; gl += GSMC_BLOCK)
      gap[gl] = gap_of_group(seed, k_first + (uint64_t)gl * GSMC_GROUP, m_draws, rho);
    if (blockIdx.x == 0 && threadIdx.x == GSMC_BLOCK - 1) ds->gap_head = gap_head(seed, m_draws, rho);
  }
  uint64_t run_q = 0;                                     // sum over the tiles of this segment done so far
  int buf = 0;
  if (WEIGHTS) {
    for (int tile = t0; tile < t1; ++tile, buf ^= 1) {
      const int64_t i = (int64_t)tile * GSMC_TILE + W * threadIdx.x;
      uint64_t q[W];
      double x[W], ex[W];
#pragma unroll
      for (int j = 0; j < W / 2; ++j) { x[2 * j] = (double)lv[j].x - mx; x[2 * j + 1] = (double)lv[j].y - mx; }
      if (tile + 1 < t1) {                                // next tile's log weights are in flight during the arithmetic
#pragma unroll
        for (int j = 0; j < W / 2; ++j) lv[j] = *reinterpret_cast<const Real2*>(lw + i + GSMC_TILE + 2 * j);
      }
      gm_exp_nonpos_v<W>(x, ex, etab);
#pragma unroll
      for (int j = 0; j < W; ++j) q[j] = (uint64_t)(ex[j] * scale);          // = floor: the product is >= 0
      if ((int64_t)(tile + 1) * GSMC_TILE > n) {                             // the tile that holds the pad lanes (block-uniform)
#pragma unroll
        for (int j = 0; j < W; ++j) if (i + j >= n) q[j] = 0;
      }
      uint64_t qs = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) qs += q[j];
      uint64_t qt;
      const uint64_t incl = block_scan_incl(qs, sm, buf, &qt);
      uint64_t c = run_q + incl - qs;
#pragma unroll
      for (int j = 0; j < W; j += 2) {
        ulonglong2 o;
        c += q[j]; o.x = c; c += q[j + 1]; o.y = c;
        *reinterpret_cast<ulonglong2*>(cl + i + j) = o;
      }
      run_q += qt;
    }
    if (threadIdx.x == 0) seg_q[blockIdx.x] = run_q;
  }
  if (GAPS) {
    // tile prefixes of the gaps (written above by other threads of this block): warp 0, 32 tiles per round
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      uint64_t carry = 0;
      for (int base = t0; base < t1; base += 32) {
        const int tile = base + lane;
        uint64_t sum = 0;
        if (tile < t1) {
          const ulonglong2* gp = reinterpret_cast<const ulonglong2*>(gap + (int64_t)tile * GSMC_GPT);
#pragma unroll
          for (int j = 0; j < GSMC_GPT / 2; ++j) { const ulonglong2 g2 = gp[j]; sum += g2.x + g2.y; }
        }
        uint64_t xs = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t y = shfl_up_u64(xs, d); if (lane >= d) xs += y; }
        if (tile < t1) tile_e[tile] = carry + xs - sum;
        carry += (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)xs, 31);
      }
      if (lane == 0) seg_e[blockIdx.x] = carry;
    }
  }
}

// What a scan_segments launch completes after the scans (bit set).
enum { SCAN_Q = 1, SCAN_SET_DRAWS = 2, SCAN_E = 4, SCAN_RESID = 8 };

// Totals of a resampling event once every rank's totals are known (thread 0 of one block):
//   SCAN_Q      cdf_total = sum of the ranks' integer weight totals  [SCAN_SET_DRAWS: M = N draws, no copies]
//   SCAN_E      S_tot = head gap + all ranks' gap totals, and the threshold ratio
//   SCAN_RESID  n_det = sum c over all ranks, M = N - n_det, cdf_total = sum of the residual fractions
__device__ __forceinline__ void finish_totals(DevScalars* ds, int nranks, uint64_t seed, uint64_t n_global, int what,
                                              uint64_t total0, uint64_t total1) {
  if (what & SCAN_RESID) {
    if (nranks == 1) { ds->det_rank_total[0] = total0; ds->frac_rank_total[0] = total1; }
    uint64_t d = 0, c = 0;
    for (int r = 0; r < nranks; ++r) { d += ds->det_rank_total[r]; c += ds->frac_rank_total[r]; ds->cdf_rank_total[r] = ds->frac_rank_total[r]; }
    ds->n_det = d;
    ds->n_draws = n_global - d;
    ds->cdf_total = c;
  }
  if (what & SCAN_Q) {
    uint64_t s = 0;
    for (int r = 0; r < nranks; ++r) s += ds->cdf_rank_total[r];
    ds->cdf_total = s;
    if (what & SCAN_SET_DRAWS) { ds->n_draws = n_global; ds->n_det = 0; }
  }
  if (what & SCAN_E) {
    uint64_t s = 0;
    for (int r = 0; r < nranks; ++r) s += ds->gap_rank_total[r];
    const uint64_t stot = ds->gap_head + s;               // the head gap was drawn by the gap pass
    ds->gap_total = stot;
    ds->thr_ratio = stot ? (double)ds->cdf_total / (double)stot : 0.0;
  }
}

// The segment prefixes stay RANK-LOCAL (they are final before a rank announces its totals, so a peer that has
// received the totals can read them without a further handshake); consumers add the integer weight of the lower
// ranks, off_r = sum of cdf_rank_total[0..r), themselves: C_i = off_r + sp[segment(i)] + cl[i].
// Block-wide exclusive scans of up to two arrays of n_segs <= 1024 segment totals (one element per thread
// of a 1024-thread block): out0/out1[i] = exclusive prefix, out[n_segs] = total. out may be shared or global.
__device__ __forceinline__ void scan_segments_block(const uint64_t* in0, const uint64_t* in1, int n_segs, uint64_t* out0, uint64_t* out1,
                                                    uint64_t (*sm)[33], uint64_t totals[2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = threadIdx.x;
  uint64_t v[2], x[2];
#pragma unroll
  for (int arr = 0; arr < 2; ++arr) {
    const uint64_t* a = arr ? in1 : in0;
    v[arr] = (a && i < n_segs) ? a[i] : 0;
    x[arr] = v[arr];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t y = shfl_up_u64(x[arr], d); if (lane >= d) x[arr] += y; }
    if (lane == 31) sm[arr][warp] = x[arr];
  }
  __syncthreads();
  if (warp < 2) {
    uint64_t w = sm[warp][lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t y = shfl_up_u64(w, d); if (lane >= d) w += y; }
    sm[warp][lane] = w;
  }
  __syncthreads();
#pragma unroll
  for (int arr = 0; arr < 2; ++arr) {
    uint64_t* o = arr ? out1 : out0;
    totals[arr] = sm[arr][31];
    if (!(arr ? in1 : in0) || !o) continue;
    const uint64_t incl = x[arr] + (warp ? sm[arr][warp - 1] : 0);
    if (i < n_segs) o[i] = incl - v[arr];
    if (i == 0) o[n_segs] = totals[arr];
  }
}

// One block: scans the raw segment totals in0/in1 into the prefix arrays out0/out1; then this rank's totals
// go to ds, are exchanged with the peers (fused LL exchange over NVLink), finish_totals runs and the
// prefixes are made global. With exchange == 0 on a multi-rank run the host performs the allgathers and
// launches totals_kernel.
__global__ void __launch_bounds__(1024) scan_segments_kernel(const uint64_t* in0, const uint64_t* in1, uint64_t* out0, uint64_t* out1,
                                                             int n_segs, DevScalars* ds, int what,
                                                             uint64_t seed, uint64_t n_global, int conditional,
                                                             PeerScalars peers, int rank, int nranks, int exchange) {
  __shared__ uint64_t sm[2][33];
  __shared__ uint64_t mine[2];
  __shared__ uint64_t got[2 * GSMC_MAX_RANKS];
  __shared__ uint32_t s_seq;
  pdl_wait();
  pdl_trigger();
  const bool skip = conditional && !ds->do_resample;
  const int n64 = (what & SCAN_RESID) ? 2 : ((what & SCAN_Q) ? 1 : 0) + ((what & SCAN_E) ? 1 : 0);
  uint64_t totals[2] = {0, 0};
  if (!skip) scan_segments_block(in0, in1, n_segs, out0, out1, sm, totals);
  __threadfence_system();                            // the (rank-local) prefixes are visible to the peers before the totals are sent
  // this rank's totals: the weights' first when both are present
  if (threadIdx.x == 0) {
    int k = 0;
    if (what & SCAN_Q) { ds->cdf_rank_total[rank] = totals[0]; mine[k++] = totals[0]; }
    if (what & SCAN_E) { const uint64_t t = (what & SCAN_Q) ? totals[1] : totals[0]; ds->gap_rank_total[rank] = t; mine[k++] = t; }
    if (what & SCAN_RESID) { ds->det_rank_total[rank] = totals[0]; ds->frac_rank_total[rank] = totals[1]; mine[0] = totals[0]; mine[1] = totals[1]; }
    if (nranks > 1 && exchange && n64) s_seq = (ds->xseq += 2);
  }
  __syncthreads();
  if (nranks > 1) {
    if (!exchange || n64 == 0) return;
    // the exchange runs on every rank even when no resample was decided (all ranks advance the sequence number alike)
    ll_allgather_u64(peers, ds, rank, nranks, s_seq, mine, n64, got);
    __syncthreads();
    if (threadIdx.x == 0 && !skip) {
      for (int r = 0; r < nranks; ++r) {
        int k = 0;
        if (what & SCAN_Q) ds->cdf_rank_total[r] = got[r * n64 + k++];
        if (what & SCAN_E) ds->gap_rank_total[r] = got[r * n64 + k++];
        if (what & SCAN_RESID) { ds->det_rank_total[r] = got[r * n64]; ds->frac_rank_total[r] = got[r * n64 + 1]; }
      }
    }
  }
  if (threadIdx.x == 0 && !skip) finish_totals(ds, nranks, seed, n_global, what, totals[0], totals[1]);
}
// multi-rank runs that exchange the totals with ncclAllGather (GSMC_NCCL_SCALARS=1) finish here
__global__ void __launch_bounds__(1024) totals_kernel(uint64_t* a0, uint64_t* a1, int n_segs, DevScalars* ds, int nranks, int rank,
                                                      uint64_t seed, uint64_t n_global, int what, int conditional) {
  if (conditional && !ds->do_resample) return;
  if (threadIdx.x == 0) finish_totals(ds, nranks, seed, n_global, what, 0, 0);
}

// residual scheme: e_i = floor(q_i * resid_scale); c_i = e_i >> 32 copies; r_i = e_i & (2^32-1)
__device__ __forceinline__ void resid_split(uint64_t q, double scale, uint64_t* c, uint64_t* r) {
  const uint64_t e = (uint64_t)floor((double)q * scale);
  *c = e >> 32; *r = e & 0xffffffffULL;
}
__global__ void resid_scale_kernel(DevScalars* ds, double n_global) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0 && blockIdx.x == 0 && ds->do_resample)
    ds->resid_scale = (n_global * 4294967296.0) / (double)ds->cdf_total;
}
// segment-local inclusive CDFs of the copy counts (cc) and of the residual fractions (cl) + the segment totals;
// same streaming structure as weights_kernel (8 elements per thread and tile, one barrier per block scan)
template <typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK, GSMC_WK_OCC) resid_cdf_kernel(const Real* lw, int64_t n, double scale, const DevScalars* ds,
                                                                            uint64_t* cc, uint64_t* seg_c, uint64_t* cl, uint64_t* seg_r,
                                                                            int nt, int seg_tiles, int conditional) {
  typedef typename Vec2T<Real>::type Real2;
  constexpr int W = GSMC_WPT;
  __shared__ uint64_t sm_c[2 * (GSMC_BLOCK / 32)];
  __shared__ uint64_t sm_r[2 * (GSMC_BLOCK / 32)];
  __shared__ double etab[64];
  if (threadIdx.x < 64) etab[threadIdx.x] = gm_exp2tab_g[threadIdx.x];
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (conditional && !ds->do_resample) return;
  const double mx = ds->max_lw, rscale = ds->resid_scale;
  const int t0 = blockIdx.x * seg_tiles, t1 = min(t0 + seg_tiles, nt);
  uint64_t run_c = 0, run_r = 0;
  int buf = 0;
  for (int tile = t0; tile < t1; ++tile, buf ^= 1) {
    const int64_t i = (int64_t)tile * GSMC_TILE + W * threadIdx.x;
    double x[W], ex[W];
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
      const Real2 l = *reinterpret_cast<const Real2*>(lw + i + 2 * j);
      x[2 * j] = (double)l.x - mx; x[2 * j + 1] = (double)l.y - mx;
    }
    gm_exp_nonpos_v<W>(x, ex, etab);
    uint64_t c[W], r[W], cs = 0, rs = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint64_t q = (uint64_t)(ex[j] * scale);                      // = floor: the product is >= 0 (as in weights_kernel)
      resid_split(q, rscale, &c[j], &r[j]);
      if (i + j >= n) { c[j] = 0; r[j] = 0; }
      cs += c[j]; rs += r[j];
    }
    uint64_t ctot, rtot;
    uint64_t ic = run_c + block_scan_incl(cs, sm_c, buf, &ctot) - cs;
    uint64_t ir = run_r + block_scan_incl(rs, sm_r, buf, &rtot) - rs;
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      ulonglong2 oc, orr;
      ic += c[j]; oc.x = ic; ic += c[j + 1]; oc.y = ic;
      ir += r[j]; orr.x = ir; ir += r[j + 1]; orr.y = ir;
      *reinterpret_cast<ulonglong2*>(cc + i + j) = oc;
      *reinterpret_cast<ulonglong2*>(cl + i + j) = orr;
    }
    run_c += ctot; run_r += rtot;
  }
  if (threadIdx.x == 0) { seg_c[blockIdx.x] = run_c; seg_r[blockIdx.x] = run_r; }
}

// ------------------------------------------------------------------------------------------------
// search: ancestor of every output slot
// ------------------------------------------------------------------------------------------------
struct CdfView {
  const uint64_t* seg[GSMC_MAX_RANKS];   // segment-local inclusive CDF of every rank (peer-mapped)
  const uint64_t* sp[GSMC_MAX_RANKS];    // exclusive segment prefixes of every rank, n_segs+1 entries (peer-mapped)
  int64_t n_per;                         // particles per rank
  int n_pad;                             // padded particles per rank (the last segment ends here)
  int seg_len;                           // particles per segment (a multiple of GSMC_TILE)
  int n_segs;                            // segments per rank
  int nranks;
};
// The threshold predicate "C > T" over integer CDF values C. iid / replay mode: T_j = floor(floor(u_j 2^53) C_N / 2^53);
// grouped order statistics: T_k = min(trunc(x_k ratio), C_N - 1) (threshold_u64 in gsmc_rng.cuh).
struct GtU64 { uint64_t T; __device__ __forceinline__ bool operator()(uint64_t c) const { return c > T; } };

// smallest j in [0, len) with gt(add + arr[j]); len if none
template <class P>
__device__ __forceinline__ int upper_pred(const uint64_t* arr, int len, uint64_t add, const P gt) {
  int l = 0, h = len;
  while (l < h) {
    const int mid = (l + h) >> 1;
    if (gt(add + __ldg(arr + mid))) h = mid; else l = mid + 1;
  }
  return l;
}
// the same by a full warp: 32 probes per round trip (32-ary search); arr may live in shared or global memory
template <class P>
__device__ __forceinline__ int upper_pred_warp(const uint64_t* arr, int len, uint64_t add, const P gt) {
  const int lane = threadIdx.x & 31;
  int lo = 0, hi = len;                           // answer in [lo, hi]; hi = len means "none"
  while (hi > lo) {
    const int step = (hi - lo + 31) >> 5;
    const int p = lo + lane * step;
    const bool pred = (p >= hi) || gt(add + arr[p]);
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    const int f = mask ? __ffs((int)mask) - 1 : 32;
    const int new_hi = (f == 32) ? hi : lo + f * step;
    const int new_lo = (f == 0) ? lo : lo + (f - 1) * step + 1;
    hi = new_hi < hi ? new_hi : hi;
    lo = new_lo;
    if (f == 0) hi = lo;
  }
  return lo;
}
// integer weight of the ranks before r
__device__ __forceinline__ uint64_t rank_offset(const DevScalars* ds, int r) {
  uint64_t off = 0;
  for (int q = 0; q < r; ++q) off += ds->cdf_rank_total[q];
  return off;
}
// Ancestor word of a threshold against the global CDF: min{i : gt(C_i)}, clamped to the last particle.
// Three levels: owner rank (its inclusive end passes the predicate), segment, position in the segment.
template <bool WARP, class P>
__device__ __forceinline__ uint32_t search_global(const CdfView& v, const DevScalars* ds, const P gt) {
  uint64_t off = 0;                               // integer weight of the ranks before r: C_i = off + sp[segment] + cl[i]
  int r = 0;
  for (; r < v.nranks - 1; ++r) {
    const uint64_t end = off + ds->cdf_rank_total[r];
    if (gt(end)) break;
    off = end;
  }
  const uint64_t* sp = v.sp[r];                   // rank-local exclusive prefixes of rank r's segments
  const int s = WARP ? upper_pred_warp(sp + 1, v.n_segs, off, gt) : upper_pred(sp + 1, v.n_segs, off, gt);
  int64_t j = v.n_per - 1;
  if (s < v.n_segs) {
    const int64_t first = (int64_t)s * v.seg_len;
    const int len = (int)(first + v.seg_len <= v.n_pad ? v.seg_len : v.n_pad - first);
    const uint64_t add = off + __ldg(sp + s);
    j = first + (WARP ? upper_pred_warp(v.seg[r] + first, len, add, gt) : upper_pred(v.seg[r] + first, len, add, gt));
    if (j > v.n_per - 1) j = v.n_per - 1;
  }
  return ((uint32_t)r << GSMC_ANC_RANK_SHIFT) | (uint32_t)j;
}

// Sorted mode, step 1: ancestor word of the order statistic that OPENS every tile, win[b] for b in [0, nt]; win[nt]
// closes the last tile. One WARP per boundary (32-ary searches: 2 probe rounds over the segment prefixes
// in shared memory, 3 over the segment in global/peer memory), 32 boundaries per 1024-thread block. Also
// turns tile_e[b] into the GLOBAL gap prefix A (head included) at which tile b opens.
// FUSED_SCAN (multinomial, Philox draws): every block first scans the raw segment totals itself (identical
// results in every block), which replaces the separate scan launch; block 0 publishes prefixes and totals and, on a
// sharded filter, sends this rank's totals to the peers' mailboxes (every block reads all ranks' totals back).
template <bool FUSED_SCAN>
__global__ void __launch_bounds__(1024) partition_kernel(CdfView v, uint64_t k_first, int rank, DevScalars* ds,
                                                         const uint64_t* raw_q, const uint64_t* raw_e, uint64_t* sp_q, uint64_t* sp_e,
                                                         uint64_t* tile_e, int seg_tiles, int nt, uint32_t* win,
                                                         uint64_t n_global, int conditional, PeerScalars peers) {
  __shared__ uint64_t spq[GSMC_MAX_SEGS + 1];
  __shared__ uint64_t spe[GSMC_MAX_SEGS + 1];
  __shared__ uint64_t sm[2][33];
  __shared__ double s_ratio;
  __shared__ uint64_t s_draws, s_cn, s_head;
  __shared__ uint64_t off_q[GSMC_MAX_RANKS + 1], off_e[GSMC_MAX_RANKS + 1];   // exclusive prefixes of the ranks' totals, [R] = sum
  __shared__ uint64_t tot_q[GSMC_MAX_RANKS], tot_e[GSMC_MAX_RANKS];
  // The decision was final before the weights pass of this event triggered this launch (it waited for the deciding kernel
  // first), so a step that does not resample leaves WITHOUT waiting for the previous kernel to drain.
  if (conditional && !ds->do_resample) { pdl_trigger(); return; }   // every rank takes the same decision: nobody sends, nobody waits
  pdl_wait();
  pdl_trigger();
  const int n_segs = v.n_segs, R = v.nranks;
  if (FUSED_SCAN) {
    const uint32_t seq = ds->xseq + 1;                    // the totals exchange of this step (see DevScalars::xseq)
    uint64_t totals[2];
    scan_segments_block(raw_q, raw_e, n_segs, spq, spe, sm, totals);
    if (R > 1) {
      // Block 0 publishes this rank's (rank-local) segment prefixes and then sends its two totals to every peer;
      // every block of every rank reads all totals from its own mailboxes. A peer that has seen rank r's totals
      // may read r's prefixes and CDF: they were complete (and fenced) before the send.
      if (blockIdx.x == 0) {
        if (threadIdx.x <= n_segs) { sp_q[threadIdx.x] = spq[threadIdx.x]; sp_e[threadIdx.x] = spe[threadIdx.x]; }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < R) {
          const uint32_t w[4] = {(uint32_t)totals[0], (uint32_t)(totals[0] >> 32), (uint32_t)totals[1], (uint32_t)(totals[1] >> 32)};
          __threadfence_system();
          ll_send(peers.ds[threadIdx.x], rank, seq, w, 4);
        }
      }
      if (threadIdx.x < R) {
        uint32_t g[4];
        if (!ll_recv(ds, threadIdx.x, seq, g, 4)) { ds->error = 2; g[0] = g[1] = g[2] = g[3] = 0; }
        tot_q[threadIdx.x] = (uint64_t)g[0] | ((uint64_t)g[1] << 32);
        tot_e[threadIdx.x] = (uint64_t)g[2] | ((uint64_t)g[3] << 32);
      }
    } else if (threadIdx.x == 0) { tot_q[0] = totals[0]; tot_e[0] = totals[1]; }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t aq = 0, ae = 0;
      for (int r = 0; r < R; ++r) { off_q[r] = aq; off_e[r] = ae; aq += tot_q[r]; ae += tot_e[r]; }
      off_q[R] = aq; off_e[R] = ae;
      const uint64_t head = ds->gap_head;                // drawn by block 0 of the gap pass (same value on every rank)
      const uint64_t stot = head + ae;
      s_ratio = stot ? (double)aq / (double)stot : 0.0;
      s_draws = n_global; s_cn = aq; s_head = head;
      if (blockIdx.x == 0) {
        for (int r = 0; r < R; ++r) { ds->cdf_rank_total[r] = tot_q[r]; ds->gap_rank_total[r] = tot_e[r]; }
        ds->cdf_total = aq; ds->n_draws = n_global; ds->n_det = 0;
        ds->gap_total = stot; ds->thr_ratio = s_ratio;
      }
    }
    __syncthreads();
    if (R == 1 && blockIdx.x == 0 && threadIdx.x <= n_segs) { sp_q[threadIdx.x] = spq[threadIdx.x]; sp_e[threadIdx.x] = spe[threadIdx.x]; }
  } else {
    if (threadIdx.x <= n_segs) { spq[threadIdx.x] = sp_q[threadIdx.x]; spe[threadIdx.x] = sp_e[threadIdx.x]; }
    if (threadIdx.x == 0) {
      s_ratio = ds->thr_ratio; s_draws = ds->n_draws; s_cn = ds->cdf_total; s_head = ds->gap_head;
      uint64_t aq = 0, ae = 0;
      for (int r = 0; r < R; ++r) { off_q[r] = aq; off_e[r] = ae; aq += ds->cdf_rank_total[r]; ae += ds->gap_rank_total[r]; }
      off_q[R] = aq; off_e[R] = ae;
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const uint64_t m_draws = s_draws, cn = s_cn;
  const double ratio = s_ratio;
  const uint64_t my_e = s_head + off_e[rank];
  for (int b = blockIdx.x * 32 + (threadIdx.x >> 5); b <= nt; b += gridDim.x * 32) {
    const uint64_t kt = k_first + (uint64_t)b * GSMC_TILE;
    uint32_t w = ((uint32_t)(R - 1) << GSMC_ANC_RANK_SHIFT) | (uint32_t)(v.n_per - 1);
    uint64_t A;                                            // the tile opens at the order statistic A / S_tot
    if (b < nt) {
      A = my_e + spe[b / seg_tiles] + tile_e[b];
      __syncwarp();
      if (lane == 0) tile_e[b] = A;
    } else {
      A = my_e + spe[n_segs];                              // the closing boundary: first draw of the next rank
    }
    if (kt < m_draws) {
      GtU64 gt; gt.T = threshold_u64((double)A, ratio, cn);
      int r = 0;
      for (; r < R - 1; ++r) if (gt(off_q[r + 1])) break;
      const uint64_t off = off_q[r];
      const uint64_t* sp = (r == rank) ? spq : v.sp[r];      // own prefixes from shared memory, a peer's over NVLink
      const int sg = upper_pred_warp(sp + 1, n_segs, off, gt);
      int64_t j = v.n_per - 1;
      if (sg < n_segs) {
        const int64_t first = (int64_t)sg * v.seg_len;
        const int len = (int)(first + v.seg_len <= v.n_pad ? v.seg_len : v.n_pad - first);
        j = first + upper_pred_warp(v.seg[r] + first, len, off + sp[sg], gt);
        if (j > v.n_per - 1) j = v.n_per - 1;
      }
      w = ((uint32_t)r << GSMC_ANC_RANK_SHIFT) | (uint32_t)j;
    }
    if (lane == 0) win[b] = w;
  }
}

// ---- mbarrier + 1-D bulk copy (TMA) helpers: one thread moves a contiguous window global -> shared ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
// dst, src 16-byte aligned; bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) {
  uint64_t v;
  asm("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}

// Key of a CDF entry inside a group's bracket [TL, TH): (nearly) the largest 32-bit word that still selects a particle
// at or before it, i.e. (C - TL) 2^32 / (TH - TL), in integer arithmetic (no int<->fp64 conversion per entry): with
// D = TH - TL normalised to 32 bits, Dn = top 32 bits of D << clz(D) (in [2^31, 2^32)), the multiplier
// M = trunc((2^63 - 2^10) / Dn) (in [2^31, 2^32), one fp64 division per group) and c' = top 32 bits of (C - TL) << clz(D)
// (<= Dn), the key is (c' M) >> 31 < 2^32: non-decreasing in C, off the exact quotient by a few units of 2^-32.
struct BracketScale { uint32_t lz, mul; };
__device__ __forceinline__ BracketScale bracket_scale(uint64_t tl, uint64_t th) {
  BracketScale b; b.lz = 0; b.mul = 0;                          // TH <= TL: every key is 0
  if (th > tl) {
    const uint64_t d = th - tl;
    b.lz = (uint32_t)__clzll((long long)d);
    const uint32_t dn = (uint32_t)((d << b.lz) >> 32);
    b.mul = __double2uint_rz(9223372036854774784.0 / (double)dn);
  }
  return b;
}
__device__ __forceinline__ uint32_t bracket_key(uint64_t c_minus_tl, BracketScale b) {
  const uint32_t cn = (uint32_t)((c_minus_tl << b.lz) >> 32);
  return (uint32_t)(((uint64_t)cn * b.mul) >> 31);
}
// global CDF value of the particle behind an ancestor word
__device__ __forceinline__ uint64_t cdf_at(const CdfView& v, const DevScalars* ds, uint32_t word) {
  const int r = (int)(word >> GSMC_ANC_RANK_SHIFT);
  const uint32_t i = word & GSMC_ANC_INDEX_MASK;
  return rank_offset(ds, r) + __ldg(v.sp[r] + i / (uint32_t)v.seg_len) + __ldg(v.seg[r] + i);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// one step of a branch-free binary search over 4-byte keys in shared memory: if (key[ad + probe] < w) ad += step, as a
// load, a compare and a PREDICATED add (the compiler's own choice is compare + select + add)
__device__ __forceinline__ void search_step(uint32_t& ad, uint32_t probe, uint32_t step, uint32_t w) {
  asm("{\n .reg .pred p;\n .reg .u32 k, a;\n add.u32 a, %0, %1;\n ld.shared.u32 k, [a];\n setp.lt.u32 p, k, %2;\n @p add.u32 %0, %0, %3;\n}"
      : "+r"(ad) : "r"(probe), "r"(w), "r"(step));
}
// Bracket of a group on the global CDF, as global positions (rank * n_per + index): low word p_lo, high word p_hi.
__device__ __noinline__ uint64_t bracket_global(const CdfView& v, const DevScalars* ds, uint64_t tl, uint64_t th) {
  GtU64 g_lo; g_lo.T = tl;
  GtU64 g_hi; g_hi.T = th;
  const uint32_t wl = search_global<false>(v, ds, g_lo), wh = search_global<false>(v, ds, g_hi);
  const uint64_t p_lo = (uint64_t)(wl >> GSMC_ANC_RANK_SHIFT) * (uint64_t)v.n_per + (wl & GSMC_ANC_INDEX_MASK);
  const uint64_t p_hi = (uint64_t)(wh >> GSMC_ANC_RANK_SHIFT) * (uint64_t)v.n_per + (wh & GSMC_ANC_INDEX_MASK);
  return p_lo | (p_hi << 32);
}
// Ancestor word of the draw with Philox word w in the bracket [p_lo, p_hi): p_lo + #{p : K_p < w}, keys evaluated per probe.
__device__ __noinline__ uint32_t draw_global(const CdfView& v, const DevScalars* ds, uint32_t p_lo, uint32_t p_hi, uint64_t tl, BracketScale r32, uint32_t w) {
  uint32_t l = p_lo, h = p_hi;
  const uint32_t n_per = (uint32_t)v.n_per;
  while (l < h) {
    const uint32_t mid = l + ((h - l) >> 1);
    const uint32_t word = ((mid / n_per) << GSMC_ANC_RANK_SHIFT) | (mid % n_per);
    if (bracket_key(cdf_at(v, ds, word) - tl, r32) < w) l = mid + 1; else h = mid;
  }
  return ((l / n_per) << GSMC_ANC_RANK_SHIFT) | (l % n_per);
}

// Sorted mode, step 2. One block iteration handles a tile of 2048 consecutive output slots = 8 groups, one group
// (256 slots, 8 per lane) per warp, and writes their ancestors.
//   1. one thread starts a bulk copy (TMA, mbarrier-tracked) of the CDF window [win[m], win[m+1]] the tile can map to
//      into shared memory; while it is in flight every thread draws its 8 Philox words (2 calls);
//   2. the window stays in the integer form the weights pass wrote (segment-local u64 values): the segment prefix and
//      rank offset are subtracted from the THRESHOLDS instead of being added to every window entry (windows that
//      straddle a segment or rank boundary are rebased once in shared memory);
//   3. every warp locates the order statistic that opens its group in the window (one 32-ary cooperative search; the
//      closing one is the next warp's) and turns the entries of its bracket, in place, into 32-bit keys in the domain of
//      the Philox words (one conversion per CDF entry, nothing per draw);
//   4. every lane runs 8 interleaved, bound-check-free binary searches over the keys of the bracket (same probe count
//      for the whole warp, no divergence, 8 independent 4-byte shared-memory load chains per lane);
//   5. positions -> ancestor words, 32-byte vector stores.
// Windows that span more than two ranks or exceed the shared-memory capacity fall back to per-draw searches in global
// memory (same definition); windows that touch a peer's CDF are staged with ordinary loads.
#define GSMC_SEARCH_TPT 8                                   // draws per thread
#ifndef GSMC_WIN_CAP
#define GSMC_WIN_CAP 5120
#endif
#ifndef GSMC_SEARCH_OCC
#define GSMC_SEARCH_OCC 5
#endif
#define GSMC_SEARCH_SMEM ((GSMC_WIN_CAP + 4) * 8)           // 41 KB of window per block: 5 blocks per SM
__global__ void __launch_bounds__(GSMC_BLOCK, GSMC_SEARCH_OCC) search_sorted_kernel(CdfView v, uint64_t k_first, const DevScalars* ds,
                                                                      const uint64_t* tile_e, const uint64_t* gap, uint32_t seg_tiles,
                                                                      uint32_t seg_magic, const PhiloxKeys keys, const uint32_t* win,
                                                                      uint32_t* anc, int64_t n_out, int nt, int det_offset,
                                                                      int conditional, int rank, AncOut anc_all) {
  extern __shared__ __align__(16) uint64_t cwin[];                 // GSMC_WIN_CAP + 4
  __shared__ uint64_t mbar;
  __shared__ int s_pos[GSMC_GPT + 1];                              // window positions of the order statistics that bracket the tile's groups
  if (conditional && !ds->do_resample) { pdl_trigger(); return; }   // (see partition_kernel: no wait on a step that does not resample)
  pdl_wait();
  // No early trigger here: the kernel that follows is the next propagate, a persistent grid of 80-register blocks that
  // would become resident as soon as search blocks retire and take their place while the rest of this grid still needs
  // the SMs (measured, cfg 3: 20.47 ms per run with an early trigger, 19.52 with the implicit one at block exit).
  const uint64_t m_draws = ds->n_draws, cn = ds->cdf_total;
  const double ratio = ds->thr_ratio;
  const uint32_t rho = ds->rho;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) mbar_init(&mbar, 1);
  __syncthreads();
  uint32_t phase = 0;
  for (int m = blockIdx.x; m < nt; m += gridDim.x) {
    const uint64_t kt = k_first + (uint64_t)m * GSMC_TILE;
    if (kt >= m_draws) break;                            // uniform per block
    const int64_t o_local = (int64_t)m * GSMC_TILE + GSMC_SEARCH_TPT * threadIdx.x;
    // window [w0, w1]: [lo, hi] of rank r0, or the tail [lo, n_per) of r0 followed by the head [0, hi] of r0+1
    const uint32_t w0 = win[m], w1 = win[m + 1];
    const int r0 = (int)(w0 >> GSMC_ANC_RANK_SHIFT), r1 = (int)(w1 >> GSMC_ANC_RANK_SHIFT);
    const int lo = (int)(w0 & GSMC_ANC_INDEX_MASK), hi = (int)(w1 & GSMC_ANC_INDEX_MASK);
    const int64_t len_a = (r1 == r0) ? (int64_t)hi - lo + 1 : v.n_per - lo;
    const int64_t len_b = (r1 == r0) ? 0 : (int64_t)hi + 1;
    const bool staged = (r1 == r0 || r1 == r0 + 1) && len_a + len_b <= GSMC_WIN_CAP;
    const bool bulk = staged && r1 == r0 && r0 == rank;
    const int la = (int)len_a, len = (int)(len_a + len_b);
    const int off = bulk ? (lo & 1) : 0;                 // the bulk copy starts at an even (16-byte aligned) element
    if (bulk && threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)((off + len + 1) & ~1) * 8u;
      fence_proxy_async();                               // earlier generic-proxy accesses of cwin come first
      mbar_expect_tx(&mbar, bytes);
      bulk_g2s(cwin, v.seg[r0] + (lo - off), bytes, &mbar);
    }
    // The group of this warp opens at A_w = A_tile + gaps of the tile's earlier groups and spans g_w.
    uint64_t gme = (lane < GSMC_GPT) ? __ldg(gap + (int64_t)m * GSMC_GPT + lane) : 0;
    uint64_t gin = gme;
#pragma unroll
    for (int d = 1; d < GSMC_GPT; d <<= 1) { const uint64_t y = shfl_up_u64(gin, d); if (lane >= d) gin += y; }
    const uint64_t A_tile = __ldg(tile_e + m);
    const uint64_t g_w = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)gme, warp);
    const uint64_t A_w = A_tile + (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)gin, warp) - g_w;
    const uint64_t TL_abs = threshold_u64((double)A_w, ratio, cn), TH_abs = threshold_u64((double)(A_w + g_w), ratio, cn);
    // words of this lane: slots 8 lane .. 8 lane + 7 of the group; slot 0 of the group is the order statistic itself
    const uint64_t k = k_first + (uint64_t)o_local;
    const PhiloxOut p0 = philox_call(keys, k >> 2, rho, GSMC_STREAM_RESAMPLE), p1 = philox_call(keys, (k >> 2) + 1, rho, GSMC_STREAM_RESAMPLE);
    const uint32_t wd[GSMC_SEARCH_TPT] = {(uint32_t)p0.a, (uint32_t)(p0.a >> 32), (uint32_t)p0.b, (uint32_t)(p0.b >> 32),
                                          (uint32_t)p1.a, (uint32_t)(p1.a >> 32), (uint32_t)p1.b, (uint32_t)(p1.b >> 32)};
    uint32_t a[GSMC_SEARCH_TPT];
    if (staged) {
      // C_i = rank offset + sp[segment(i)] + cl[i] with rank-local segment prefixes; segment(i) = (i / GSMC_TILE) / seg_tiles by multiply-high
      const uint64_t* sp_a = v.sp[r0];
      const uint32_t tl0 = (uint32_t)lo >> GSMC_TILE_SHIFT, tl1 = (uint32_t)(lo + la - 1) >> GSMC_TILE_SHIFT;
      const uint32_t sg0 = seg_tiles == 1 ? tl0 : __umulhi(tl0, seg_magic), sg1 = seg_tiles == 1 ? tl1 : __umulhi(tl1, seg_magic);
      const uint64_t sp0 = __ldg(sp_a + sg0);
      const uint64_t base = rank_offset(ds, r0) + sp0;   // window entries are kept relative to `base`: C_i - base
      if (bulk) {
        mbar_wait(&mbar, phase);
        phase ^= 1;
        if (sg0 != sg1) {                                  // rare: the window straddles a segment boundary of this rank
#pragma unroll 4
          for (int j = threadIdx.x; j < la; j += GSMC_BLOCK) {
            const uint32_t tl = (uint32_t)(lo + j) >> GSMC_TILE_SHIFT;
            const uint32_t sg = seg_tiles == 1 ? tl : __umulhi(tl, seg_magic);
            cwin[off + j] += __ldg(sp_a + sg) - sp0;
          }
          __syncthreads();
        }
      } else {
        const uint64_t* seg_a = v.seg[r0] + lo;
#pragma unroll 4
        for (int j = threadIdx.x; j < la; j += GSMC_BLOCK) {
          const uint32_t tl = (uint32_t)(lo + j) >> GSMC_TILE_SHIFT;
          const uint32_t sg = seg_tiles == 1 ? tl : __umulhi(tl, seg_magic);
          cwin[j] = (__ldg(sp_a + sg) - sp0) + __ldg(seg_a + j);
        }
        if (len_b) {
          const uint64_t* seg_b = v.seg[r1];
          const uint64_t* sp_b = v.sp[r1];
          const uint64_t roff_b = rank_offset(ds, r1) - base;       // >= 0: rank r1 starts where r0's total ends
          for (int j = threadIdx.x; j < (int)len_b; j += GSMC_BLOCK) {
            const uint32_t tl = (uint32_t)j >> GSMC_TILE_SHIFT;
            const uint32_t sg = seg_tiles == 1 ? tl : __umulhi(tl, seg_magic);
            cwin[la + j] = roff_b + __ldg(sp_b + sg) + __ldg(seg_b + j);
          }
        }
        __syncthreads();
      }
      // bracket of this warp's group: position of its opening order statistic (warp-uniform); the closing one is the
      // next warp's opening one (the last warp also looks up the tile's closing one)
      uint64_t* cw = cwin + off;
      const uint64_t TL = TL_abs - base, TH = TH_abs - base;
      {
        GtU64 g_lo; g_lo.T = TL;
        const int p = upper_pred_warp(cw, len, 0, g_lo);
        if (lane == 0) s_pos[warp] = p;
        if (warp == GSMC_GPT - 1) {
          GtU64 g_hi; g_hi.T = TH;
          const int q = upper_pred_warp(cw, len, 0, g_hi);
          if (lane == 0) s_pos[GSMC_GPT] = q;
        }
      }
      __syncthreads();                                     // all bracket searches read the u64 entries: done before any key is written
      const int p_lo = s_pos[warp];
      int p_hi = s_pos[warp + 1];
      p_hi = p_hi < len - 1 ? p_hi : len - 1;
      // keys of the bracket, in place (low word of each 8-byte entry); brackets of different warps are disjoint
      const BracketScale r32 = bracket_scale(TL, TH);
      for (int p = p_lo + lane; p < p_hi; p += 32) {
        const uint32_t key = bracket_key(cw[p] - TL, r32);
        *reinterpret_cast<uint32_t*>(cw + p) = key;
      }
      __syncwarp();
      // pos_j = p_lo + #{p in [p_lo, p_hi) : K_p < w_j}: every probe is in bounds and the probe count depends on
      // the bracket length only (32-bit shared-memory addresses: one add per step)
      uint32_t cw_base = (uint32_t)__cvta_generic_to_shared(cw);
      asm volatile("" : "+r"(cw_base) :: "memory");        // the probes below (plain asm loads) stay behind the key stores
      uint32_t ad[GSMC_SEARCH_TPT];
#pragma unroll
      for (int j = 0; j < GSMC_SEARCH_TPT; ++j) ad[j] = cw_base + (uint32_t)p_lo * 8u;
      // (halving by n/2, n/4, ... rather than a power-of-two ladder with immediate offsets: power-of-two strides put
      // every lane's probe into the same shared-memory bank, measured 143 us against 107 us)
      const int n_search = __shfl_sync(0xffffffffu, p_hi - p_lo, 0);     // warp-uniform (broadcast: lets the loop control run on the uniform datapath)
      for (int rem = n_search; rem > 1;) {
        const int half = rem >> 1;
        const uint32_t probe = (uint32_t)(half - 1) * 8u, step = (uint32_t)half * 8u;
#pragma unroll
        for (int j = 0; j < GSMC_SEARCH_TPT; ++j) search_step(ad[j], probe, step, wd[j]);
        rem -= half;
      }
      if (n_search > 0) {
#pragma unroll
        for (int j = 0; j < GSMC_SEARCH_TPT; ++j) search_step(ad[j], 0u, 8u, wd[j]);
      }
      if (lane == 0) ad[0] = cw_base + (uint32_t)p_lo * 8u;        // the order statistic that opens the group
      if (len_b == 0) {                                    // the usual case (block-uniform): the window lies in one rank
#pragma unroll
        for (int j = 0; j < GSMC_SEARCH_TPT; ++j) a[j] = w0 + ((ad[j] - cw_base) >> 3);
      } else {
#pragma unroll
        for (int j = 0; j < GSMC_SEARCH_TPT; ++j) {
          const int pj = (int)((ad[j] - cw_base) >> 3);
          a[j] = pj < la ? (w0 + (uint32_t)pj) : ((((uint32_t)r1) << GSMC_ANC_RANK_SHIFT) | (uint32_t)(pj - la));
        }
      }
    } else {
      // same definition on the global CDF (out of line: rare, and the hot loop stays small)
      const uint64_t br = bracket_global(v, ds, TL_abs, TH_abs);
      const BracketScale r32 = bracket_scale(TL_abs, TH_abs);
#pragma unroll
      for (int j = 0; j < GSMC_SEARCH_TPT; ++j)
        a[j] = draw_global(v, ds, (uint32_t)br, (j == 0 && lane == 0) ? (uint32_t)br : (uint32_t)(br >> 32), TL_abs, r32, wd[j]);
    }
    // output slot of draw k: k - k_first (multinomial: this rank's own slots); residual scheme: GLOBAL slot n_det + k,
    // stored into the rank that owns it
    if (!det_offset) {
      const int64_t o = o_local;
      if (k + GSMC_SEARCH_TPT <= m_draws && o + GSMC_SEARCH_TPT <= n_out) {
        *reinterpret_cast<uint4*>(anc + o) = make_uint4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<uint4*>(anc + o + 4) = make_uint4(a[4], a[5], a[6], a[7]);
      } else {
#pragma unroll
        for (int j = 0; j < GSMC_SEARCH_TPT; ++j) if (k + j < m_draws && o + j < n_out) anc[o + j] = a[j];
      }
    } else {
      const uint64_t og = ds->n_det + k;
#pragma unroll
      for (int j = 0; j < GSMC_SEARCH_TPT; ++j) if (k + j < m_draws) anc_store(anc_all, og + j, a[j]);
    }
    __syncthreads();                                     // everybody is done with the window before the next tile's copy lands
  }
}

// iid mode (replay / sample_unweighted): T_j = floor(floor(u_j 2^53) * C_N / 2^53)
__global__ void __launch_bounds__(GSMC_BLOCK) search_iid_kernel(CdfView v, const DevScalars* ds, const double* u, uint64_t seed,
                                                                uint32_t event, uint32_t stream, int64_t m, int64_t out_offset_det,
                                                                uint32_t* anc32, int64_t* anc64, int conditional) {
  if (conditional && !ds->do_resample) return;
  const int64_t j = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  int64_t count = m;
  if (out_offset_det) count = (int64_t)ds->n_draws;
  if (j >= count) return;
  double uj;
  if (u) uj = u[j];
  else { double u0, u1; uniform_pair(seed, (uint64_t)j >> 1, event, stream, &u0, &u1); uj = (j & 1) ? u1 : u0; }
  const uint64_t t = (uint64_t)floor(uj * 9007199254740992.0);
  const uint64_t cn = ds->cdf_total;
  GtU64 gt; gt.T = (__umul64hi(t, cn) << 11) | ((t * cn) >> 53);
  const uint32_t w = search_global<false>(v, ds, gt);
  const int64_t o = j + (out_offset_det ? (int64_t)ds->n_det : 0);
  if (anc32) anc32[o] = w;
  if (anc64) anc64[o] = (int64_t)(w >> GSMC_ANC_RANK_SHIFT) * v.n_per + (int64_t)(w & GSMC_ANC_INDEX_MASK);
}

// residual scheme, deterministic part: particle i owns the output slots [Cc_{i-1}, Cc_i) of the copies CDF (two levels:
// segment prefix + segment-local inclusive count). A scatter: one thread per particle writes its own c_i = floor(N p_i)
// copies (usually 0..2, neighbouring threads write neighbouring slots); a particle with 32 or more copies is written by
// its whole warp. Same result as searching min{i : Cc_i > o} for every slot o.
// Sharded filter: the copies of rank r's particles start at global slot D_r = copies of the lower ranks, and a slot is
// stored into the ancestor column of the rank that owns it (peer memory); the word names the owner rank of the particle.
__global__ void __launch_bounds__(GSMC_BLOCK) det_copies_kernel(const uint64_t* cc, const uint64_t* seg_c, int n_segs, int seg_len, int n_pad,
                                                                int64_t n, int64_t n_global, int rank, const DevScalars* ds, AncOut anc, int conditional) {
  pdl_wait();
  pdl_trigger();
  if (conditional && !ds->do_resample) return;
  const int lane = threadIdx.x & 31;
  uint64_t rank_off = 0;
  for (int q = 0; q < rank; ++q) rank_off += ds->det_rank_total[q];
  const uint32_t rank_word = (uint32_t)rank << GSMC_ANC_RANK_SHIFT;
  for (int64_t base = ((int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x) & ~(int64_t)31; base < n_pad; base += (int64_t)gridDim.x * GSMC_BLOCK) {
    const int64_t i = base + lane;                       // a warp covers 32 consecutive particles of ONE segment
    const int seg = (int)(base / seg_len);
    const uint64_t off = rank_off + __ldg(seg_c + seg);
    const uint64_t incl = (i < n) ? __ldg(cc + i) : 0;
    uint64_t prev = shfl_up_u64(incl, 1);
    if (lane == 0) prev = (base % seg_len == 0) ? 0 : __ldg(cc + base - 1);
    uint64_t first = off + prev;
    uint64_t cnt = (i < n) ? incl - prev : 0;
    if (first >= (uint64_t)n_global) cnt = 0;
    else if (first + cnt > (uint64_t)n_global) cnt = (uint64_t)n_global - first;       // never more than N slots
    const bool big = cnt >= 32;
    if (!big) for (uint64_t j = 0; j < cnt; ++j) anc_store(anc, first + j, rank_word | (uint32_t)i);
    unsigned mask = __ballot_sync(0xffffffffu, big);
    while (mask) {
      const int src = __ffs((int)mask) - 1;
      mask &= mask - 1;
      const uint64_t f0 = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)first, src);
      const uint64_t c0 = (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)cnt, src);
      const uint32_t who = rank_word | (uint32_t)(base + src);
      for (uint64_t j = lane; j < c0; j += 32) anc_store(anc, f0 + j, who);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
// importance.jl:31,50: log_normalized_weights = log_weights .- log_total_weight
// The rank's (max, s1, s2) triple is shifted with the weights (thread 0), so a later statistics pass -- e.g. the draw of
// sample_unweighted_traces / importance_resampling -- sees the maximum of the NORMALISED log weights.
template <typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK) normalize_lw_kernel(Real* lw, int64_t n, DevScalars* ds, int rank) {
  const int64_t i = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  const double log_total = ds->log_total;
  if (i < n) lw[i] = (Real)((double)lw[i] - log_total);
  if (i == 0) { ds->triples[rank].m -= log_total; ds->max_lw -= log_total; }
}
template <typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK) column_to_f64_kernel(const Real* src, double* dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}

// get_traces view: value of time step t for the particle that is now at position i.
// Walks the per-step ancestor columns back from the newest (anc_cols[s] was applied when step s
// was produced from step s-1, if resampled[s]).
template <typename Real>
struct HistView {
  const Real* slab[GSMC_MAX_RANKS];      // [cap][D][stride] per rank
  const uint32_t* anc_slab[GSMC_MAX_RANKS];
  const int* resampled;                  // [cap + 2], index = step
  int64_t stride;
  int D;
  int64_t cap;                           // columns in the slabs; step t lives in column (t-1) % cap
  int64_t flag_mod;                      // resampled[] is indexed by step % flag_mod
};
template <typename Real>
__device__ __forceinline__ uint32_t walk_back(const HistView<Real>& h, uint32_t word, int64_t from_step, int64_t to_step) {
  // word = (rank, index) position in the ordering that exists after step `from_step` ... down to `to_step`
  for (int64_t s = from_step; s > to_step; --s) {
    if (h.resampled[s % h.flag_mod]) {
      const uint32_t* a = h.anc_slab[word >> GSMC_ANC_RANK_SHIFT] + ((s - 1) % h.cap) * h.stride;
      word = a[word & GSMC_ANC_INDEX_MASK];
    }
  }
  return word;
}
// out[d][i] (column-major, ld = n) of time step t for local particles i in [0, n); newest = newest existing step,
// pending = 1 if a resample has been decided for step newest+1 but not yet applied
template <typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK) get_state_kernel(HistView<Real> h, int rank, int64_t n, int64_t t, int64_t newest,
                                                               int pending, double* out) {
  const int64_t i = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  if (i >= n) return;
  uint32_t word = ((uint32_t)rank << GSMC_ANC_RANK_SHIFT) | (uint32_t)i;
  word = walk_back(h, word, newest + (pending ? 1 : 0), t);
  const Real* col = h.slab[word >> GSMC_ANC_RANK_SHIFT] + ((t - 1) % h.cap) * h.D * h.stride + (word & GSMC_ANC_INDEX_MASK);
  for (int d = 0; d < h.D; ++d) out[(int64_t)d * n + i] = (double)col[d * h.stride];
}
// out[s][t-1][d] for selected local particles idx[s]
template <typename Real>
__global__ void __launch_bounds__(GSMC_BLOCK) trajectories_kernel(HistView<Real> h, int64_t n_per, const int64_t* idx, int64_t n_idx,
                                                                  int64_t newest, int pending, double* out) {
  const int64_t s = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  if (s >= n_idx) return;
  // idx holds GLOBAL particle indices: owner rank = idx / n_per; rows of other ranks are read through the peer mappings
  const int owner = (int)(idx[s] / n_per);
  uint32_t word = ((uint32_t)owner << GSMC_ANC_RANK_SHIFT) | (uint32_t)(idx[s] - (int64_t)owner * n_per);
  if (pending && h.resampled[(newest + 1) % h.flag_mod]) {
    const uint32_t* a = h.anc_slab[owner] + (newest % h.cap) * h.stride;
    word = a[word & GSMC_ANC_INDEX_MASK];
  }
  for (int64_t t = newest; t >= 1; --t) {
    const Real* col = h.slab[word >> GSMC_ANC_RANK_SHIFT] + ((t - 1) % h.cap) * h.D * h.stride + (word & GSMC_ANC_INDEX_MASK);
    for (int d = 0; d < h.D; ++d) out[(s * newest + (t - 1)) * h.D + d] = (double)col[d * h.stride];
    if (t > 1 && h.resampled[t % h.flag_mod]) {
      const uint32_t* a = h.anc_slab[word >> GSMC_ANC_RANK_SHIFT] + ((t - 1) % h.cap) * h.stride;
      word = a[word & GSMC_ANC_INDEX_MASK];
    }
  }
}
// ancestor words -> global int64 indices (state.parents)
__global__ void __launch_bounds__(GSMC_BLOCK) anc_to_global_kernel(const uint32_t* anc, int64_t n, int64_t n_per, int64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * GSMC_BLOCK + threadIdx.x;
  if (i < n) out[i] = (int64_t)(anc[i] >> GSMC_ANC_RANK_SHIFT) * n_per + (anc[i] & GSMC_ANC_INDEX_MASK);
}

#endif
