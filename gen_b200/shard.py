"""Host-side sharding rules of the multi-GPU filter; pure functions so that the
world_size>1 logic is testable on CPU with gloo.

  * rank r of R owns the contiguous global particle indices [r*N/R, (r+1)*N/R); Philox counters use
    the GLOBAL index, so draws do not depend on R.
  * logsumexp/ESS: every rank contributes (max, sum exp(lw-max), sum exp(2(lw-max))); the triples are
    allgathered and merged in rank order on every rank (deterministic, identical everywhere).
  * resampling: integer weights make the prefix sum associative; the global CDF is the per-rank local
    CDFs offset by the exclusive prefix of the allgathered per-rank totals. Output slot k (global) of the
    sorted-uniform scheme belongs to the rank that owns index k; its ancestor may live on any rank. The sorted
    uniforms are grouped order statistics (GROUP slots per group): a rank draws the Gamma gaps of its own groups,
    the ranks exchange the gap totals, and group j opens at (head + gaps of all groups before j) / S_tot.
"""
import math

TILE = 2048
GROUP = 256          # output slots per group of the sorted draws (GSMC_GROUP in gsmc_rng.cuh)


def check_partition(num_particles, world_size):
    if world_size < 1 or world_size > 8:
        raise ValueError("1 <= world_size <= 8")
    if world_size > 1 and num_particles % (world_size * TILE) != 0:
        raise ValueError("num_particles must be a multiple of %d * world_size" % TILE)


def partition(num_particles, world_size, rank):
    """(first_global, count) of a rank."""
    check_partition(num_particles, world_size)
    n = num_particles // world_size
    return rank * n, n


def merge_lse(a, b):
    """Merge two (m, s1, s2) triples (same rule as lse_merge in kernels.cuh)."""
    if a[0] == -math.inf and not math.isnan(a[1]):
        return b
    if b[0] == -math.inf and not math.isnan(b[1]):
        return a
    m = max(a[0], b[0])
    ea, eb = math.exp(a[0] - m), math.exp(b[0] - m)
    return (m, a[1] * ea + b[1] * eb, a[2] * ea * ea + b[2] * eb * eb)


def combine_lse(triples):
    """(log_total, ess) from the rank-ordered triples."""
    t = triples[0]
    for x in triples[1:]:
        t = merge_lse(t, x)
    log_total = t[0] + math.log(t[1])
    ess = math.exp(-(2.0 * (t[0] - log_total) + math.log(t[2])))
    return log_total, ess


def cdf_offsets(rank_totals):
    """Exclusive prefix of the per-rank integer weight totals and the grand total C_N."""
    offs, acc = [], 0
    for w in rank_totals:
        offs.append(acc)
        acc += int(w)
    return offs, acc


def owner_of_ancestor(global_index, num_particles, world_size):
    n = num_particles // world_size
    return global_index // n, global_index % n
