"""Multi-GPU path. GPU part: torchrun with 2 (and 4/8 if present) ranks, every rank checks its shard
against the full CPU oracle bit for bit. CPU part: the sharding rules (gen_b200/shard.py) under a
world_size-2 gloo group reproduce the single-rank oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_filter_matches_oracle(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count(" ok on %d ranks" % world) == 8, res.stdout[-2000:]


@pytest.mark.gpu
def test_sharded_filter_with_nccl_scalar_exchanges():
    """GSMC_NCCL_SCALARS=1: the per-step scalars travel by ncclAllGather instead of the fused peer-memory mailboxes."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, GSMC_NCCL_SCALARS="1", GSMC_MGPU_BIG_LOG2="16"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count(" ok on 2 ranks") == 8, res.stdout[-2000:]


def _gloo_worker(rank, world, port, N, T, ret):
    """Each rank runs the oracle on ITS shard only, exchanging exactly what the CUDA library exchanges:
    logsumexp triples, per-rank integer weight totals (and, standing in for the peer-memory loads of
    the GPU path, the CDF segments and state rows)."""
    import math
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gen_b200 import shard
    from oracle import closed_forms as cf
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = O.Oracle()
    params = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    ys = cf.simulate_lgssm(T, params, 3)
    first, n = shard.partition(N, world, rank)
    seed = 5

    def gather_obj(x):
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    # init: global-index Philox draws restricted to the shard
    z = orc.normals(seed, 1, first, n)
    x = params[0] + params[1] * z
    lw = np.array([orc.L.orc_logpdf_normal(float(ys[0]), params[5] * xi, params[6]) for xi in x])
    log_ml_est, rho = 0.0, 0
    for t in range(1, T):
        m = float(lw.max())
        tri = (m, float(np.sum(np.exp(lw - m))), float(np.sum(np.exp(2 * (lw - m)))))
        log_total, ess = shard.combine_lse(gather_obj(tri))
        if ess < N * 0.8:
            gmax = max(tr[0] for tr in gather_obj(tri))
            k = orc.L.orc_weight_shift(N)
            q = np.floor(np.array([orc.L.orc_exp(float(v - gmax)) for v in lw]) * 2.0 ** k).astype(np.uint64)
            local_cdf = np.cumsum(q, dtype=np.uint64)
            totals = gather_obj(int(local_cdf[-1]))
            offs, c_n = shard.cdf_offsets(totals)
            cdf_all = np.concatenate([np.asarray(c, dtype=np.uint64) + np.uint64(o) for c, o in zip(gather_obj(local_cdf), offs)])
            x_all = np.concatenate(gather_obj(x))
            # sorted draws as grouped order statistics: a rank generates the Gamma gaps of ITS groups only and
            # exchanges their total; the opening order statistic of its first group is head + lower ranks' totals
            gl = [int(v) for v in orc.gaps(seed, rho, N, first // shard.GROUP, n // shard.GROUP)]
            g_offs, g_sum = shard.cdf_offsets(gather_obj(sum(gl)))
            head = orc.gap_head(seed, rho, N)
            g_all = [int(v) for v in orc.gaps(seed, rho, N, 0, N // shard.GROUP)]
            assert head + g_offs[rank] == head + sum(g_all[:first // shard.GROUP]) and g_sum == sum(g_all)
            anc = orc.search_sorted(cdf_all, seed, rho, N)[first:first + n]     # this rank's output slots
            owners = [shard.owner_of_ancestor(int(a), N, world)[0] for a in anc[[0, -1]]]
            assert owners[0] <= owners[1]
            x = x_all[anc]
            lw = np.zeros(n)
            log_ml_est += log_total - math.log(N)
            rho += 1
        z = orc.normals(seed, t + 1, first, n)
        x = (x * params[2] + params[3]) + params[4] * z
        lw = lw + np.array([orc.L.orc_logpdf_normal(float(ys[t]), params[5] * xi, params[6]) for xi in x])
    m = float(lw.max())
    tri = (m, float(np.sum(np.exp(lw - m))), float(np.sum(np.exp(2 * (lw - m)))))
    log_total, _ = shard.combine_lse(gather_obj(tri))
    ret[rank] = (first, n, x, lw, log_ml_est + log_total - math.log(N), rho)
    dist.destroy_process_group()


def test_sharding_rules_gloo_world2(orc):
    """world_size=2 over gloo on CPU == single-rank oracle (states and log weights bit for bit)."""
    import multiprocessing as mp
    from oracle import closed_forms as cf
    from oracle import oracle as O
    N, T, world = 4096, 8, 2
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, 29611, N, T, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    params = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    ys = cf.simulate_lgssm(T, params, 3)
    pf = orc.particle_filter(O.LGSSM, params, N, seed=5)
    pf.init([ys[0]])
    n_res = 0
    for t in range(1, T):
        n_res += pf.maybe_resample(N * 0.8)
        pf.step([ys[t]])
    assert n_res >= 1
    for r in range(world):
        first, n, x, lw, lml, rho = ret[r]
        assert rho == n_res
        assert np.array_equal(x.view(np.uint64), pf.state()[0, first:first + n].view(np.uint64))
        assert np.array_equal(lw.view(np.uint64), pf.log_weights()[first:first + n].view(np.uint64))
        assert lml == pytest.approx(pf.log_ml_estimate(), rel=1e-12)


def test_shard_helpers():
    from gen_b200 import shard
    assert shard.partition(8192, 2, 1) == (4096, 4096)
    with pytest.raises(ValueError):
        shard.partition(1000, 2, 0)
    offs, total = shard.cdf_offsets([5, 7, 11])
    assert offs == [0, 5, 12] and total == 23
    lt, ess = shard.combine_lse([(0.0, 4.0, 4.0), (0.0, 4.0, 4.0)])
    assert lt == pytest.approx(np.log(8.0)) and ess == pytest.approx(8.0)
    assert shard.merge_lse((-np.inf, 0.0, 0.0), (1.0, 2.0, 3.0)) == (1.0, 2.0, 3.0)
