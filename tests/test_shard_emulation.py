"""The sharded (multi-GPU) data path on ONE GPU: R logical ranks of one filter as R handles on one device
(gsmc_group_*, gen_b200.distributed.LocalShardGroup). Rank offsets of the integer CDF and of the group gaps, CDF
windows and ancestor gathers that cross shard boundaries, history walks through "peer" columns and the collective
sample_unweighted are the code that runs on R GPUs; only the scalar exchanges are direct reads. Every shard is
compared bit for bit with the FULL CPU oracle, like tests/mgpu_worker.py does on R real GPUs."""
import numpy as np
import pytest

import gen_b200 as g
from gen_b200.distributed import LocalShardGroup
from oracle import closed_forms as cf
from oracle import oracle as O

pytestmark = pytest.mark.gpu

LG = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def families():
    return {
        "lgssm": (O.LGSSM, g.LinearGaussianSSM(*LG), LG, cf.simulate_lgssm(16, LG, 3), 0),
        "lgssm_prop": (O.LGSSM, g.LinearGaussianSSM(*LG), LG, cf.simulate_lgssm(16, LG, 3), 1),
        "bearings": (O.BEARINGS, g.BearingsOnly(), cf.BEARINGS_PARAMS, cf.simulate_bearings(16), 1),
        "sv": (O.SV, g.StochasticVolatility(-1.0, 0.97, 0.2), [-1.0, 0.97, 0.2], cf.simulate_sv(16, [-1.0, 0.97, 0.2], 0), 0),
        "hmm": (O.HMM, g.HMM(cf.HMM_PRIOR, cf.HMM_EMISSION, cf.HMM_TRANSITION), list(cf.hmm_params()), np.array(cf.HMM_OBS * 4, dtype=float), 0),
    }


def run_group(orc, fam, model, params, ys, prop, world, n_per, T, thr_frac, dtype="f64", scheme="multinomial"):
    N = n_per * world
    oscheme = 1 if scheme == "residual" else 0
    grp = LocalShardGroup(world)
    shards = [g.ParticleFilterState(model, N, seed=5, dtype=dtype, resample=scheme, keep_history=True, history_capacity=T, comm=grp.rank(r)) for r in range(world)]
    pf = orc.particle_filter(fam, params, N, seed=5, keep_history=True)
    proposal = model.custom_proposal() if prop else None
    for r, st in enumerate(shards):
        assert st.num_local == n_per and st.first_global == r * n_per
    grp.init([ys[0]], proposal)
    pf.init([ys[0]], proposal=prop)
    sls = [slice(r * n_per, (r + 1) * n_per) for r in range(world)]
    for st, sl in zip(shards, sls):
        assert np.array_equal(bits(st.log_weights()), bits(pf.log_weights()[sl]))
        assert np.array_equal(bits(st.state()), bits(pf.state()[:, sl]))
    n_res, remote = 0, 0
    for t in range(1, T):
        dg, do = grp.maybe_resample(N * thr_frac), pf.maybe_resample(N * thr_frac, scheme=oscheme)
        assert dg == do, (t, shards[0].last_ess, pf.last_ess)
        assert abs(shards[0].last_ess - pf.last_ess) <= 1e-10 * pf.last_ess
        if dg:
            n_res += 1
            par = pf.parents()
            for r, (st, sl) in enumerate(zip(shards, sls)):
                anc = st.ancestors()
                assert np.array_equal(anc, par[sl]), "rank %d: ancestors differ at t=%d" % (r, t)
                remote += int(np.sum((anc < r * n_per) | (anc >= (r + 1) * n_per)))
                assert np.array_equal(bits(st.state()), bits(pf.state()[:, sl])), "rank %d: gathered state differs at t=%d" % (r, t)
        grp.step([ys[t]], proposal)
        pf.step([ys[t]], proposal=prop)
        for r, (st, sl) in enumerate(zip(shards, sls)):
            assert np.array_equal(bits(st.state()), bits(pf.state()[:, sl])), "rank %d: state differs after step %d" % (r, t)
            assert np.array_equal(bits(st.log_weights()), bits(pf.log_weights()[sl])), "rank %d: log weights differ after step %d" % (r, t)
    lml = pf.log_ml_estimate()
    for st in shards:
        assert abs(st.log_ml_estimate() - lml) <= 1e-11 * abs(lml)
    for t in (1, T // 2, T):
        for st, sl in zip(shards, sls):
            assert np.array_equal(bits(st.state(t)), bits(pf.history(t)[:, sl])), "history differs at t=%d" % t
    return grp, shards, pf, n_res, remote


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("family", ["lgssm", "bearings", "sv", "lgssm_prop", "hmm"])
def test_emulated_shards_match_oracle(orc, world, family):
    fam, model, params, ys, prop = families()[family]
    grp, shards, pf, n_res, remote = run_group(orc, fam, model, params, ys, prop, world, 2048 * 4, 12, 0.8)
    assert n_res >= 2
    assert remote > 0, "no ancestor crossed a shard boundary: the cross-rank gather was not exercised"
    # collective sample_unweighted + trajectories through the peers' columns
    N = shards[0].num_particles
    ig, io = grp.sample_unweighted(333), pf.sample_unweighted(333)
    assert np.array_equal(ig, io)
    assert io.min() < shards[0].num_local and io.max() >= N - shards[0].num_local
    T = shards[0].T
    tr = shards[world - 1].trajectories(ig[:64])
    for t in (1, T // 2, T):
        assert np.array_equal(bits(tr[:, t - 1, :].T), bits(pf.history(t)[:, io[:64]]))
    grp.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("family", ["sv", "lgssm", "bearings"])
def test_emulated_shards_residual_resampling(orc, world, family):
    """Residual resampling on a sharded filter (BASELINE.json configs[3] names the scheme): the deterministic copies of a
    rank's particles and its share of the multinomial draws land in output slots that other ranks own."""
    fam, model, params, ys, prop = families()[family]
    grp, shards, pf, n_res, remote = run_group(orc, fam, model, params, ys, prop, world, 2048 * 4, 12, 0.8, scheme="residual")
    assert n_res >= 2 and remote > 0
    grp.close()


def test_emulated_shards_residual_large(orc):
    fam, model, params, ys, prop = families()["lgssm"]
    grp, shards, pf, n_res, remote = run_group(orc, fam, model, params, cf.simulate_lgssm(8, LG, 3), prop, 4, 1 << 19, 8, 0.4, scheme="residual")
    assert n_res >= 1 and remote > 0
    grp.close()


def test_emulated_shards_large_and_skewed(orc):
    """2^20 particles per rank on 4 ranks (the oracle still finishes in seconds): many tiles per segment, windows that
    straddle segment and rank boundaries; a low ESS threshold makes the weights skewed when a resample finally fires."""
    fam, model, params, ys, prop = families()["lgssm"]
    grp, shards, pf, n_res, remote = run_group(orc, fam, model, params, cf.simulate_lgssm(8, LG, 3), prop, 4, 1 << 20, 8, 0.35)
    assert n_res >= 1 and remote > 0
    grp.close()


def test_emulated_shards_replay_uniforms(orc):
    """north_star protocol on a sharded filter: exported uniforms (one per output slot) -> bit-exact ancestors."""
    world, n_per = 2, 4096
    N = world * n_per
    model = g.LinearGaussianSSM(*LG)
    ys = cf.simulate_lgssm(4, LG, 3)
    grp = LocalShardGroup(world)
    shards = [g.ParticleFilterState(model, N, seed=9, keep_history=True, history_capacity=4, comm=grp.rank(r)) for r in range(world)]
    pf = orc.particle_filter(O.LGSSM, LG, N, seed=9, keep_history=True)
    grp.init([ys[0]])
    pf.init([ys[0]])
    u = np.random.default_rng(1).random(N)
    for r, st in enumerate(shards):
        st.set_replay(uniforms=u[r * n_per:(r + 1) * n_per])
    assert grp.maybe_resample(N) and pf.maybe_resample(N, u_replay=u)
    for r, st in enumerate(shards):
        assert np.array_equal(st.ancestors(), pf.parents()[r * n_per:(r + 1) * n_per])
    grp.step([ys[1]])
    pf.step([ys[1]])
    for r, st in enumerate(shards):
        assert np.array_equal(bits(st.state()), bits(pf.state()[:, r * n_per:(r + 1) * n_per]))
    grp.close()
