"""Worker of tests/test_multigpu.py: one process per GPU (torchrun), NCCL inside libgensmc.so.
Every rank runs the full CPU oracle and compares its own shard bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import gen_b200 as g
    from gen_b200.distributed import Communicator
    from oracle import closed_forms as cf
    from oracle import oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator(dist, rank, world, device=local)
    orc = O.Oracle()

    def bits(a):
        return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)

    def same(a, b, what):
        a, b = bits(a).reshape(-1), bits(b).reshape(-1)
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, "[rank %d] %s: %d of %d values differ, first at %s" % (rank, what, bad.size, a.size, bad[:8])

    LG = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    SVP = [-1.0, 0.97, 0.2]
    big = int(os.environ.get("GSMC_MGPU_BIG_LOG2", "20"))          # particles per rank of the large cases
    cases = [
        (O.LGSSM, g.LinearGaussianSSM(*LG), LG, cf.simulate_lgssm(16, LG, 3), 0, 1024 * 16, 12, 0.8, "multinomial"),
        (O.BEARINGS, g.BearingsOnly(), cf.BEARINGS_PARAMS, cf.simulate_bearings(16), 1, 1024 * 16, 12, 0.8, "multinomial"),
        (O.SV, g.StochasticVolatility(*SVP), SVP, cf.simulate_sv(16, SVP, 0), 0, 1024 * 16, 12, 0.8, "multinomial"),
        (O.HMM, g.HMM(cf.HMM_PRIOR, cf.HMM_EMISSION, cf.HMM_TRANSITION), list(cf.hmm_params()), np.array(cf.HMM_OBS * 4, dtype=float), 0, 1024 * 16, 12, 0.8, "multinomial"),
        # 2^20 particles per rank: many tiles per segment, windows across segment and rank boundaries, skewed weights
        (O.LGSSM, g.LinearGaussianSSM(*LG), LG, cf.simulate_lgssm(16, LG, 3), 0, 1 << big, 7, 0.4, "multinomial"),
        # the cfg-5 shape (bearings-only, custom proposal, D = 4 rows gathered across shard boundaries) at 2^18 per rank
        (O.BEARINGS, g.BearingsOnly(), cf.BEARINGS_PARAMS, cf.simulate_bearings(16), 1, 1 << (big - 2), 7, 0.5, "multinomial"),
        # residual resampling on a sharded filter (the cfg-4 shape): copies and draws are stored into the rank that owns the slot
        (O.SV, g.StochasticVolatility(*SVP), SVP, cf.simulate_sv(16, SVP, 0), 0, 1024 * 16, 12, 0.8, "residual"),
        (O.LGSSM, g.LinearGaussianSSM(*LG), LG, cf.simulate_lgssm(16, LG, 3), 0, 1 << (big - 1), 7, 0.4, "residual"),
    ]
    for fam, model, params, ys, prop, n_per, T, thr, scheme in cases:
        N = n_per * world
        oscheme = 1 if scheme == "residual" else 0
        st = g.ParticleFilterState(model, N, seed=5, resample=scheme, keep_history=True, history_capacity=T, device=local, comm=comm)
        n, first = st.num_local, st.first_global
        assert n == N // world and first == rank * n
        pf = orc.particle_filter(fam, params, N, seed=5, keep_history=True)
        proposal = model.custom_proposal() if prop else None
        st.init([ys[0]], proposal)
        pf.init([ys[0]], proposal=prop)
        sl = slice(first, first + n)
        same(st.log_weights(), pf.log_weights()[sl], "init log weights")
        same(st.state(), pf.state()[:, sl], "init state")
        n_res = 0
        for t in range(1, T):
            dg, do = st.maybe_resample(N * thr), pf.maybe_resample(N * thr, scheme=oscheme)
            assert dg == do, (t, st.last_ess, pf.last_ess)
            assert abs(st.last_ess - pf.last_ess) <= 1e-10 * pf.last_ess
            if dg:
                n_res += 1
                assert np.array_equal(st.ancestors(), pf.parents()[sl]), "ancestors differ at t=%d" % t
                same(st.state(), pf.state()[:, sl], "gathered state at t=%d" % t)
            st.step([ys[t]], proposal)
            pf.step([ys[t]], proposal=prop)
            same(st.state(), pf.state()[:, sl], "state after step t=%d (resampled=%s)" % (t, dg))
            same(st.log_weights(), pf.log_weights()[sl], "log weights after step t=%d (resampled=%s)" % (t, dg))
        assert n_res >= (2 if thr >= 0.8 else 1)
        assert np.array_equal(bits(st.state()), bits(pf.state()[:, sl]))
        a, b = st.log_ml_estimate(), pf.log_ml_estimate()
        assert abs(a - b) <= 1e-11 * abs(b), (a, b)
        for t in (1, T // 2, T):
            assert np.array_equal(bits(st.state(t)), bits(pf.history(t)[:, sl])), "history differs at t=%d" % t
        # sample_unweighted_traces (particle_filter.jl:62-70) on the sharded filter: collective, every rank gets the
        # same GLOBAL indices as the oracle and reads the trajectories (its own rows and its peers') identically
        ig, io = st.sample_unweighted(333), pf.sample_unweighted(333)
        assert np.array_equal(ig, io), "sample_unweighted differs"
        assert io.min() < n and io.max() >= (world - 1) * n, "the draw should touch the first and the last shard"
        tr = st.trajectories(ig[:64])
        for t in (1, T // 2, T):
            assert np.array_equal(bits(tr[:, t - 1, :].T), bits(pf.history(t)[:, io[:64]])), "sampled trajectories differ at t=%d" % t
        trs = g.sample_unweighted_traces(st, 7)
        assert len(trs) == 7
        # the sync-free loop gives the same answer
        st2 = g.ParticleFilterState(model, N, seed=5, resample=scheme, keep_history=False, device=local, comm=comm)
        st2.init([ys[0]], proposal)
        st2.run_steps(ys[1:T], N * thr, proposal)
        assert st2.log_ml_estimate() == a
        assert np.array_equal(bits(st2.log_weights()), bits(st.log_weights()))
        if scheme == "residual":
            # a repeated run shape is captured into a CUDA graph (conditional node per step) on its second occurrence:
            # every rank replays its own graph, the exchanges inside carry device-side sequence numbers
            for rep in range(3):
                st2.reset()
                st2.init([ys[0]], proposal)
                st2.run_steps(ys[1:T], N * thr, proposal)
                assert st2.log_ml_estimate() == a and np.array_equal(bits(st2.log_weights()), bits(st.log_weights())), rep
            # first run above: plain launches; rep 0: captured + launched; reps 1, 2: replayed
            assert st2.stats()["graph_replays"] == 3, st2.stats()["graph_replays"]
        st.close()
        st2.close()
        dist.barrier()
        if rank == 0:
            print("family %d %s (N = %d x %d, T = %d) ok on %d ranks: log_ml %.12f, %d resamples" % (fam, scheme, world, n_per, T, world, a, n_res), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
