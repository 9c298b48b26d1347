#!/usr/bin/env python
"""Runs the five BASELINE.json configurations at their per-GPU sizes through the public API and prints one
JSON line each (device-timed; observations simulated here with numpy; no oracle involved):
  cfg1 HMM N=10^4 (the reference's own test)      cfg2 importance sampling, regression, 10^7 samples
  cfg3 LG-SSM T=100 N=2^24                         cfg4 stochastic volatility, residual, T=1000 N=2^22
  cfg5 bearings-only, custom proposal, T=200, N=2^24 per GPU
usage: python scripts/run_configs.py [cfg ...]   (default: all)"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gen_b200 as g  # noqa: E402

PEAK = 6552.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def sim_lgssm(T, p, seed=0):
    m0, s0, a, b, q, c, r = p
    rng = np.random.default_rng(seed)
    x = m0 + s0 * rng.standard_normal()
    ys = []
    for t in range(T):
        if t > 0:
            x = a * x + b + q * rng.standard_normal()
        ys.append(c * x + r * rng.standard_normal())
    return np.array(ys)


def sim_sv(T, p, seed=0):
    mu, phi, sigma = p
    rng = np.random.default_rng(seed)
    h = mu + sigma / math.sqrt(1 - phi * phi) * rng.standard_normal()
    ys = []
    for t in range(T):
        if t > 0:
            h = mu + phi * (h - mu) + sigma * rng.standard_normal()
        ys.append(math.exp(h / 2) * rng.standard_normal())
    return np.array(ys)


def sim_bearings(T, sw=0.001, st=0.005, seed=0, truth=(-0.05, 0.001, 12.0, -0.055)):
    rng = np.random.default_rng(seed)
    x, vx, y, vy = truth
    obs = []
    for t in range(T):
        if t > 0:
            wx, wy = sw * rng.standard_normal(2)
            x, vx, y, vy = x + vx + 0.5 * wx, vx + wx, y + vy + 0.5 * wy, vy + wy
        obs.append(math.atan2(y, x) + st * rng.standard_normal())
    return np.array(obs)


def filter_run(name, model, ys, N, S, resample="multinomial", proposal=None, keep_history=True, reps=2, thr=None):
    T = len(ys)
    st = g.ParticleFilterState(model, N, seed=0, resample=resample, keep_history=keep_history, history_capacity=T)
    thr = N / 2 if thr is None else thr

    def one():
        st.reset()
        st.init([ys[0]], proposal)
        st.run_steps(ys[1:], thr, proposal)
        return st.log_ml_estimate()

    one()
    one()                      # the second occurrence of a run shape is captured into a CUDA graph; later ones replay it
    st.synchronize()
    st.timer_start()
    for _ in range(reps):
        lml = one()
    ms = st.timer_stop() / reps
    n_res = st.stats()["num_resamples"]
    alg = N * (T * (2 * S + 16) + n_res * (2 * S + 36))
    kernel_ms = None
    if os.environ.get("GSMC_CFG_PROFILE"):             # per-kernel-class CUDA-event times (per-call API, same workload)
        st.set_profiling(True)
        st.reset()
        st.init([ys[0]], proposal)
        for t in range(1, T):
            st.maybe_resample(thr)
            st.step([ys[t]], proposal)
        prof = st.stats()
        kernel_ms = {k[3:]: round(prof[k], 3) for k in prof if k.startswith("ms_")}
        st.set_profiling(False)
    st.close()
    return {"kernel_ms_profile_pass": kernel_ms,"config": name, "particles": N, "time_steps": T, "resample": resample, "resamples_per_run": n_res,
            "ms_per_run": ms, "particle_steps_per_s": N * T / (ms * 1e-3), "log_ml": lml,
            "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": alg / (ms * 1e-3) / 1e9 / PEAK}


def cfg1():
    prior = [0.2, 0.3, 0.5]
    emis = [[0.1, 0.2, 0.7], [0.2, 0.7, 0.1], [0.7, 0.2, 0.1]]
    trans = [[0.4, 0.4, 0.2], [0.2, 0.3, 0.5], [0.9, 0.05, 0.05]]
    obs = np.array([1, 1, 2, 3], dtype=float)
    # the constructor takes the reference's (Julia, column-per-state) matrices: emission_dists[x, z], transition_dists[z, z_prev]
    model = g.HMM(prior, np.array(emis).T, np.array(trans).T)
    N = 10 ** 4
    out = filter_run("cfg1 HMM (test/inference/particle_filter.jl:52-81), resample every step", model, obs, N, 8, thr=N, reps=20)
    out["log_ml_exact"] = -4.87645083351704
    out["abs_err"] = abs(out["log_ml"] - out["log_ml_exact"])
    return out


def cfg2():
    # examples/regression/quickstart.jl:26-27
    xs = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0]
    ys = [8.23, 5.87, 3.99, 2.59, 0.23, -0.66, -3.53, -6.91, -7.24, -9.90]
    model = g.LinearRegression(2.0, 10.0, 1.0)
    n = 10 ** 7
    cm = g.choicemap(*[("y-%d" % (i + 1), float(y)) for i, y in enumerate(ys)])
    g.importance_sampling(model, (xs,), cm, 1 << 16)
    t0 = time.perf_counter()
    traces, lw, lml = g.importance_sampling(model, (xs,), cm, n)
    dt = time.perf_counter() - t0
    traces._state.close()
    return {"config": "cfg2 importance_sampling, Bayesian linear regression (quickstart.jl data), 10^7 samples (wall clock incl. D2H of the 80 MB weight vector)",
            "samples": n, "ms": dt * 1e3, "samples_per_s": n / dt, "log_ml": lml, "log_ml_exact": -18.150487182903948,
            "logsumexp_normalised": float(np.log(np.sum(np.exp(lw - lw.max()))) + lw.max())}


def cfg3():
    p = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    return filter_run("cfg3 LG-SSM T=100 N=2^24, bootstrap, multinomial at ESS<N/2", g.LinearGaussianSSM(*p), sim_lgssm(100, p), 1 << 24, 8, reps=3)


def cfg4():
    p = [-1.0, 0.97, 0.2]
    return filter_run("cfg4 stochastic volatility T=1000 N=2^22, residual resampling at ESS<N/2", g.StochasticVolatility(*p), sim_sv(1000, p),
                      1 << 22, 8, resample="residual", reps=2)


def cfg5():
    model = g.BearingsOnly()
    ys = sim_bearings(200)
    return filter_run("cfg5 bearings-only T=200 N=2^24 (one GPU's shard of 2^27), custom proposal, history dropped", model, ys, 1 << 24, 32,
                      proposal=model.custom_proposal(), keep_history=False, reps=1)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    for c in todo:
        try:
            print(json.dumps(globals()[c]()), flush=True)
        except Exception as e:  # keep going: one configuration must not hide the others
            print(json.dumps({"config": c, "error": repr(e)}), flush=True)
